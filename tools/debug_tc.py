import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "humanoid-gym_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gemm_tc_gpu import _run, _rel
torch.set_printoptions(precision=4, linewidth=200)
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 128, 128, 32
A = torch.randn(M, K, device="cuda", generator=g)
Bt = torch.randn(K, N, device="cuda", generator=g)     # (K, N): MN-major B
ref = A.double() @ Bt.double()
for passes in (1, 3):
    C = _run(A, Bt, M, N, K, 0, 1, passes)
    print("B mn-major passes", passes, "rel", _rel(C, ref), "nonzero", int((C != 0).sum()), "nan", int(torch.isnan(C).sum()))
    print(C[:3, :6]); print(ref[:3, :6].float())
At = torch.randn(K, M, device="cuda", generator=g)     # (K, M): MN-major A
Bk = torch.randn(N, K, device="cuda", generator=g)
ref = At.double().t() @ Bk.double().t()
C = _run(At, Bk, M, N, K, 1, 0, 1)
print("A mn-major rel", _rel(C, ref), "nonzero", int((C != 0).sum()))
print(C[:3, :6]); print(ref[:3, :6].float())
# does the result match some other contraction?
C2 = _run(A, Bt, M, N, K, 0, 1, 1)
cands = {"A@Bt": A @ Bt}
print({k: _rel(C2, v.double()) for k, v in cands.items()})
