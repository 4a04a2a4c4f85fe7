"""Micro-benchmark of hg_gemm_bf16x3 at the update's shapes (M = 61440): epilogue variants, to separate main-loop time from
epilogue time.  Prints us per launch (CUDA events, 10 back-to-back launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from humanoid import _native as nat  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def planes(r, c):
    t = torch.randint(-3000, 3000, (2, r, (c + 7) // 8 * 8), dtype=torch.int16, device=dev)
    return t


M = 61440
for name, N, K, a_mn, b_mn in (("fwd  actor L1", 512, 705, 0, 0), ("fwd  critic L1", 768, 219, 0, 0), ("fwd  actor L2", 256, 512, 0, 0),
                                ("fwd  L3", 128, 256, 0, 0), ("dgrad critic L2", 768, 256, 0, 1), ("dgrad actor L2", 512, 256, 0, 1)):
    A = planes(M, K)
    B = planes(N, K) if not b_mn else planes(K, N)
    Cs = torch.zeros(2, M, N, dtype=torch.int16, device=dev)
    Hs = planes(M, N)
    C = torch.zeros(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    cs = torch.zeros(N, device=dev)
    out = []
    for label, epi in (("split+ELU store", 2), ("math only, no store", 6), ("fp32 store", 0), ("dgrad (H, colsum)", 3), ("dgrad no colsum", 33)):
        if (epi in (3, 33)) != bool(b_mn):
            continue
        d = nat.GemmSplit()
        d.A, d.B, d.Cs, d.Hs = nat.Split.of(A), nat.Split.of(B), nat.Split.of(Cs), nat.Split.of(Hs)
        d.C, d.ldc, d.bias = C.data_ptr(), N, bias.data_ptr()
        d.colsum = cs.data_ptr() if epi == 3 else None
        d.M, d.N, d.K, d.a_mn_major, d.b_mn_major, d.epilogue, d.split_k = M, N, K, a_mn, b_mn, (3 if epi == 33 else epi), 1
        t = timed(lambda: nat.check(nat.lib.hg_gemm_bf16x3(d, torch.cuda.current_stream().cuda_stream)))
        out.append(f"{label} {t:7.1f}")
    fl = 2.0 * M * N * K * 3
    print(f"{name:16s} N={N:4d} K={K:4d}: " + " | ".join(out) + f"   [us; 100 % tensor pipe = {fl / 2.25e15 * 1e6:.0f} us]")
