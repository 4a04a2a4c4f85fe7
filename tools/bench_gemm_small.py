"""Micro-benchmark: the rollout-sized GEMMs (M = 4096) on both tensor-core engines, timed inside a CUDA graph
(20 back-to-back launches per shape), plus the fused PPO.act kernel alone.  Prints us per launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from humanoid import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.Stream()


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 5 / reps


def tc(M, N, K, presplit):
    ld = (K + 3) // 4 * 4
    X = torch.randn(M, ld, device=dev)
    W = torch.randn(N, ld, device=dev)
    Wlo = torch.empty_like(W)
    nat.check(nat.lib.hg_tf32_residual(W.data_ptr(), Wlo.data_ptr(), W.numel(), 0))
    b = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev)
    d = nat.Gemm()
    d.A, d.B, d.C, d.bias = X.data_ptr(), W.data_ptr(), C.data_ptr(), b.data_ptr()
    d.M, d.N, d.K, d.lda, d.ldb, d.ldc = M, N, K, ld, ld, N
    d.epilogue, d.passes, d.split_k, d.trust_hw_truncation = 2, 3, 1, 1
    d.B_lo = Wlo.data_ptr() if presplit else None
    keep = (X, W, Wlo, b, C)

    def fn():
        nat.check(nat.lib.hg_gemm_tf32(d, torch.cuda.current_stream().cuda_stream))
    return timed(fn), keep


def bf3(M, N, K):
    ld = (K + 7) // 8 * 8
    X = torch.zeros(2, M, ld, dtype=torch.int16, device=dev)
    W = torch.zeros(2, N, ld, dtype=torch.int16, device=dev)
    b = torch.randn(N, device=dev)
    Cs = torch.zeros(2, M, N, dtype=torch.int16, device=dev)
    d = nat.GemmSplit()
    d.A, d.B, d.Cs, d.bias = nat.Split.of(X), nat.Split.of(W), nat.Split.of(Cs), b.data_ptr()
    d.M, d.N, d.K, d.epilogue, d.split_k = M, N, K, 2, 1
    keep = (X, W, b, Cs)

    def fn():
        nat.check(nat.lib.hg_gemm_bf16x3(d, torch.cuda.current_stream().cuda_stream))
    return timed(fn), keep


M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
print(f"M = {M}")
for N, K in ((512, 705), (512, 352), (512, 64), (256, 512), (128, 256), (768, 219), (256, 768)):
    a, _ = tc(M, N, K, False)
    b_, _ = tc(M, N, K, True)
    c, _ = bf3(M, N, K)
    print(f"N={N:4d} K={K:4d}:  3xTF32 in-kernel split {a:7.1f} us   3xTF32 B_lo by TMA {b_:7.1f} us   bf16x3 presplit {c:7.1f} us")

from humanoid.algo import ActorCritic  # noqa: E402
ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128]).cuda()
ac.flat_params()
obs = torch.randn(M, 736, device=dev)[:, :705]
cobs = torch.randn(M, 224, device=dev)[:, :219]
mu, val = torch.empty(M, 12, device=dev), torch.empty(M, 1, device=dev)
act, lp, sg = torch.empty(M, 12, device=dev), torch.empty(M, device=dev), torch.empty(M, 12, device=dev)
sample = dict(std=ac.std, actions=act, log_prob=lp, sigma=sg, seed=1, step=0)
ac.refresh_lo()
print(f"fused PPO.act kernel alone: {timed(lambda: ac.native_act(obs, cobs, mu, val, sample)):7.1f} us")
os.environ["HG_FUSED_ACT"] = "0"


def sep():
    ac.native_forward("critic", cobs, val)
    ac.native_forward("actor", obs, mu, sample=sample)


print(f"per-layer launches (8 kernels, one stream): {timed(sep):7.1f} us")
