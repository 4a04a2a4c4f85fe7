"""Where hg_gemm_bf16x3's time goes at the update's shapes (M = 61440): per-CTA cycle sums written by the kernel itself
(hg_gemm_bf16x3_set_trace) -- the MMA thread's loop, how much of it is spent waiting for operands (TMA) or for a drained
accumulator (epilogue), the TMA thread's wait for a free stage, and the epilogue's wait/busy split.

    python tools/bf3_trace.py            (HG_BF3_PAIR=0/1, HG_BF3_STAGES=n to vary)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from humanoid import _native as nat  # noqa: E402

dev = torch.device("cuda:0")


def planes(r, c):
    return torch.randint(-3000, 3000, (2, r, (c + 7) // 8 * 8), dtype=torch.int16, device=dev)


M = 61440
trace_buf = torch.zeros(148 * 8 + 148 * 64 * 4, dtype=torch.int64, device=dev)
trace = trace_buf[:148 * 8].view(148, 8)
stamps = trace_buf[148 * 8:].view(148, 64, 4)
shapes = (("fwd actor L1", 512, 705, 0, 0, 2), ("fwd actor L1 nostore", 512, 705, 0, 0, 6), ("fwd critic L1", 768, 219, 0, 0, 2),
          ("fwd actor L2", 256, 512, 0, 0, 2), ("dgrad critic L2", 768, 256, 0, 1, 3), ("dgrad actor L2", 512, 256, 0, 1, 3),
          ("wgrad actor L1 (split-K)", 705, 61440, 1, 1, 4))
print(f"pair={os.environ.get('HG_BF3_PAIR', '1')} stages cap={os.environ.get('HG_BF3_STAGES', '-')}")
print(f"{'shape':26s} {'us':>7s} {'ideal':>6s} | per busiest CTA, k-cycles: {'loop':>6s} {'w.full':>7s} {'w.acc':>6s} {'tma w.empty':>11s} {'epi wait':>8s} {'epi busy':>8s} items kb  mma-only")
for name, N, K, a_mn, b_mn, epi in shapes:
    Mx = M
    if epi == 4:                                   # wgrad: C[N x Nout] = A^T B over the batch
        Mx, N, K = 512, 705, 61440
        A = planes(K, Mx)
        B = planes(K, N)
    else:
        A = planes(Mx, K)
        B = planes(N, K) if not b_mn else planes(K, N)
    Cs = torch.zeros(2, Mx, (N + 7) // 8 * 8, dtype=torch.int16, device=dev)
    Hs = planes(Mx, N)
    Cf = torch.zeros(Mx, (N + 7) // 8 * 8, device=dev)
    bias = torch.randn(N, device=dev)
    cs = torch.zeros(N, device=dev)
    d = nat.GemmSplit()
    d.A, d.B, d.Cs, d.Hs = nat.Split.of(A), nat.Split.of(B), nat.Split.of(Cs), nat.Split.of(Hs)
    d.C, d.ldc, d.bias = Cf.data_ptr(), Cf.shape[1], bias.data_ptr()
    d.colsum = cs.data_ptr() if epi == 3 else None
    d.M, d.N, d.K, d.a_mn_major, d.b_mn_major, d.epilogue = Mx, N, K, a_mn, b_mn, epi
    d.split_k = 1 if epi != 4 else 74
    st = torch.cuda.current_stream().cuda_stream

    def run():
        nat.check(nat.lib.hg_gemm_bf16x3(d, st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    trace_buf.zero_()
    nat.lib.hg_gemm_bf16x3_set_trace(trace_buf.data_ptr())
    run()
    torch.cuda.synchronize()
    nat.lib.hg_gemm_bf16x3_set_trace(None)
    t = trace.cpu()
    lead = t[t[:, 0] > 0]
    i = int(lead[:, 0].argmax())
    r = lead[i]
    # the TMA / epilogue rows of the same CTA (pair: the leader's)
    kb = int(r[7])
    n_mma = kb * 12
    bn = min(256, (N + 63) // 64 * 64)
    mma_clk = n_mma * (bn / 2)                     # 128 x BN x 16 bf16 MMA = BN / 2 clocks at 8192 flop/clk/SM
    fl = 2.0 * Mx * N * K * 3
    print(f"{name:26s} {us:7.1f} {fl / 2.25e15 * 1e6:6.0f} | {'':26s} {r[0] / 1e3:6.1f} {r[1] / 1e3:7.1f} {r[2] / 1e3:6.1f} "
          f"{t[0::2, 3].max() / 1e3:5.0f}/{t[1::2, 3].max() / 1e3:5.0f} {t[:, 4].max() / 1e3:8.1f} {t[:, 5].max() / 1e3:8.1f} {int(r[6]):5d} {kb:3d} {mma_clk / 1e3:7.1f}")
    if os.environ.get("HG_BF3_STAMPS") == "1" and name.startswith("fwd actor L1 nostore"):
        sp = stamps.cpu()
        t0 = int(sp[0, 0, 0])
        print("   k-block stamps of CTAs 0 (leader) and 1 (peer), ns since the first: free / issued / relay / full(leader MMA thread)")
        for it in range(20, 32):
            a, b = sp[0, it], sp[1, it]
            print(f"   it {it:2d}: L free {int(a[0]) - t0:7d} issued {int(a[1]) - t0:7d} | P free {int(b[0]) - t0:7d} issued {int(b[1]) - t0:7d} "
                  f"relay {int(b[2]) - t0:7d} | full {int(a[3]) - t0:7d}")
