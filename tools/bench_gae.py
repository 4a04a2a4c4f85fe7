#!/usr/bin/env python
"""CUDA-event timing of RolloutStorage.compute_returns' kernel pair (hg_gae + advantage normalisation) at T = 60:
    HG_GAE=scan python tools/bench_gae.py      (warp scan over time, default)
    HG_GAE=serial python tools/bench_gae.py    (one thread per env walks the 60 steps)
Algorithmic bytes per sample: 9 read (reward, value, done) + 8 written (return, advantage) + 8 for the normalisation pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    from humanoid.algo import RolloutStorage
    dev = torch.device("cuda:0")
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for N in (4096, 16384, 65536):
        T = 60
        st = RolloutStorage(N, T, [4], [4], [12], "cuda:0")
        st.rewards.uniform_(), st.values.normal_()
        st.dones.copy_((torch.rand(T, N, 1, device=dev) < 0.02).byte())
        lv = torch.randn(N, 1, device=dev)
        ts = []
        for i in range(13):
            flush.zero_()
            e0.record()
            st.compute_returns(lv, 0.994, 0.9)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        b = N * T * 25
        print(f"HG_GAE={os.environ.get('HG_GAE', 'scan')} N={N}: {min(ts):.2f} us best, {sum(ts) / len(ts):.2f} us mean -> "
              f"{b / (sum(ts) / len(ts)) * 1e-3:.1f} GB/s algorithmic ({b / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
