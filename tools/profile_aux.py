#!/usr/bin/env python
"""Launch sequence for the ncu capture of the round-2 kernels outside the three hot launches (profiles/r02d_aux_*):
the rough-terrain kernels (hg_terrain_get_heights / reset_prepare / priv_frames around the two-launch env step), the
warp-scan GAE + advantage normalisation, and the native minibatch permutation.

    ncu --set full --clock-control none --import-source on -k regex:"get_heights|reset_prepare|priv_frames|gae_|adv_normalise|randperm" \\
        -c 12 -o gpurun_out/r02d_aux python tools/profile_aux.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from bench_terrain import terrain_cfg
    from parity_utils import make_env
    from humanoid import _native as nat
    from humanoid.algo import RolloutStorage
    N = int(os.environ.get("HG_AUX_ENVS", "4096"))
    dev = torch.device("cuda:0")
    np.random.seed(5)
    env = make_env(N, physics="synthetic", cfg=terrain_cfg())
    env.episode_length_buf = torch.randint(0, 2400, (N,), device=dev)
    env.episode_length_buf[::50] = 2400
    for _ in range(2):
        env.step(torch.randn(N, 12, device=dev))
    st = RolloutStorage(N, 60, [4], [4], [12], "cuda:0")
    st.rewards.uniform_(), st.values.normal_()
    st.dones.copy_((torch.rand(60, N, 1, device=dev) < 0.02).byte())
    st.compute_returns(torch.randn(N, 1, device=dev), 0.994, 0.9)
    perm = torch.empty(N * 60, dtype=torch.int64, device=dev)
    nat.check(nat.lib.hg_randperm(N * 60, 7, 0, perm.data_ptr(), nat.stream_ptr(0)), "hg_randperm")
    torch.cuda.synchronize()
    print("aux launches done")


if __name__ == "__main__":
    main()
