#!/usr/bin/env python
"""CUDA-event timings of the rough-terrain path, to run on the GPU box WITHOUT a profiler:
  * hg_terrain_get_heights alone at N = 4096 / 65536 (187 points per env; algorithmic bytes = the (N,187) fp32 result
    + 52 B of root state per env; the int16 gathers hit a few cache lines per env and stay in L1 / L2),
  * hg_terrain_priv_frames alone (3 x 892 fp32 in + out per env),
  * a whole env.step() on rough terrain (two fused launches + curriculum + heights + critic frames) next to the plane
    step (one fused launch), synthetic physics, N = 4096.
L2 is flushed between timed launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def terrain_cfg():
    from humanoid.envs import XBotLCfg

    class Cfg(XBotLCfg):
        class env(XBotLCfg.env):
            single_num_privileged_obs = XBotLCfg.env.num_observations + 17 * 11
            num_privileged_obs = int(XBotLCfg.env.c_frame_stack * single_num_privileged_obs)

        class terrain(XBotLCfg.terrain):
            mesh_type, curriculum, measure_heights = "trimesh", True, True      # XBotLCfg's commented-out alternative
    cfg = Cfg()
    cfg.seed = 5
    return cfg


def timed(fn, reps, flush):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts), sum(ts) / len(ts)


def main():
    from parity_utils import make_env
    from humanoid import _native as nat
    dev = torch.device("cuda:0")
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    np.random.seed(5)
    for N in (4096, 65536):
        env = make_env(N, physics="synthetic", cfg=terrain_cfg())
        hs = env.height_samples
        print(f"terrain {tuple(hs.shape)} int16 ({hs.numel() * 2 / 1e6:.1f} MB), N={N}")
        for _ in range(3):
            env.step(torch.randn(N, 12, device=dev))
        best, mean = timed(env._get_heights, 10, flush)
        b = N * (187 * 4 + 52)
        print(f"  get_heights: {best:.2f} us best, {mean:.2f} us mean -> {b / mean * 1e-3:.1f} GB/s algorithmic ({b / 1e6:.2f} MB)")
        src = env._privh_pp[0]

        def frames():
            nat.check(nat.lib.hg_terrain_priv_frames(
                env.obs_buf.data_ptr(), env.obs_buf.stride(0), env.num_obs, env.root_states.data_ptr(), env._heights.data_ptr(), 187,
                5.0, 18.0, env.reset_buf.data_ptr(), src.data_ptr(), env._privh_pp[1].data_ptr(), src.stride(0), 3, N,
                nat.stream_ptr(0)), "priv_frames")
        best, mean = timed(frames, 10, flush)
        b = N * (2 * 892 * 4 + 705 * 4 + 187 * 4 + 3 * 892 * 4)
        print(f"  priv_frames: {best:.2f} us best, {mean:.2f} us mean -> {b / mean * 1e-3:.1f} GB/s algorithmic ({b / 1e6:.2f} MB)")
        if N == 4096:
            plane = make_env(N, physics="synthetic")
            for e, name in ((env, "rough terrain (trimesh, curriculum, heights)"), (plane, "plane (XBotLCfg default)")):
                a = torch.randn(N, 12, device=dev)
                for _ in range(5):
                    e.step(a)
                best, mean = timed(lambda e=e, a=a: e.step(a), 20, flush)
                print(f"  env.step() {name}: {best:.1f} us best, {mean:.1f} us mean (eager launches)")
        del env
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
