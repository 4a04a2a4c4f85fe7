#!/usr/bin/env python
"""Short, deterministic launch sequence for ncu captures (see profiles/README.md):
  env   : 4 x hg_env_post_physics at --num-envs (synthetic physics state)
  mlp   : 2 x one PPO minibatch step (gather, MLP fwd, loss, MLP bwd, clip+Adam) at B = num_envs*60/4
Also prints CUDA-event timings of the same launches when run WITHOUT a profiler."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["env", "mlp", "act"])
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    from humanoid import _native as nat
    from parity_utils import make_env
    dev = torch.device("cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(48 * 1024 * 1024, device=dev)
    if a.what == "env":
        env = make_env(a.num_envs, physics="synthetic")
        env.episode_length_buf = torch.randint(0, 2400, (a.num_envs,), device=dev)
        for _ in range(3):
            env.step(torch.randn(a.num_envs, 12, device=dev))
        ts = []
        for _ in range(a.reps):
            flush.zero_()
            e0.record()
            env._launch_post_physics(nat.PHASE_STEP_ALL)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        b = 7985 * a.num_envs
        print(f"post_physics N={a.num_envs}: {min(ts):.2f} us best, {sum(ts) / len(ts):.2f} us mean -> "
              f"{b / (sum(ts) / len(ts)) * 1e-3:.1f} GB/s algorithmic ({b / 1e6:.1f} MB)")
        return
    from humanoid.algo import ActorCritic, PPO
    ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128]).cuda()
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, learning_rate=1e-5, schedule="adaptive", entropy_coef=0.001,
              gamma=0.994, lam=0.9, device="cuda:0")
    N, T = a.num_envs, 60
    alg.init_storage(N, T, [705], [219], [12])
    s = alg.storage
    s.observations.normal_(), s.privileged_observations.normal_(), s.actions.normal_(), s.mu.normal_()
    s.sigma.fill_(1.0), s.values.normal_(), s.returns.normal_(), s.advantages.normal_(), s.actions_log_prob.fill_(-17.0)
    if a.what == "act":
        obs = torch.randn(N, 708, device=dev)[:, :705]          # row pitch 708 / 220, as the env produces them
        cobs = torch.randn(N, 220, device=dev)[:, :219]
        for _ in range(3):
            s.step = 0
            alg.act(obs, cobs)
        ts = []
        for _ in range(a.reps):
            s.step = 0
            e0.record()
            alg.act(obs, cobs)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        fl = (1052672 + 795392) * N
        print(f"PPO.act N={N}: {min(ts):.1f} us best, {sum(ts) / len(ts):.1f} us mean -> {fl / (sum(ts) / len(ts)) * 1e-6:.2f} TFLOP/s")
        return
    B = N * T // 4
    perm = torch.randperm(N * T, device=dev)
    ts = []
    for r in range(a.reps + 1):
        mb = s.gather(perm[:B], split=alg.use_split_path())
        e0.record()
        alg.minibatch_step(mb)
        e1.record()
        torch.cuda.synchronize()
        if r:
            ts.append(e0.elapsed_time(e1))
    fl = (1052672 + 795392 + 1383424 + 1254400) * B
    print(f"minibatch step B={B}: {min(ts):.2f} ms best, {sum(ts) / len(ts):.2f} ms mean -> {fl / (sum(ts) / len(ts)) * 1e-9:.2f} TFLOP/s")


if __name__ == "__main__":
    main()
