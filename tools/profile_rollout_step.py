"""Where does a rollout step go?

  (1) under ncu:  ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/profile_rollout_step.py ncu
      -> two eager steps between cudaProfilerStart/Stop (per-kernel durations)
  (2) stand-alone: python tools/profile_rollout_step.py
      -> CUDA-graph replays of the step's pieces, 60 repetitions each, timed with CUDA events:
         act only / env.step only / process_env_step only / the whole step / the runner's own rollout graph
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

bench._use_product()
mode = sys.argv[1] if len(sys.argv) > 1 else "time"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
env, runner = bench._make_runner(N, "cuda:0", "synthetic")
alg = runner.alg
obs, cobs = env.get_observations(), env.get_privileged_observations()

if mode == "ncu":
    with torch.inference_mode():
        for t in range(6):
            if t == 4:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
            a = alg.act(obs, cobs)
            obs, cobs, r, d, info = env.step(a)
            alg.process_env_step(r, d, info)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sys.exit(0)

T = 60
env.use_device_counters(True)
step_dev = env.noise_step_dev_ptr
state = {"obs": obs, "cobs": cobs}


def piece_act():
    alg.storage.step = 0
    for t in range(T):
        alg.storage.step = t
        alg.act(state["obs"], state["cobs"], step_dev=step_dev)
    alg._join_critic(torch.cuda.current_stream())


def piece_env():
    a = alg.storage.actions[0]
    for t in range(T):
        o, c, r, d, info = env.step(a)
        state["obs"], state["cobs"] = o, c


def piece_process():
    for t in range(T):
        alg.storage.step = t
        alg.process_env_step(env.rew_buf, env.reset_buf, env.extras)


def piece_step():
    o, c = state["obs"], state["cobs"]
    alg.storage.step = 0
    for t in range(T):
        a = alg.act(o, c, step_dev=step_dev)
        o, c, r, d, info = env.step(a)
        alg.process_env_step(r, d, info)
    state["obs"], state["cobs"] = o, c


def timed_graph(fn, name):
    with torch.inference_mode():
        fn()                                   # eager warm-up
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps / T
        print(f"{name:28s} {us:8.1f} us per step")


timed_graph(piece_act, "act (chain + obs copy)")
timed_graph(piece_env, "env.step")
timed_graph(piece_process, "process_env_step")
timed_graph(piece_step, "whole step")
os.environ["HG_FUSED_ACT"] = "0"
timed_graph(piece_act, "act, per-layer launches")
os.environ["HG_FUSED_ACT"] = "1"
os.environ["HG_FUSED_DECIMATION"] = "0"
env.gym.fused = False
timed_graph(piece_env, "env.step, decimation loop")
