"""Two eager rollout steps (act, env.step, process_env_step) between cudaProfilerStart/Stop:
ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/profile_rollout_step.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env, runner = bench._make_runner(N, "cuda:0", "synthetic")
alg = runner.alg
obs, cobs = env.get_observations(), env.get_privileged_observations()
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
