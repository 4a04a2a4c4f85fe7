"""Host-resident physics frames through the captured rollout: eager warm-up, capture, replays; prints ms / iteration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
kind = sys.argv[2] if len(sys.argv) > 2 else "synthetic_host"
dev = os.environ.get("HG_DEV", "cuda:0")
torch.cuda.set_device(dev)
env, runner = bench._make_runner(N, dev, kind)
state = (env.get_observations(), env.get_privileged_observations())
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.time()
    state, losses = bench._iterate(runner, state)
    r = float(env.rew_buf.mean().item())
    torch.cuda.synchronize()
    print(f"iter {i}: {1e3 * (time.time() - t0):.2f} ms  losses {losses}  mean reward {r:.5f}", flush=True)
