"""Intra-kernel timeline of post_physics_kernel (hg_env_set_trace): mean / max time of each phase over the CTAs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from humanoid import _native as nat  # noqa: E402
from parity_utils import make_env  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = make_env(N, physics="synthetic")
env.episode_length_buf = torch.randint(0, 2400, (N,), device="cuda")
for _ in range(3):
    env.step(torch.randn(N, 12, device="cuda"))
grid = (N + 31) // 32
trace = torch.zeros(grid, 12, dtype=torch.int64, device="cuda")
flush = torch.empty(48 * 1024 * 1024, device="cuda")
flush.zero_()
torch.cuda.synchronize()
nat.lib.hg_env_set_trace(trace.data_ptr())
env._launch_post_physics(nat.PHASE_STEP_ALL)
torch.cuda.synchronize()
nat.lib.hg_env_set_trace(None)
t = trace.cpu().double()
t0 = t[:, 0].min()
names = ["start", "staging issued", "staged data landed", "rewards done", "aliasing barrier", "compute warps done", "history shift done (thread 0)",
         "all warps joined", "new frames stored", "write-backs done", "kernel end"]
print(f"N={N}: grid {grid}, span {(t[:, 10].max() - t0) / 1e3:.1f} us; CTA start spread {(t[:, 0].max() - t0) / 1e3:.1f} us")
for k in range(1, 11):
    d = (t[:, k] - t[:, k - 1]) / 1e3
    print(f"  {names[k]:32s} +{d.mean():6.2f} us (max {d.max():6.2f})   at {((t[:, k] - t0) / 1e3).mean():6.1f} us")
