import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from humanoid import _native as nat
from parity_utils import make_env
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = make_env(N, physics="synthetic")
env.episode_length_buf = torch.randint(0, 2400, (N,), device="cuda")
for _ in range(3):
    env.step(torch.randn(N, 12, device="cuda"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(48 * 1024 * 1024, device="cuda")
ALL = nat.PHASE_STEP_ALL
C, CB, T, R, RS, O, L = (nat.PHASE_COUNTERS, nat.PHASE_CALLBACK, nat.PHASE_TERMINATE, nat.PHASE_REWARD, nat.PHASE_RESET,
                         nat.PHASE_OBS, nat.PHASE_LAST)
env.step(torch.randn(N, 12, device="cuda"))
print("reset fraction of the last step:", float(env.reset_buf.float().mean()))
cases = {"all": ALL, "c+cb": C | CB, "c+cb+t": C | CB | T, "c+cb+t+r": C | CB | T | R, "c+cb+t+r+rs": C | CB | T | R | RS,
         "c+t+rs": C | T | RS, "c+last": C | L, "no_obs": ALL & ~nat.PHASE_OBS, "obs_only": nat.PHASE_OBS | nat.PHASE_LAST, "obs_only_nolast": nat.PHASE_OBS,
         "counters_only": nat.PHASE_COUNTERS, "reward_only": nat.PHASE_REWARD}
for name, ph in cases.items():
    ts = []
    for _ in range(6):
        flush.zero_()
        e0.record()
        env._launch_post_physics(ph)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:16s} {min(ts):8.1f} us")
# plain copy of the same histories for reference
a = torch.randn(N, 736, device="cuda"); b = torch.empty_like(a)
c = torch.randn(N, 224, device="cuda"); d = torch.empty_like(c)
ts = []
for _ in range(6):
    flush.zero_()
    e0.record(); b.copy_(a); d.copy_(c); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"torch copy of (N,736)+(N,224): {min(ts):.1f} us -> {2 * 4 * N * 960 / min(ts) * 1e-3:.0f} GB/s")
