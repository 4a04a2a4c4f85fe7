"""Intra-kernel timeline of hg_actor_critic_forward (mlp_chain_kernel): per work item the %globaltimer stamps of
  0 producer reaches the item   1 dependency satisfied   2 last TMA of the item issued
  3 MMA warp starts the item    4 last MMA committed
  5 epilogue sees the accumulator   6 tile published
Prints per-layer averages and the timeline of a few CTAs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from humanoid import _native as nat  # noqa: E402
from humanoid.algo import ActorCritic  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128]).cuda()
ac.flat_params()
obs = torch.randn(M, 736, device=dev)[:, :705]
cobs = torch.randn(M, 224, device=dev)[:, :219]
mu, val = torch.empty(M, 12, device=dev), torch.empty(M, 1, device=dev)
act, lp, sg = torch.empty(M, 12, device=dev), torch.empty(M, device=dev), torch.empty(M, 12, device=dev)
sample = dict(std=ac.std, actions=act, log_prob=lp, sigma=sg, seed=1, step=0)
ac.refresh_lo()
for _ in range(3):
    ac.native_act(obs, cobs, mu, val, sample)
torch.cuda.synchronize()
trace = torch.zeros(148, 16, 16, dtype=torch.int64, device=dev)
nat.lib.hg_actor_critic_set_trace(trace.data_ptr())
ac.native_act(obs, cobs, mu, val, sample)
torch.cuda.synchronize()
nat.lib.hg_actor_critic_set_trace(None)
t = trace.cpu()
valid = t[:, :, 0] > 0
t0 = int(t[:, :, 0][valid].min())
names = ["A1", "C1", "A2", "C2", "A3", "C3", "A4", "C4"]
print(f"kernel span: {(int(t[:, :, 6].max()) - t0) / 1e3:.1f} us")
print("layer  items   begin(us)  dep-wait   load-issue  mma-start->done   acc->published   published(us)  [means; min/max of published]")
for l in range(8):
    sel = valid & ((t[:, :, 7] >> 32) == l)
    if not sel.any():
        continue
    r = t[sel].double()
    rel = lambda k: (r[:, k] - t0) / 1e3
    print(f"{names[l]:5s} {int(sel.sum()):5d}  {rel(0).mean():9.1f}  {((r[:, 1] - r[:, 0]) / 1e3).mean():8.1f}  {((r[:, 2] - r[:, 1]) / 1e3).mean():10.1f}"
          f"  {((r[:, 4] - r[:, 3]) / 1e3).mean():15.1f}  {((r[:, 6] - r[:, 5]) / 1e3).mean():15.1f}  {rel(6).mean():13.1f}   [{rel(6).min():.1f} / {rel(6).max():.1f}]")
print("epilogue detail (warp 0 lane 0, first 32-column chunk): tmem ld+wait / math / rest of the tile's chunks + stores / membar   [us, means]")
for l in range(8):
    sel = valid & ((t[:, :, 7] >> 32) == l)
    if not sel.any():
        continue
    r = t[sel].double()
    d = lambda a, b: ((r[:, b] - r[:, a]) / 1e3).mean()
    print(f"{names[l]:5s} acc->ld issue {d(5, 8):6.2f}  ld {d(8, 9):6.2f}  math {d(9, 10):6.2f}  remaining chunks+stores {d(10, 11):6.2f}  membar {d(11, 12):6.2f}  barrier+atomic {d(12, 6):6.2f}")
for c in (0, 100):
    print(f"-- CTA {c}")
    for i in range(16):
        if t[c, i, 0] == 0:
            continue
        l, w = int(t[c, i, 7]) >> 32, int(t[c, i, 7]) & 0xFFFFFFFF
        s = [(int(t[c, i, k]) - t0) / 1e3 for k in range(7)]
        print(f"   {names[l]} w={w:4d}: begin {s[0]:6.1f} dep {s[1]:6.1f} issued {s[2]:6.1f} | mma {s[3]:6.1f} .. {s[4]:6.1f} | acc {s[5]:6.1f} pub {s[6]:6.1f}")
