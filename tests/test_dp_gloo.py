"""Env-sharded data parallelism on CPU with the gloo backend, world_size = 2.

The product's PPO.minibatch_step makes every rank compute partial sums scaled by 1/B_global and SUM-all-reduces
ONE flat buffer [gradient | surrogate, value loss, entropy, KL]; these tests check, with the oracle as the compute
stand-in, that this protocol reproduces the single-process gradient / losses / learning-rate decision exactly, and
that the optional 3-number all-reduce gives the global advantage normalisation."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ppo_oracle as po  # noqa: E402

NA, NC, A = 60, 40, 12
HID_A, HID_C = (32, 24, 16), (48, 24, 16)


def _batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    p = po.init_params(NA, NC, A, HID_A, HID_C, 1.0, generator=g)
    p["std"] = 0.5 + torch.rand(A, generator=g)
    obs, cobs = torch.randn(B, NA, generator=g), torch.randn(B, NC, generator=g)
    with torch.no_grad():
        mu, sg = po.actor_dist(obs, p)
        mu_old = mu + 0.05 * torch.randn(B, A, generator=g)
        sg_old = sg * (1 + 0.05 * torch.randn(B, A, generator=g)).clamp(0.8, 1.2)
        acts = mu_old + sg_old * torch.randn(B, A, generator=g)
        old_lp = po.log_prob(acts, mu_old, sg_old).unsqueeze(1)
        val = po.mlp(cobs, p, "critic")
    tv = val + 0.3 * torch.randn(B, 1, generator=g)
    ret = val + 0.5 * torch.randn(B, 1, generator=g)
    adv = torch.randn(B, 1, generator=g)
    return p, (obs, cobs, acts, tv, adv, ret, old_lp, mu_old, sg_old)


def _sharded_loss(p, batch, B_global, ecoef=0.001):
    """What one rank's kernels compute: every mean uses 1/B_global; entropy enters with weight B_local/B_global."""
    B_local = batch[0].shape[0]
    loss, sur, vl, kl = po.ppo_loss(p, batch, entropy_coef=0.0)
    frac = B_local / B_global
    ent = po.entropy(p["std"].expand(B_local, -1)).mean()
    loss = loss * frac - ecoef * ent * frac
    return loss, torch.stack([sur.detach() * frac, vl.detach() * frac, ent.detach() * frac, kl * frac])


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    p, batch = _batch(B, seed=3)
    lo, hi = rank * B // world, (rank + 1) * B // world
    shard = tuple(t[lo:hi] for t in batch)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss, scalars = _sharded_loss(q, shard, B)
    loss.backward()
    flat = torch.cat([q[k].grad.reshape(-1) for k in q] + [scalars])
    dist.all_reduce(flat)                                      # the ONE collective of an optimizer step
    n = flat.numel() - 4
    lr = po.adapt_lr(1e-5, flat[n + 3])
    # advantage normalisation from all-reduced (sum, sum of squares, count)
    adv = torch.randn(64, generator=torch.Generator().manual_seed(100 + rank)).double()
    st = torch.tensor([adv.sum(), (adv * adv).sum(), float(adv.numel())], dtype=torch.float64)
    dist.all_reduce(st)
    mean = st[0] / st[2]
    std = ((st[1] - st[0] * mean) / (st[2] - 1)).sqrt()
    out[rank] = (flat.clone(), lr, ((adv - mean) / (std + 1e-8)).float())
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_gradient_equals_single_process():
    world, B = 2, 256
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, B, out), nprocs=world, join=True)
    p, batch = _batch(B, seed=3)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss, sur, vl, kl = po.ppo_loss(q, batch)
    loss.backward()
    ref = torch.cat([q[k].grad.reshape(-1) for k in q])
    f0, lr0, a0 = out[0]
    f1, lr1, a1 = out[1]
    assert torch.equal(f0, f1) and lr0 == lr1, "ranks must hold identical reduced buffers and lr decisions"
    n = ref.numel()
    rel = float((f0[:n].double() - ref.double()).norm() / ref.double().norm())
    assert rel < 1e-5, rel
    assert abs(float(f0[n]) - float(sur)) < 1e-6 and abs(float(f0[n + 1]) - float(vl)) < 1e-5 and abs(float(f0[n + 3]) - float(kl)) < 1e-6
    assert lr0 == po.adapt_lr(1e-5, kl)
    # global advantage normalisation == normalising the concatenation
    full = torch.cat([torch.randn(64, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)])
    want = (full - full.mean()) / (full.std() + 1e-8)
    got = torch.cat([a0, a1])
    assert torch.allclose(got, want, atol=1e-6)


def test_policy_sample_seed_differs_per_rank():
    """the Philox key of hg_policy_sample must include the rank, or every env shard draws the same
    exploration noise.  (Host-side check of the seed derivation; the draw itself is keyed (seed, env, step).)"""
    import importlib.util
    import os
    import re
    src = open(os.path.join(ROOT, "humanoid-gym_b200", "humanoid", "algo", "ppo", "ppo.py")).read()
    body = re.search(r"    def rank_seed\(seed, rank\):\n((?:        .*\n)+)", src).group(1)
    ns = {}
    exec("def rank_seed(seed, rank):\n" + body.replace("        ", "    ", 1).replace("\n        ", "\n    "), ns)
    seeds = {ns["rank_seed"](123456789, r) for r in range(8)}
    assert len(seeds) == 8
    assert all(0 <= s < (1 << 64) for s in seeds)
