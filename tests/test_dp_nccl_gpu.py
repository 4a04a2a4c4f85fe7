"""Product-path multi-rank parity: the CUDA path at 2 ranks x B/2 samples (NCCL, one logical
all-reduce of the flat gradient buffer per optimizer step, in two overlapped pieces) against the same CUDA path at
1 rank x B samples, on identical injected inputs: gradient rel-L2 < 1e-4 per tensor, identical adaptive-lr decision,
identical post-Adam weights (within fp32 reduction-order noise).  Needs >= 2 GPUs: `gpurun --gpus 2`."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(B, gen):
    obs = torch.randn(B, 705, generator=gen).clamp(-18, 18)
    cobs = torch.randn(B, 219, generator=gen).clamp(-18, 18)
    acts = torch.randn(B, 12, generator=gen)
    mu_old = 0.3 * torch.randn(B, 12, generator=gen)
    sg_old = 0.7 + 0.3 * torch.rand(B, 12, generator=gen)
    old_lp = (-((acts - mu_old) ** 2) / (2 * sg_old ** 2) - sg_old.log() - 0.9189385).sum(1, keepdim=True)
    tv = torch.randn(B, 1, generator=gen)
    ret = tv + 0.5 * torch.randn(B, 1, generator=gen)
    adv = torch.randn(B, 1, generator=gen)
    return dict(obs=obs, priv_obs=cobs, actions=acts, values=tv, advantages=adv, returns=ret, old_log_prob=old_lp, old_mu=mu_old,
                old_sigma=sg_old)


def _to_dev(mb, alg, dev):
    from humanoid import _native as nat
    out = {k: v.to(dev).contiguous() for k, v in mb.items()}
    for k in ("obs", "priv_obs"):                          # TMA-addressable row pitch for the fp32 engines
        x = out[k]
        buf = torch.zeros(x.shape[0], (x.shape[1] + 3) // 4 * 4, device=dev)
        buf[:, :x.shape[1]] = x
        out[k] = buf[:, :x.shape[1]]
    if alg.use_split_path():
        for src, dst in (("obs", "obs_split"), ("priv_obs", "priv_split")):
            x = out[src]
            planes = torch.zeros(2, x.shape[0], (x.shape[1] + 7) // 8 * 8, dtype=torch.int16, device=dev)
            nat.check(nat.lib.hg_split_bf16(x.data_ptr(), x.stride(0), nat.Split.of(planes), x.shape[0], x.shape[1],
                                            torch.cuda.current_stream(dev).cuda_stream))
            out[dst] = planes
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from humanoid.algo import ActorCritic, PPO

    def make():
        torch.manual_seed(7)
        ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128]).to(dev)
        with torch.no_grad():
            ac.std.copy_(0.5 + torch.rand(12, generator=torch.Generator().manual_seed(1)).to(dev))
        return ac, PPO(ac, num_learning_epochs=1, num_mini_batches=1, learning_rate=1e-5, schedule="adaptive", entropy_coef=0.001,
                       gamma=0.994, lam=0.9, device=str(dev))

    B = 4096
    full = _batch(B, torch.Generator().manual_seed(5))
    half = {k: v[rank * (B // world):(rank + 1) * (B // world)] for k, v in full.items()}
    ac, alg = make()
    alg.minibatch_step(_to_dev(half, alg, dev), world=world)
    torch.cuda.synchronize(dev)
    g_dp = alg._grad.clone()
    w_dp = ac.flat_params().clone()
    lr_dp = alg.learning_rate
    # every rank holds the same reduced gradient and takes the same decision
    ref = [torch.empty_like(g_dp) for _ in range(world)]
    dist.all_gather(ref, g_dp)
    same = all(torch.equal(ref[0], r) for r in ref)
    if rank == 0:
        ac1, alg1 = make()
        alg1.minibatch_step(_to_dev(full, alg1, dev), world=1)
        torch.cuda.synchronize(dev)
        rels = {}
        for name, _ in ac.named_parameters():
            a, b = ac.view_of(g_dp, name).double(), ac1.view_of(alg1._grad, name).double()
            rels[name] = float((a - b).norm() / b.norm().clamp_min(1e-30))
        n = ac.num_params
        q.put(dict(same=same, rels=rels, lr=(lr_dp, alg1.learning_rate),
                   stats=(g_dp[n:n + 4].tolist(), alg1._grad[n:n + 4].tolist()),
                   w_max=float((w_dp - ac1.flat_params()).abs().max()),
                   seeds=(alg._seed, None)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one_rank_on_the_product_path():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out["same"], "ranks disagree on the reduced gradient"
    worst = max(out["rels"].items(), key=lambda kv: kv[1])
    assert worst[1] < 1e-4, worst
    assert out["lr"][0] == out["lr"][1], out["lr"]
    for a, b in zip(*out["stats"]):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), out["stats"]
    assert out["w_max"] < 1.1e-5, out["w_max"]          # Adam's first step is ~lr * sign(g): tiny-gradient elements may differ by one step
