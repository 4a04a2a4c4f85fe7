"""GPU parity of the learning-side kernels (through the C ABI / the drop-in algo classes) against the
torch-fp32 oracle and the reference-generated goldens.  Tolerances: 1e-5 relative on forward values /
returns, 1e-4 relative L2 on gradients (north_star)."""
import numpy as np
import pytest
import torch

from golden_io import Golden
from oracle import ppo_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["simt_fp32", "tcgen05_3xtf32", "tcgen05_bf16x3"])
def gemm_engine(request):
    """Every test runs on all three GEMM engines: the exact-fp32 CUDA-core path, the tcgen05 3xTF32 path, and the
    default: the split-precision bf16x3 path for PPO.update (3xTF32 for the rollout forward)."""
    from humanoid import _native as nat
    prev = nat.lib.hg_set_gemm_mode({"simt_fp32": 0, "tcgen05_3xtf32": 1, "tcgen05_bf16x3": 4}[request.param])
    yield request.param
    nat.lib.hg_set_gemm_mode(prev)


def _pad4(t):
    """Same values, row pitch rounded up to 4 floats (what RolloutStorage.gather produces)."""
    R, Cc = t.shape
    buf = torch.zeros(R, (Cc + 3) // 4 * 4, device=t.device)
    buf[:, :Cc] = t
    return buf[:, :Cc]


def _with_split(alg, mb):
    """Minibatch dicts built by hand carry fp32 observations; on the bf16x3 engine add their split images."""
    from humanoid import _native as nat
    if not alg.use_split_path():
        return mb
    for src, dst in (("obs", "obs_split"), ("priv_obs", "priv_split")):
        x = mb[src]
        R, Cc = x.shape
        planes = torch.zeros(2, R, (Cc + 7) // 8 * 8, dtype=torch.int16, device="cuda")
        nat.check(nat.lib.hg_split_bf16(x.data_ptr(), x.stride(0), nat.Split.of(planes), R, Cc, 0), "hg_split_bf16")
        mb[dst] = planes
    return mb


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _cat(ac, which="data"):
    """Unpadded concatenation of parameters / gradients in named_parameters() order."""
    return torch.cat([(p.data if which == "data" else p.grad).reshape(-1) for p in ac.parameters()])


def _make_ac(na, nc, act, ah, ch, params=None):
    from humanoid.algo import ActorCritic
    ac = ActorCritic(na, nc, act, actor_hidden_dims=ah, critic_hidden_dims=ch).cuda()
    if params is not None:
        ac.load_state_dict({k: v.cuda() for k, v in params.items()})
    ac.flat_params()
    return ac


def test_policy_example_known_answers():
    """The reference's shipped actor (705-512-256-128-12) through hg_mlp_forward."""
    k = Golden("policy_example_kat.npz")
    w = k.group("w.")
    ac = _make_ac(705, 219, 12, [512, 256, 128], [768, 256, 128])
    sd = ac.state_dict()
    for n, v in w.items():
        sd["actor." + n] = v.cuda()
    ac.load_state_dict(sd)
    y = ac.act_inference(k.t("x").cuda())
    # 705-term fp32 dot products: compare norm-wise at 1e-5 and element-wise with an absolute floor of 1e-5 * max|y|
    assert _rel(y.cpu(), k.t("y")) < 1e-5, _rel(y.cpu(), k.t("y"))
    np.testing.assert_allclose(y.cpu().numpy(), k["y"], rtol=1e-5, atol=1e-5 * float(np.abs(k["y"]).max()))
    kat = [0.0847, -0.0234, 0.0057, 0.2348, 0.6382, -0.2275, -0.1129, -0.1501, 0.2042, 0.3535, 0.0077, -0.4530]
    np.testing.assert_allclose(y[0].cpu().numpy(), kat, atol=5e-5)


def test_forward_sample_logprob_vs_golden():
    g = Golden("ppo_learning.npz")
    p0 = g.group("w0.")
    ac = _make_ac(60, 40, 12, [32, 24, 16], [48, 24, 16], p0)
    obs, cobs, acts = g.t("fwd.obs").cuda(), g.t("fwd.cobs").cuda(), g.t("fwd.actions").cuda()
    ac.update_distribution(obs)
    np.testing.assert_allclose(ac.action_mean.cpu().numpy(), g["fwd.mean"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ac.get_actions_log_prob(acts).cpu().numpy(), g["fwd.logp"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ac.entropy.cpu().numpy(), g["fwd.entropy"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ac.evaluate(cobs).cpu().numpy(), g["fwd.value"], rtol=1e-5, atol=1e-6)
    # fused sample + log-prob kernel with injected eps vs the oracle
    from humanoid import _native as nat
    M = obs.shape[0]
    eps = torch.randn(M, 12, generator=torch.Generator().manual_seed(1))
    mean = ac.action_mean.contiguous()
    a, lp, sg = torch.empty(M, 12, device="cuda"), torch.empty(M, device="cuda"), torch.empty(M, 12, device="cuda")
    nat.check(nat.lib.hg_policy_sample(mean.data_ptr(), ac.std.data_ptr(), eps.cuda().data_ptr(), 0, 0, None, a.data_ptr(),
                                       lp.data_ptr(), sg.data_ptr(), M, 12, 0))
    ra, rv, rlp, rmu, rsg = po.act(obs.cpu(), cobs.cpu(), p0, eps)
    np.testing.assert_allclose(a.cpu().numpy(), ra.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lp.cpu().numpy(), rlp.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sg.cpu().numpy(), rsg.numpy(), rtol=0, atol=0)
    # Philox path: standard-normal statistics
    M2 = 1 << 16
    mean0 = torch.zeros(M2, 12, device="cuda")
    one = torch.ones(12, device="cuda")
    a2, lp2, sg2 = torch.empty(M2, 12, device="cuda"), torch.empty(M2, device="cuda"), torch.empty(M2, 12, device="cuda")
    nat.check(nat.lib.hg_policy_sample(mean0.data_ptr(), one.data_ptr(), None, 1234, 7, None, a2.data_ptr(), lp2.data_ptr(),
                                       sg2.data_ptr(), M2, 12, 0))
    assert abs(float(a2.mean())) < 0.01 and abs(float(a2.std()) - 1) < 0.01
    assert abs(float((a2 ** 4).mean()) - 3) < 0.1


def test_act_fused_sampling_matches_standalone_kernel():
    """PPO.act samples inside the actor's output-layer epilogue on the tensor-core engines; the stand-alone
    hg_policy_sample kernel on the same mean must give bit-identical actions / log-prob / sigma (Philox and injected eps)."""
    from humanoid.algo import PPO
    from humanoid import _native as nat
    ac = _make_ac(705, 219, 12, [512, 256, 128], [768, 256, 128])
    with torch.no_grad():
        ac.std.copy_(0.5 + torch.rand(12, device="cuda"))
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=1, device="cuda:0")
    N = 1000
    alg.init_storage(N, 4, [705], [219], [12])
    s = alg.storage
    g = torch.Generator(device="cuda").manual_seed(2)
    for t, eps in enumerate((None, torch.randn(N, 12, device="cuda", generator=g))):
        obs = _pad4(torch.randn(N, 705, device="cuda", generator=g))
        cobs = _pad4(torch.randn(N, 219, device="cuda", generator=g))
        alg.act(obs, cobs, eps=eps, step=77 + t)
        torch.cuda.synchronize()
        mu = s.mu[t].contiguous()
        a, lp, sg = torch.empty(N, 12, device="cuda"), torch.empty(N, device="cuda"), torch.empty(N, 12, device="cuda")
        nat.check(nat.lib.hg_policy_sample(mu.data_ptr(), ac.std.data_ptr(), nat.ptr(eps), alg._seed, 77 + t, None, a.data_ptr(),
                                           lp.data_ptr(), sg.data_ptr(), N, 12, 0))
        torch.cuda.synchronize()
        assert torch.equal(a, s.actions[t]) and torch.equal(sg, s.sigma[t]), t
        assert torch.equal(lp, s.actions_log_prob[t].view(-1)), float((lp - s.actions_log_prob[t].view(-1)).abs().max())
        ref_mu = po.mlp(obs.cpu(), {k: v.detach().cpu() for k, v in ac.state_dict().items()}, "actor")
        assert _rel(mu.cpu(), ref_mu) < 1e-5
        s.step += 1


@pytest.mark.parametrize("chain", ["f16x3", "3xtf32"])
@pytest.mark.parametrize("N", [4096, 1000, 16384])
def test_fused_act_matches_separate_chains(N, gemm_engine, chain, monkeypatch):
    """PPO.act as one persistent launch (both nets, all layers, on-device layer dependencies) against the per-layer launches,
    over several calls (the tile counters must come back to zero) and with a ragged last row tile.
    3xtf32 chain (hg_actor_critic_forward): same tile shapes and MMA order as the per-layer kernel -> bit-identical mean / value /
    actions / log-prob.  f16x3 chain (hg_actor_critic_forward_f16, the default): a different operand format -> mean / value within
    1e-5 of the oracle (measured 3.8e-6 against fp64, the 3xTF32 engines 8e-6: tools/act_accuracy.py -- the error of both is the
    truncating fp32 accumulation of the tensor core, proportional to the number of MMAs per accumulator, and fp16x3 issues half as
    many) and so within 1.5e-5 of the per-layer 3xTF32 results; sigma bit-identical; actions = mean + sigma z on the same z."""
    import os
    from humanoid.algo import PPO
    if gemm_engine == "simt_fp32":
        pytest.skip("the fused kernel is a tensor-core path")
    monkeypatch.setenv("HG_CHAIN_F16", "1" if chain == "f16x3" else "0")
    ac = _make_ac(705, 219, 12, [512, 256, 128], [768, 256, 128])
    with torch.no_grad():
        ac.std.copy_(0.5 + torch.rand(12, device="cuda"))
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=1, device="cuda:0")
    alg.init_storage(N, 6, [705], [219], [12])
    s = alg.storage
    g = torch.Generator(device="cuda").manual_seed(N)
    obs = [_pad4(torch.randn(N, 705, device="cuda", generator=g)) for _ in range(3)]
    cobs = [_pad4(torch.randn(N, 219, device="cuda", generator=g)) for _ in range(3)]
    for t in range(3):                                   # fused (default)
        alg.act(obs[t], cobs[t], step=100 + t)
        s.step += 1
    torch.cuda.synchronize()
    for key, buf in ac._scratch.items():
        if key[0] == "chain_counters":
            assert buf.abs().sum() == 0, f"tile counters {key} were not re-zeroed"
    os.environ["HG_FUSED_ACT"] = "0"
    try:
        for t in range(3):
            alg.act(obs[t], cobs[t], step=100 + t)
            s.step += 1
    finally:
        os.environ["HG_FUSED_ACT"] = "1"
    torch.cuda.synchronize()
    for t in range(3):
        for k in ("mu", "values", "actions", "actions_log_prob", "sigma"):
            a, b = getattr(s, k)[t], getattr(s, k)[t + 3]
            if chain == "3xtf32" or k == "sigma":
                assert torch.equal(a, b), (t, k, float((a - b).abs().max()))
            elif k in ("mu", "values"):
                assert _rel(a, b) < 1.5e-5, (t, k, _rel(a, b))
            elif k == "actions":                          # same noise: the action differs by exactly the difference of the means
                assert float(((a - s.mu[t]) - (b - s.mu[t + 3])).abs().max()) < 2e-6, (t, k)
            else:
                assert float((a - b).abs().max()) < 1e-4, (t, k, float((a - b).abs().max()))
    p = {k: v.detach().cpu() for k, v in ac.state_dict().items()}
    assert _rel(s.mu[0].cpu(), po.mlp(obs[0].cpu(), p, "actor")) < 1e-5
    assert _rel(s.values[0].cpu(), po.mlp(cobs[0].cpu(), p, "critic")) < 1e-5


@pytest.fixture(params=["warp_scan", "serial"])
def gae_mode(request):
    from humanoid import _native as nat
    prev = nat.lib.hg_set_gae_mode(1 if request.param == "warp_scan" else 0)
    yield request.param
    nat.lib.hg_set_gae_mode(prev)


@pytest.mark.parametrize("T,N", [(1, 33), (31, 64), (32, 1), (33, 100), (60, 1000), (200, 96)])
def test_gae_shapes_vs_oracle(gae_mode, T, N):
    """Ragged env counts and rollout lengths on both sides of the 32-lane width (chunks of 1, 2, 7 steps per lane)."""
    from humanoid.algo import RolloutStorage
    gen = torch.Generator().manual_seed(T * 1000 + N)
    st = RolloutStorage(N, T, [4], [4], [12], "cuda:0")
    r, v = torch.rand(T, N, 1, generator=gen), torch.randn(T, N, 1, generator=gen)
    d = (torch.rand(T, N, 1, generator=gen) < 0.1).byte()
    lv = torch.randn(N, 1, generator=gen)
    st.rewards.copy_(r), st.values.copy_(v), st.dones.copy_(d)
    st.compute_returns(lv.cuda(), 0.994, 0.9)
    ret, adv = po.gae(r, v, d, lv, 0.994, 0.9)
    np.testing.assert_allclose(st.returns.cpu().numpy(), ret.numpy(), rtol=1e-5, atol=1e-5)
    if T * N > 1:
        np.testing.assert_allclose(st.advantages.cpu().numpy(), adv.numpy(), rtol=1e-5, atol=2e-5)


def test_gae_vs_golden_and_large(gae_mode):
    from humanoid.algo import RolloutStorage
    g = Golden("ppo_learning.npz")
    T, N = g["gae.rewards"].shape[:2]
    st = RolloutStorage(N, T, [4], [4], [12], "cuda:0")
    st.rewards.copy_(g.t("gae.rewards")), st.values.copy_(g.t("gae.values")), st.dones.copy_(g.t("gae.dones"))
    st.compute_returns(g.t("gae.last_values").cuda(), 0.994, 0.9)
    np.testing.assert_allclose(st.returns.cpu().numpy(), g["gae.returns"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.advantages.cpu().numpy(), g["gae.advantages"], rtol=1e-5, atol=2e-6)
    # benchmark size vs the oracle
    T, N = 60, 4096
    gen = torch.Generator().manual_seed(0)
    st = RolloutStorage(N, T, [4], [4], [12], "cuda:0")
    r, v = torch.rand(T, N, 1, generator=gen), torch.randn(T, N, 1, generator=gen)
    d = (torch.rand(T, N, 1, generator=gen) < 0.02).byte()
    lv = torch.randn(N, 1, generator=gen)
    st.rewards.copy_(r), st.values.copy_(v), st.dones.copy_(d)
    st.compute_returns(lv.cuda(), 0.994, 0.9)
    ret, adv = po.gae(r, v, d, lv, 0.994, 0.9)
    np.testing.assert_allclose(st.returns.cpu().numpy(), ret.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(st.advantages.cpu().numpy(), adv.numpy(), rtol=1e-5, atol=1e-5)
    assert abs(float(st.advantages.mean())) < 1e-5 and abs(float(st.advantages.std()) - 1) < 1e-5


def _storage_from(gd, alg):
    s = alg.storage
    for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "returns", "advantages",
              "actions_log_prob", "mu", "sigma"):
        getattr(s, k).copy_(gd[k].cuda())


def test_full_update_vs_golden():
    """PPO.update(): 2 epochs x 4 minibatches with the reference's permutation; per-step gradients, adaptive
    learning rates, losses and final weights against what the unmodified reference produced."""
    from humanoid.algo import PPO
    g = Golden("ppo_learning.npz")
    ac = _make_ac(60, 40, 12, [32, 24, 16], [48, 24, 16], g.group("w0."))
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
              entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device="cuda:0")
    gd = g.group("upd.storage.")
    T, N = gd["rewards"].shape[:2]
    alg.init_storage(N, T, [60], [40], [12])
    _storage_from(gd, alg)
    perm = g.t("upd.perm").cuda()
    mini = (T * N) // 4
    ref_grads, ref_lrs = g["upd.grads"], g["upd.lrs"]
    n = ac.num_params
    alg._loss_sums.zero_()
    step = 0
    for _ in range(2):
        for i in range(4):
            mb = alg.storage.gather(perm[i * mini:(i + 1) * mini], split=alg.use_split_path())
            alg.minibatch_step(mb)
            torch.cuda.synchronize()
            grad = _cat(ac, "grad").cpu().numpy()
            rel = np.linalg.norm(grad - ref_grads[step]) / np.linalg.norm(ref_grads[step])
            assert rel < 1e-4, f"optimizer step {step}: gradient rel-L2 {rel:.3g}"
            assert abs(alg.learning_rate - ref_lrs[step]) <= 1e-12 * ref_lrs[step], (step, alg.learning_rate, ref_lrs[step])
            step += 1
    sums = alg._loss_sums.tolist()
    assert abs(sums[1] / 8 - float(g["upd.mean_value_loss"])) < 1e-5 * max(1.0, abs(float(g["upd.mean_value_loss"])))
    assert abs(sums[0] / 8 - float(g["upd.mean_surrogate_loss"])) < 1e-5
    for k, v in g.group("w1.").items():
        got = ac.state_dict()[k].cpu()
        assert torch.allclose(got, v, rtol=1e-4, atol=1e-7), f"{k}: rel {_rel(got, v):.3g}"
    # the public entry point runs too (fresh permutation) and returns finite means
    _storage_from(gd, alg)
    alg.storage.step = T
    vl, sl = alg.update()
    assert np.isfinite(vl) and np.isfinite(sl) and alg.storage.step == 0


@pytest.mark.parametrize("B", [4096, 61440])
def test_flagship_gradients_vs_autograd(B, gemm_engine):
    """Full XBot-L architecture, one minibatch (4096 samples, and 61,440 = the real minibatch of the 4096-env
    configuration): native loss+backward vs torch autograd (oracle)."""
    from humanoid.algo import PPO
    if B > 4096 and gemm_engine == "simt_fp32":
        pytest.skip("the exact-fp32 CUDA-core engine is covered at B=4096; 61,440 is for the tensor-core engines")
    torch.manual_seed(3)
    ac = _make_ac(705, 219, 12, [512, 256, 128], [768, 256, 128])
    with torch.no_grad():
        ac.std.copy_(0.5 + torch.rand(12, device="cuda"))
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=1, learning_rate=1e-5, schedule="adaptive", entropy_coef=0.001,
              gamma=0.994, lam=0.9, device="cuda:0")
    gen = torch.Generator().manual_seed(5)
    p = {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}
    obs = torch.randn(B, 705, generator=gen).clamp(-18, 18)
    cobs = torch.randn(B, 219, generator=gen).clamp(-18, 18)
    with torch.no_grad():
        mu_old, sg_old = po.actor_dist(obs, p)
        mu_old = mu_old + 0.05 * torch.randn(B, 12, generator=gen)
        sg_old = sg_old * (1 + 0.05 * torch.randn(B, 12, generator=gen)).clamp(0.8, 1.2)
        acts = mu_old + sg_old * torch.randn(B, 12, generator=gen)
        old_lp = po.log_prob(acts, mu_old, sg_old).unsqueeze(1)
        val = po.mlp(cobs, p, "critic")
    tv = val + 0.3 * torch.randn(B, 1, generator=gen)
    ret = val + 0.5 * torch.randn(B, 1, generator=gen)
    adv = torch.randn(B, 1, generator=gen)
    batch = (obs, cobs, acts, tv, adv, ret, old_lp, mu_old, sg_old)
    L = po.Learner(p, lr=1e-5)
    loss, sur, vl, kl = po.ppo_loss(L.p, batch)
    loss.backward()
    ref = L.flat_grad()
    mb = dict(obs=obs, priv_obs=cobs, actions=acts, values=tv, advantages=adv, returns=ret, old_log_prob=old_lp,
              old_mu=mu_old, old_sigma=sg_old)
    mb = {k: v.cuda().contiguous() for k, v in mb.items()}
    mb["obs"], mb["priv_obs"] = _pad4(mb["obs"]), _pad4(mb["priv_obs"])
    w_before = _cat(ac).clone()
    alg.minibatch_step(_with_split(alg, mb))
    torch.cuda.synchronize()
    got = _cat(ac, "grad").cpu()
    assert _rel(got, ref) < 1e-4, _rel(got, ref)
    off = 0
    for name in L.names:                                     # per-tensor check (worst tensor in SURVEY: critic.0.weight)
        k = L.p[name].numel()
        assert _rel(got[off:off + k], ref[off:off + k]) < 1e-4, (name, _rel(got[off:off + k], ref[off:off + k]))
        off += k
    sc = alg._scalars.cpu()
    assert abs(float(sc[0]) - float(sur)) < 1e-5 and abs(float(sc[1]) - float(vl)) < 1e-5 * max(1, float(vl))
    assert abs(float(sc[3]) - float(kl)) < 1e-5
    # clip + Adam against torch.optim.Adam on the oracle side
    L.lr = po.adapt_lr(L.lr, kl)
    for gp in L.opt.param_groups:
        gp["lr"] = L.lr
    torch.nn.utils.clip_grad_norm_([L.p[k] for k in L.names], 1.0)
    L.opt.step()
    want = torch.cat([L.p[k].detach().reshape(-1) for k in L.names])
    # the fp32 update (~1e-5) is only a few ulp of the weights it lands on, so compare the new parameters
    # tightly and the update itself within that quantisation
    # Adam's first step moves every weight by lr * g / (|g| + 1e-8): for the few elements whose gradient is itself
    # ~1e-8 an absolute gradient difference of 1e-9 (well inside the 1e-4 relative budget) changes the step by a
    # sizeable fraction of lr, for ANY fp32 summation order.  So: every weight within one lr-step (1e-5) of the
    # oracle's, 99.9 % of them within 2 % of a step, and the update vector as a whole within 1 %.
    diff = (_cat(ac).cpu() - want).abs()
    assert float(diff.max()) < 1.0e-5, float(diff.max())
    assert float((diff > 2e-7).float().mean()) < 1e-3, float((diff > 2e-7).float().mean())
    upd_got, upd_want = (_cat(ac).cpu() - w_before.cpu()), (want - w_before.cpu())
    assert _rel(upd_got, upd_want) < 1e-2, _rel(upd_got, upd_want)
    assert abs(alg.learning_rate - L.lr) < 1e-18


def test_ragged_small_net_gradients():
    """Odd sizes on purpose: batch not a multiple of any tile, widths that are not multiples of 4 (scalar bias
    sums, the generic GEMM fallback) next to ones that are (128-bit sums, skinny-head kernels)."""
    from humanoid.algo import PPO
    torch.manual_seed(11)
    na, nc, nact, B = 37, 29, 3, 777
    ac = _make_ac(na, nc, nact, [52, 20], [24, 8])
    with torch.no_grad():
        ac.std.copy_(0.5 + torch.rand(nact, device="cuda"))
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=1, learning_rate=1e-5, schedule="adaptive", entropy_coef=0.001,
              gamma=0.994, lam=0.9, device="cuda:0")
    gen = torch.Generator().manual_seed(12)
    p = {k: v.detach().cpu().clone() for k, v in ac.state_dict().items()}
    obs = torch.randn(B, na, generator=gen)
    cobs = torch.randn(B, nc, generator=gen)
    with torch.no_grad():
        mu_old, sg_old = po.actor_dist(obs, p)
        mu_old = mu_old + 0.05 * torch.randn(B, nact, generator=gen)
        acts = mu_old + sg_old * torch.randn(B, nact, generator=gen)
        old_lp = po.log_prob(acts, mu_old, sg_old).unsqueeze(1)
        val = po.mlp(cobs, p, "critic")
    tv = val + 0.3 * torch.randn(B, 1, generator=gen)
    ret = val + 0.5 * torch.randn(B, 1, generator=gen)
    adv = torch.randn(B, 1, generator=gen)
    L = po.Learner(p, lr=1e-5)
    loss, sur, vl, kl = po.ppo_loss(L.p, (obs, cobs, acts, tv, adv, ret, old_lp, mu_old, sg_old))
    loss.backward()
    ref = L.flat_grad()
    mb = dict(obs=obs, priv_obs=cobs, actions=acts, values=tv, advantages=adv, returns=ret, old_log_prob=old_lp,
              old_mu=mu_old, old_sigma=sg_old)
    mb = {k: v.cuda().contiguous() for k, v in mb.items()}
    mb["obs"], mb["priv_obs"] = _pad4(mb["obs"]), _pad4(mb["priv_obs"])
    alg.minibatch_step(_with_split(alg, mb))
    torch.cuda.synchronize()
    got = _cat(ac, "grad").cpu()
    off = 0
    for name in L.names:
        k = L.p[name].numel()
        assert _rel(got[off:off + k], ref[off:off + k]) < 1e-4, (name, _rel(got[off:off + k], ref[off:off + k]))
        off += k


def test_checkpoint_roundtrip(tmp_path):
    """model_<it>.pt keeps the reference's format: keys and optimizer state reload into fresh objects."""
    from humanoid.algo import PPO
    ac = _make_ac(60, 40, 12, [32, 24, 16], [48, 24, 16])
    alg = PPO(ac, learning_rate=3e-4, device="cuda:0")
    alg._exp_avg.normal_(), alg._exp_avg_sq.uniform_(), alg._adam_step.fill_(17)
    alg.sync_optimizer_container()
    path = tmp_path / "model_0.pt"
    torch.save({"model_state_dict": ac.state_dict(), "optimizer_state_dict": alg.optimizer.state_dict(), "iter": 3, "infos": None}, path)
    sd = torch.load(path)
    assert list(sd["model_state_dict"].keys()) == po.param_names()
    ac2 = _make_ac(60, 40, 12, [32, 24, 16], [48, 24, 16])
    alg2 = PPO(ac2, learning_rate=1e-3, device="cuda:0")
    ac2.load_state_dict(sd["model_state_dict"])
    alg2.optimizer.load_state_dict(sd["optimizer_state_dict"])
    alg2.load_optimizer_container()
    assert torch.equal(_cat(ac2), _cat(ac))
    for name, _ in ac.named_parameters():
        assert torch.equal(ac2.view_of(alg2._exp_avg, name), ac.view_of(alg._exp_avg, name)), name
        assert torch.equal(ac2.view_of(alg2._exp_avg_sq, name), ac.view_of(alg._exp_avg_sq, name)), name
    assert int(alg2._adam_step) == 17 and abs(alg2.learning_rate - 3e-4) < 1e-12
    # the reference's own ActorCritic accepts the weights (same module tree): checked structurally
    assert ac2.actor[0].weight.shape == (32, 60) and ac2.critic[6].weight.shape == (1, 16)


def test_runner_end_to_end_small():
    """task_registry.make_env -> make_alg_runner -> learn(2) on the synthetic physics source."""
    from parity_utils import make_args
    from humanoid.envs import XBotLCfg  # noqa: F401
    from humanoid.utils import task_registry
    args = make_args(256)
    args.max_iterations = 2
    env, _ = task_registry.make_env("humanoid_ppo", args=args)
    runner, cfg = task_registry.make_alg_runner(env, name="humanoid_ppo", args=args, log_root=None)
    w0 = runner.alg.actor_critic.flat_params().clone()
    runner.learn(2, init_at_random_ep_len=True)
    assert runner.current_learning_iteration == 2
    assert torch.isfinite(runner.alg.actor_critic.flat_params()).all()
    assert not torch.equal(w0, runner.alg.actor_critic.flat_params())
    assert runner.last_perf["fps"] > 0


def test_graph_replay_matches_eager_rollout():
    """The CUDA-graph rollout (default) must reproduce the eager rollout bit for bit: same storage slabs, same final
    observations / env state, for two consecutive collection phases (capture happens on the second)."""
    from parity_utils import make_args
    from humanoid.envs import XBotLCfg  # noqa: F401
    from humanoid.utils import task_registry
    import os

    def build():
        torch.manual_seed(0)
        np.random.seed(0)
        args = make_args(256)
        env, _ = task_registry.make_env("humanoid_ppo", args=args)
        runner, _ = task_registry.make_alg_runner(env, name="humanoid_ppo", args=args, log_root=None)
        return env, runner

    def run(runner, env, phases):
        obs, cobs = env.get_observations(), env.get_privileged_observations()
        out = []
        with torch.inference_mode():
            for _ in range(phases):
                obs, cobs = runner.collect(obs, cobs)
                torch.cuda.synchronize()
                s = runner.alg.storage
                out.append({k: getattr(s, k).clone() for k in ("observations", "privileged_observations", "actions", "rewards", "dones",
                                                               "values", "actions_log_prob", "mu", "sigma", "returns", "advantages")})
                out[-1]["obs"], out[-1]["rew"] = obs.clone(), env.rew_buf.clone()
                s.clear()
        return out

    os.environ["HG_CUDA_GRAPH"] = "0"
    try:
        env_e, run_e = build()
        eager = run(run_e, env_e, 3)
    finally:
        os.environ["HG_CUDA_GRAPH"] = "1"
    env_g, run_g = build()
    graph = run(run_g, env_g, 3)
    assert getattr(run_g, "_graph", None) is not None, "the graph path did not engage"
    for ph, (a, b) in enumerate(zip(eager, graph)):
        for k in a:
            assert torch.equal(a[k], b[k]), (ph, k, float((a[k].float() - b[k].float()).abs().max()))


def test_episode_book_kernel_vs_oracle():
    """hg_episode_book_step (one launch per env step, no host sync) against the oracle restatement of the reference's
    per-step bookkeeping (on_policy_runner.py:140-154): same finished-episode rewards / lengths in the same order."""
    from humanoid.algo.ppo.on_policy_runner import _EpisodeBook
    from oracle.runner_oracle import episode_book_step
    N, T = 1000, 24
    g = torch.Generator().manual_seed(3)
    book = _EpisodeBook(N, T, 22, "cuda:0")
    cur_r, cur_l = torch.zeros(N), torch.zeros(N)
    means_buf = torch.zeros(22, device="cuda")
    want_r, want_l, want_infos = [], [], []
    for t in range(T):
        rew = torch.rand(N, generator=g)
        dones = torch.rand(N, generator=g) < 0.07
        means = torch.rand(22, generator=g)
        means_buf.copy_(means)
        infos = {"episode": {f"rew_{k}": means_buf[k] for k in range(22)}, "time_outs": dones.cuda()}
        book.step(t, rew.cuda(), dones.cuda(), infos)
        r, ln = episode_book_step(cur_r, cur_l, rew, dones)
        want_r += r.tolist()
        want_l += ln.tolist()
        want_infos.append(means)
    torch.cuda.synchronize()
    out = book.drain_infos()
    assert list(book.lenbuffer) == want_l[-100:]
    np.testing.assert_allclose(list(book.rewbuffer), want_r[-100:], rtol=1e-6)
    np.testing.assert_allclose([out[f"rew_{k}"] for k in range(22)], torch.stack(want_infos).mean(0).numpy(), rtol=1e-5)
    assert torch.equal(book.cur_reward_sum.cpu(), cur_r) and torch.equal(book.cur_episode_length.cpu(), cur_l)


def test_export_policy_as_jit_roundtrip(tmp_path):
    """export_policy_as_jit (reference utils/helpers.py:248-253): the TorchScript actor written from the product's
    ActorCritic reproduces the reference's shipped policy_example.pt answers on CPU and agrees with the native forward."""
    from humanoid.utils import export_policy_as_jit
    k = Golden("policy_example_kat.npz")
    ac = _make_ac(705, 219, 12, [512, 256, 128], [768, 256, 128])
    sd = ac.state_dict()
    for n, v in k.group("w.").items():
        sd["actor." + n] = v.cuda()
    ac.load_state_dict(sd)
    export_policy_as_jit(ac, str(tmp_path))
    pol = torch.jit.load(str(tmp_path / "policy_1.pt"), map_location="cpu")
    x = k.t("x")
    with torch.no_grad():
        y = pol(x)
    np.testing.assert_allclose(y.numpy(), k["y"], rtol=1e-5, atol=1e-6)
    y_native = ac.act_inference(x.cuda()).cpu()
    assert _rel(y_native, y) < 1e-5
    assert ac.actor[0].weight.is_cuda                      # exporting must not move the live module off the GPU


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 4096, 245760, 1 << 20, (1 << 20) + 1])
def test_native_randperm_is_a_permutation(n):
    """hg_randperm (the minibatch permutation of PPO.update, rollout_storage.py:155): every index exactly once, a pure
    function of (seed, counter), different for different counters / seeds."""
    from humanoid import _native as nat
    out = torch.empty(3, n, dtype=torch.int64, device="cuda")
    for k, (seed, ctr) in enumerate(((11, 0), (11, 1), (12, 0))):
        nat.check(nat.lib.hg_randperm(n, seed, ctr, out[k].data_ptr(), nat.stream_ptr(0)), "hg_randperm")
    again = torch.empty(n, dtype=torch.int64, device="cuda")
    nat.check(nat.lib.hg_randperm(n, 11, 0, again.data_ptr(), nat.stream_ptr(0)), "hg_randperm")
    torch.cuda.synchronize()
    ar = torch.arange(n, device="cuda")
    for k in range(3):
        assert torch.equal(torch.sort(out[k]).values, ar)
    assert torch.equal(out[0], again)
    if n >= 1000:
        assert not torch.equal(out[0], out[1]) and not torch.equal(out[0], out[2])
        # no visible structure: fixed points are rare, neighbours are not kept together, the first half is a fair sample
        assert int((out[0] == ar).sum()) < 12
        assert int(((out[0][1:] - out[0][:-1]).abs() == 1).sum()) < 24
        assert abs(float((out[0][: n // 2] < n // 2).double().mean()) - 0.5) < 5.0 / np.sqrt(n)


def test_native_randperm_uniform_positions():
    """Over many keys each position receives each value about equally often (chi-square on a 64-element permutation)."""
    from humanoid import _native as nat
    n, trials = 64, 8192
    out = torch.empty(trials, n, dtype=torch.int64, device="cuda")
    for t in range(trials):
        nat.check(nat.lib.hg_randperm(n, 99, t, out[t].data_ptr(), nat.stream_ptr(0)), "hg_randperm")
    torch.cuda.synchronize()
    counts = torch.zeros(n, n, device="cuda")
    counts.index_put_((torch.arange(n, device="cuda").repeat(trials), out.flatten()), torch.ones(trials * n, device="cuda"),
                      accumulate=True)
    expected = trials / n
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    dof = (n - 1) ** 2
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), (chi2, dof)


@pytest.mark.parametrize("n,seed,ctr", [(1, 3, 0), (5, 3, 1), (64, 99, 7), (1000, 2 ** 63 + 5, 2 ** 33 + 1), (245760, 11, 3)])
def test_native_randperm_bit_exact_vs_oracle(n, seed, ctr):
    """Integer work: the kernel's permutation equals the numpy restatement (oracle/perm_oracle.py) element for element."""
    from humanoid import _native as nat
    from oracle.perm_oracle import randperm
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    nat.check(nat.lib.hg_randperm(n, seed, ctr, out.data_ptr(), nat.stream_ptr(0)), "hg_randperm")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), randperm(n, seed, ctr))
