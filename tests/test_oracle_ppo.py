"""Pin the learning-side oracle (oracle/ppo_oracle.py) against vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from golden_io import Golden
from oracle import ppo_oracle as po


@pytest.fixture(scope="module")
def g():
    return Golden("ppo_learning.npz")


def test_param_order(g):
    assert list(g["param_order"]) == po.param_names()


def test_forward_logprob_entropy(g):
    p = g.group("w0.")
    mean, sigma = po.actor_dist(g.t("fwd.obs"), p)
    np.testing.assert_allclose(mean.numpy(), g["fwd.mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sigma.numpy(), g["fwd.std"], rtol=0, atol=0)
    np.testing.assert_allclose(po.log_prob(g.t("fwd.actions"), mean, sigma).numpy(), g["fwd.logp"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(po.entropy(sigma).numpy(), g["fwd.entropy"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(po.mlp(g.t("fwd.cobs"), p, "critic").numpy(), g["fwd.value"], rtol=1e-6, atol=1e-7)


def test_policy_example_known_answers():
    """The reference's only shipped fixture: actor 705-512-256-128-12."""
    k = Golden("policy_example_kat.npz")
    p = {"actor." + n: v for n, v in k.group("w.").items()}
    y = po.mlp(k.t("x"), p, "actor")
    np.testing.assert_allclose(y.numpy(), k["y"], rtol=1e-5, atol=1e-6)
    kat = [0.0847, -0.0234, 0.0057, 0.2348, 0.6382, -0.2275, -0.1129, -0.1501, 0.2042, 0.3535, 0.0077, -0.4530]
    np.testing.assert_allclose(y[0].numpy(), kat, atol=5e-5)


def test_gae(g):
    r, a = po.gae(g.t("gae.rewards"), g.t("gae.values"), g.t("gae.dones"), g.t("gae.last_values"), 0.994, 0.9)
    np.testing.assert_allclose(r.numpy(), g["gae.returns"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(a.numpy(), g["gae.advantages"], rtol=1e-5, atol=1e-6)


def test_full_update(g):
    st = g.group("upd.storage.")
    # returns/advantages in the golden came from compute_returns on the same storage: recheck
    last_v = po.mlp(g.t("upd.last_cobs"), g.group("w0."), "critic")
    r, a = po.gae(st["rewards"], st["values"], st["dones"], last_v, 0.994, 0.9)
    np.testing.assert_allclose(r.numpy(), st["returns"].numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(a.numpy(), st["advantages"].numpy(), rtol=1e-5, atol=1e-6)

    L = po.Learner(g.group("w0."), lr=1e-5)
    L.grad_log = []
    mv, ms = L.update(st, g.t("upd.perm"))
    grads = np.stack([x[0].numpy() for x in L.grad_log])
    lrs = np.array([x[1] for x in L.grad_log])
    np.testing.assert_allclose(lrs, g["upd.lrs"], rtol=1e-12)
    ref = g["upd.grads"]
    for i in range(ref.shape[0]):
        rel = np.linalg.norm(grads[i] - ref[i]) / np.linalg.norm(ref[i])
        assert rel < 1e-5, f"gradient {i}: rel L2 {rel}"
    assert abs(mv - float(g["upd.mean_value_loss"])) < 1e-5 * max(1, abs(mv))
    assert abs(ms - float(g["upd.mean_surrogate_loss"])) < 1e-5
    assert L.lr == float(g["upd.final_lr"])
    for k, v in g.group("w1.").items():
        np.testing.assert_allclose(L.p[k].detach().numpy(), v.numpy(), rtol=1e-5, atol=1e-7, err_msg=k)


def test_perm_oracle_known_answers():
    """oracle/perm_oracle.py: Philox4x32-10 against the Random123 known-answer vectors; the restated hg_randperm is a
    permutation for every size, a pure function of (seed, counter)."""
    import numpy as np
    from oracle.perm_oracle import philox4x32_10, randperm
    assert philox4x32_10(0, 0, 0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32_10(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == \
        (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    for n in (1, 2, 3, 31, 32, 33, 1000, 4097):
        p = randperm(n, 5, 0)
        assert np.array_equal(np.sort(p), np.arange(n))
        assert np.array_equal(p, randperm(n, 5, 0))
        if n >= 31:
            assert not np.array_equal(p, randperm(n, 5, 1)) and not np.array_equal(p, randperm(n, 6, 0))
