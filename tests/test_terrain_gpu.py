"""GPU parity of the rough-terrain path (csrc/hg_terrain.cu + the two-launch form of the fused env
kernel), through the C ABI via the drop-in XBotLFreeEnv:
  (1) the two-launch form of a step is bit-identical to the single fused launch (plane),
  (2) golden vectors of the UNMODIFIED reference on rough terrain, step by step (tests/golden/env_terrain.npz),
  (3) the CPU oracle on seeded random states at N = 4096 on a larger terrain,
  (4) properties of the height sampling at the benchmark sizes.
Tolerance 1e-5 relative / 1e-6 absolute; integer quantities (levels, origins picked from the table) exact.  A height
sample may differ only where the sampled point sits within 2e-4 cells of a cell boundary (index truncation)."""
import numpy as np
import pytest
import torch

from golden_io import Golden, oracle_state_from_golden
from parity_utils import (make_env, random_state, random_noise, load_state, compare_step, near_threshold_envs,
                          terrain_cfg_from_golden, terrain_params_from_golden, CHECK_KEYS)
from oracle import env_oracle as eo

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _near_cell_boundary(root_states, T, margin=2e-4):
    """(N, P) mask of height points whose cell coordinate (float64 restatement of legged_robot.py:778-783) lies within
    `margin` of an integer in x or y: fp32 rounding may put them in either cell."""
    r = root_states.double()
    qz, qw = r[:, 5], r[:, 6]
    n = torch.sqrt(qz * qz + qw * qw).clamp(min=1e-9)
    s, c = 2 * (qz / n) * (qw / n), 1 - 2 * (qz / n) ** 2          # sin / cos of the yaw
    p = T["height_points"].double()
    x = c[:, None] * p[None, :, 0] - s[:, None] * p[None, :, 1] + r[:, None, 0]
    y = s[:, None] * p[None, :, 0] + c[:, None] * p[None, :, 1] + r[:, None, 1]
    out = torch.zeros_like(x, dtype=torch.bool)
    for v in (x, y):
        cell = (v + T["border_size"]) / T["horizontal_scale"]
        out |= (cell - torch.round(cell)).abs() < margin
    return out


def _check_heights(got, want, root_states, T, what):
    bad = ~torch.isclose(got.cpu(), want, rtol=RTOL, atol=ATOL)
    if bad.any():
        unexplained = bad & ~_near_cell_boundary(root_states, T)
        assert not unexplained.any(), f"{what}: {int(unexplained.sum())} height samples off away from any cell boundary"
        assert int(bad.sum()) <= max(2, bad.numel() // 2000), f"{what}: {int(bad.sum())} boundary flips"


def test_two_launch_step_equals_fused_launch():
    """Rough terrain runs the fused kernel as {counters, callback, termination, rewards} + {reset, observations, last_*}
    with the curriculum kernel in between; on identical inputs the pair must reproduce the single launch bit for bit."""
    from humanoid import _native as nat
    from humanoid.envs.base import legged_robot as lr
    N = 4096
    g = torch.Generator().manual_seed(11)
    S, noise = random_state(N, g), random_noise(N, g)
    actions = 3.0 * torch.randn(N, 12, generator=g)
    outs = []
    for split in (False, True):
        env = make_env(N, physics="external")
        if split:
            fused = env._launch_post_physics

            def two(phases, fused=fused):
                if phases == nat.PHASE_STEP_ALL:
                    fused(lr._PHASES_BEFORE_RESET)
                    fused(lr._PHASES_FROM_RESET)
                else:
                    fused(phases)
            env._launch_post_physics = two
        load_state(env, S)
        env.inject_noise(**noise)
        env.step(actions.cuda())
        torch.cuda.synchronize()
        from parity_utils import env_value
        outs.append({k: env_value(env, k).detach().cpu().clone() for k in CHECK_KEYS})
        outs[-1]["reset_count"] = torch.tensor(env.last_reset_count)
    assert int(outs[0]["reset_buf"].sum()) > 10
    for k in outs[0]:
        if k == "episode_means":          # float atomics across CTAs: the summation order differs from launch to launch
            assert torch.allclose(outs[0][k], outs[1][k], rtol=1e-5, atol=0), k
        else:
            assert torch.equal(outs[0][k], outs[1][k]), k


def test_golden_terrain_rollout_step_by_step():
    g = Golden("env_terrain.npz")
    n, steps = int(g["meta.n_envs"]), int(g["meta.n_steps"])
    np.random.seed(int(g["meta.np_seed"]))                        # the reference's set_seed() before HumanoidTerrain is built
    env = make_env(n, physics="external", cfg=terrain_cfg_from_golden(g))
    T = terrain_params_from_golden(g)
    np.testing.assert_array_equal(env.height_samples.cpu().numpy(), g["meta.height_samples"])
    np.testing.assert_array_equal(env.terrain_origins.cpu().numpy(), g["meta.terrain_origins"])
    np.testing.assert_array_equal(env.terrain_types.cpu().numpy(), g["meta.terrain_types"])
    assert env.num_height_points == 187 and env.privileged_obs_buf.shape == (n, 3 * 892) and env.custom_origins
    assert int(env.terrain_levels.max()) <= int(g["meta.max_init_terrain_level"])

    S = oracle_state_from_golden(g)
    if not g.t("step000.post.reset_buf").any():
        # extras["episode"] is only refreshed on steps with a reset: until the first one the reference still reports the
        # means of its warm-up steps, which the init snapshot does not carry
        S["episode_means"] = g.t("step000.post.episode_means")
    hist_o, hist_p = S["obs_hist"].clone(), S["critic_hist"].clone()
    problems = []
    for t in range(steps):
        p = f"step{t:03d}."
        noise = g.group(p + "noise.")
        S["obs_hist"], S["critic_hist"] = hist_o, hist_p
        S["episode_length_buf"], S["terrain_levels"] = g.t(p + "pre.episode_length_buf"), g.t(p + "pre.terrain_levels")
        load_state(env, S)
        frames = {k: g.t(p + "pre." + k) for k in ("root_states", "contact_forces", "rigid_state")}
        dof_seq = [(g.t(p + "torque_in.dof_pos"), g.t(p + "torque_in.dof_vel")), (g.t(p + "pre.dof_pos"), g.t(p + "pre.dof_vel"))]
        calls = {"n": 0}

        def on_simulate(ph, frames=frames, dof_seq=dof_seq, calls=calls):
            calls["n"] += 1
            if calls["n"] in (9, 10):
                ds = ph.dof_state.view(n, 12, 2)
                ds[..., 0], ds[..., 1] = dof_seq[calls["n"] - 9][0].cuda(), dof_seq[calls["n"] - 9][1].cuda()
            if calls["n"] == 10:
                ph.root_states.copy_(frames["root_states"].cuda())
                ph.contact_forces.copy_(frames["contact_forces"].reshape(-1, 3).cuda())
                ph.rigid_state.copy_(frames["rigid_state"].reshape(-1, 13).cuda())
        env.gym.on_simulate = on_simulate
        env.inject_noise(**noise)
        obs, priv, rew, reset, extras = env.step(g.t(p + "actions_in").cuda())
        torch.cuda.synchronize()
        ref = {k: v for k, v in g.group(p + "post.").items()}
        keys = [k for k in ref if k not in ("obs_frame", "priv_frame", "obs_buf", "privileged_obs_buf", "measured_heights",
                                            "terrain_levels", "extras_terrain_level")]
        ref["torques"] = g.t(p + "pre.torques")
        bad = compare_step(env, ref, RTOL, ATOL, max_outlier_frac=0.0, keys=keys + ["torques"])
        if not torch.equal(env.terrain_levels.cpu(), ref["terrain_levels"]):
            bad.append("terrain_levels")
        if abs(float(extras["episode"]["terrain_level"]) - float(ref["extras_terrain_level"])) > 1e-6:
            bad.append("extras terrain_level")
        _check_heights(env.measured_heights, ref["measured_heights"], frames["root_states"], T, f"step {t}")
        flips = ~torch.isclose(env.measured_heights.cpu(), ref["measured_heights"], rtol=RTOL, atol=ATOL)
        for name, got, want in (("obs_frame", obs[:, -47:], ref["obs_frame"]), ("priv_frame", priv[:, -892:], ref["priv_frame"])):
            ok = torch.isclose(got.cpu(), want, rtol=RTOL, atol=ATOL)
            if name == "priv_frame":
                ok[:, 705:] |= flips                              # a boundary flip carries over into the frame
            if not ok.all():
                bad.append(f"{name}: max abs err {(got.cpu() - want).abs().max().item():.3g}")
        if "privileged_obs_buf" in ref and not flips.any():
            if not torch.allclose(priv.cpu(), ref["privileged_obs_buf"], rtol=RTOL, atol=ATOL):
                bad.append("privileged_obs_buf (full 3 x 892 history)")
        if bad:
            problems.append((t, bad))
        hist_o = obs.detach().cpu().view(n, 15, 47).clone()
        hist_p = priv.detach().cpu().view(n, 3, 892).clone()
        post = g.group(p + "post.")
        for k, v in post.items():
            if k in S and k not in ("obs_buf", "privileged_obs_buf"):
                S[k] = v.clone()
        S["common_step_counter"] = int(g["init.common_step_counter"]) + t + 1
    assert not problems, problems[:3]


def _big_terrain_cfg(N, measure_heights=True):
    from humanoid.envs import XBotLCfg

    class Cfg(XBotLCfg):
        class env(XBotLCfg.env):
            single_num_privileged_obs = XBotLCfg.env.num_observations + 17 * 11
            num_privileged_obs = int(XBotLCfg.env.c_frame_stack * single_num_privileged_obs)

        class terrain(XBotLCfg.terrain):
            mesh_type, curriculum = "heightfield", True
            num_rows, num_cols, border_size, max_init_terrain_level = 6, 8, 10, 5
            terrain_proportions = [0.1, 0.25, 0.25, 0.1, 0.1, 0.1, 0.1]
    Cfg.terrain.measure_heights = measure_heights
    if not measure_heights:
        Cfg.env.single_num_privileged_obs, Cfg.env.num_privileged_obs = 73, 219
    cfg = Cfg()
    cfg.seed = 5
    return cfg


@pytest.mark.parametrize("N,seed", [(4096, 21), (1000, 22)])
def test_random_state_vs_oracle_on_terrain(N, seed):
    """One full step on a 6 x 8 grid of sub-terrains: robots anywhere on the map (incl. off it: index clipping), random
    levels incl. the last one, time-outs and base contacts -> curriculum moves in every direction, spawn jitter,
    height-augmented critic frames."""
    g = torch.Generator().manual_seed(seed)
    np.random.seed(seed)
    env = make_env(N, physics="external", cfg=_big_terrain_cfg(N))
    tc = env.cfg.terrain
    T = eo.make_terrain_params(env.height_samples.cpu().numpy(), env.terrain_origins.cpu().numpy(), tc.border_size,
                               tc.horizontal_scale, tc.vertical_scale, env.terrain.env_length, True, True,
                               tc.measured_points_x, tc.measured_points_y, env.obs_scales.height_measurements)
    P = eo.make_params()
    P["terrain"] = T
    S, noise = random_state(N, g), random_noise(N, g)
    levels = torch.randint(0, tc.num_rows, (N,), generator=g)
    types = env.terrain_types.cpu()
    S["terrain_levels"], S["terrain_types"] = levels, types
    S["env_origins"] = T["terrain_origins"][levels, types].clone()
    spread = torch.where(torch.rand(N, 1, generator=g) < 0.5, 1.5, 6.0) * torch.randn(N, 2, generator=g)
    S["root_states"][:, 0:2] = S["env_origins"][:, 0:2] + spread
    S["root_states"][::53, 0:2] = torch.tensor([-30.0, 500.0])           # far off the map: clipped indices
    S["root_states"][:, 2] += S["env_origins"][:, 2]
    S["critic_hist"] = torch.randn(N, 3, 892, generator=g).clamp(-18, 18)
    S["privileged_obs_buf"] = S["critic_hist"].reshape(N, -1).clone()
    S["measured_heights"] = torch.zeros(N, 187)
    noise["u_root"] = torch.rand(N, 2, generator=g)
    noise["r_level"] = torch.randint(0, tc.num_rows, (N,), generator=g)
    actions = 2.0 * torch.randn(N, 12, generator=g)
    load_state(env, S)
    env.inject_noise(**noise)
    env.step(actions.cuda())
    torch.cuda.synchronize()

    R = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in S.items()}
    eo.pre_physics(R, P, actions, noise["u_delay"], noise["z_act"])
    eo.compute_torques(R, P)
    eo.post_physics(R, P, noise)
    reset = R["reset_buf"].bool()
    assert int(reset.sum()) > N // 100
    d = R["terrain_levels"] - levels
    assert int((d > 0).sum()) > 0 and int((d < 0).sum()) > 0 and int(((levels == tc.num_rows - 1) & reset & (d <= 0)).sum()) > 0

    _check_heights(env.measured_heights, R["measured_heights"], S["root_states"], T, "heights")
    flips = ~torch.isclose(env.measured_heights.cpu(), R["measured_heights"], rtol=RTOL, atol=ATOL)
    # curriculum: a level may differ only where the walked distance sits on one of its two thresholds
    dist = torch.norm(S["root_states"][:, :2] - S["env_origins"][:, :2], dim=1)
    need = torch.norm(S["commands"][:, :2], dim=1) * 24.0 * 0.5      # pre-step commands: resampled only after the curriculum
    edge = ((dist - env.terrain.env_length / 2).abs() < 1e-5) | ((dist - need).abs() < 1e-5)
    lv_bad = (env.terrain_levels.cpu() != R["terrain_levels"]) & ~edge
    assert not lv_bad.any(), int(lv_bad.sum())
    same_lv = env.terrain_levels.cpu() == R["terrain_levels"]
    assert torch.equal(env.env_origins.cpu()[same_lv], R["env_origins"][same_lv])
    near = near_threshold_envs(S, R, noise) | ~same_lv
    keys = [k for k in CHECK_KEYS if k != "privileged_obs_buf"]
    bad = compare_step(env, R, RTOL, ATOL, near=near, keys=keys)
    assert not bad, bad
    # spawn: base_init + (origin + U(-1, 1)) for the reset envs
    got, want = env.privileged_obs_buf.cpu().view(N, 3, 892), R["privileged_obs_buf"].view(N, 3, 892)
    ok = torch.isclose(got, want, rtol=RTOL, atol=ATOL)
    ok[:, 2, 705:] |= flips
    ok[near] = True
    assert ok.all(), int((~ok).sum())
    assert (got[reset][:, :2] == 0).all()                                   # history of a reset env is zeroed before the append


def test_heights_properties_large():
    """N = 65536 on the same terrain: flat border -> 0, constant field -> that constant, yaw invariance of the centre point."""
    from humanoid import _native as nat
    N = 65536
    np.random.seed(3)
    env = make_env(N, physics="external", cfg=_big_terrain_cfg(N, measure_heights=True))
    r = env.root_states
    r[:, 0:2] = -5.0                                                            # inside the flat border
    h = env._get_heights()
    assert h.shape == (N, 187) and float(h.abs().max()) == 0.0
    env.height_samples.fill_(37)
    r[:, 0:2] = 40.0 * torch.rand(N, 2, device=r.device)
    yaw = 6.28 * torch.rand(N, device=r.device)
    r[:, 3:7] = torch.stack((torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(yaw / 2), torch.cos(yaw / 2)), 1)
    h = env._get_heights()
    assert torch.equal(h, torch.full_like(h, 37 * np.float32(env.cfg.terrain.vertical_scale)))
    assert env._get_heights(torch.tensor([3, 5], device=r.device)).shape == (2, 187)


def test_runner_on_rough_terrain_graph_matches_eager():
    """task_registry -> OnPolicyRunner on rough terrain with the critic sized for the 3 x 892 frames: the CUDA-graph rollout
    (two fused launches + curriculum + heights + critic frames per step, per-step counters on the device) reproduces the
    eager rollout bit for bit over three collection phases, and a learning iteration runs on top of it."""
    import os
    from parity_utils import make_args
    from humanoid.utils import task_registry

    def build():
        torch.manual_seed(0)
        np.random.seed(0)
        args = make_args(256)
        env, _ = task_registry.make_env("humanoid_ppo", args=args, env_cfg=_big_terrain_cfg(256))
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
                out.append({k: getattr(s, k).clone() for k in ("observations", "privileged_observations", "actions", "rewards",
                                                               "dones", "values", "actions_log_prob")})
                out[-1]["obs"], out[-1]["cobs"], out[-1]["rew"] = obs.clone(), cobs.clone(), env.rew_buf.clone()
                out[-1]["levels"], out[-1]["origins"] = env.terrain_levels.clone(), env.env_origins.clone()
                out[-1]["heights"] = env.measured_heights.clone()
                s.clear()
        return out

    os.environ["HG_CUDA_GRAPH"] = "0"
    try:
        env_e, run_e = build()
        eager = run(run_e, env_e, 3)
    finally:
        os.environ["HG_CUDA_GRAPH"] = "1"
    env_g, run_g = build()
    assert env_g.privileged_obs_buf.shape == (256, 3 * 892) and run_g.alg.actor_critic.critic[0].in_features == 3 * 892
    graph = run(run_g, env_g, 3)
    assert getattr(run_g, "_graph", None) is not None, "the graph path did not engage"
    for ph, (a, b) in enumerate(zip(eager, graph)):
        for k in a:
            assert torch.equal(a[k], b[k]), (ph, k, float((a[k].float() - b[k].float()).abs().max()))
    assert float(eager[-1]["heights"].abs().max()) > 0 and int(eager[-1]["dones"].sum()) > 0
    w0 = run_g.alg.actor_critic.flat_params().clone()
    run_g.learn(2, init_at_random_ep_len=True)
    assert torch.isfinite(run_g.alg.actor_critic.flat_params()).all() and not torch.equal(w0, run_g.alg.actor_critic.flat_params())
