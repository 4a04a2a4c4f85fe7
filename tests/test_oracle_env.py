"""Pin the env oracle (oracle/env_oracle.py) against golden vectors captured from the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from golden_io import Golden, oracle_state_from_golden
from oracle import env_oracle as eo

# The oracle restates the same fp32 torch ops in the same order -> expect (near) bit equality on CPU.
RTOL, ATOL = 1e-6, 1e-7


@pytest.fixture(scope="module")
def g():
    return Golden("env_rollout.npz")


def test_params_match_reference_derivation(g):
    P = eo.make_params()
    assert P["dt"] == float(g["meta.dt"])
    assert P["resample_period"] == int(g["meta.resample_period"]) == 799
    assert P["max_episode_length"] == float(g["meta.max_episode_length"]) == 2400.0
    assert P["push_interval"] == float(g["meta.push_interval"]) == 400.0
    assert list(g["meta.reward_names"]) == list(eo.REWARD_NAMES)
    np.testing.assert_array_equal(np.array(P["reward_scales"]), g["meta.reward_scales"])
    np.testing.assert_array_equal(P["p_gains"].numpy(), g["meta.p_gains"])
    np.testing.assert_array_equal(P["d_gains"].numpy(), g["meta.d_gains"])
    np.testing.assert_array_equal(P["torque_limits"].numpy(), g["meta.torque_limits"])
    np.testing.assert_array_equal(P["noise_scale_vec"].numpy(), g["meta.noise_scale_vec"])
    assert tuple(g["meta.feet_indices"]) == P["feet"] and tuple(g["meta.knee_indices"]) == P["knees"]
    assert tuple(g["meta.termination_contact_indices"]) == P["term_bodies"]
    np.testing.assert_array_equal(np.array(P["base_init_state"], np.float32), g["meta.base_init_state"])


def test_grid_origins(g):
    o = eo.grid_origins(int(g["meta.n_envs"]))
    np.testing.assert_array_equal(o.numpy(), g["init.env_origins"])


def _cmp(name, a, b, t):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    if a.dtype == np.bool_ or b.dtype == np.bool_:
        assert np.array_equal(a.astype(bool).reshape(b.shape), b.astype(bool)), f"{name} @ step {t}"
    else:
        np.testing.assert_allclose(a.reshape(b.shape), b, rtol=RTOL, atol=ATOL, err_msg=f"{name} @ step {t}")


def test_rollout_replay_matches_reference(g):
    """Chain all recorded steps from the initial snapshot; every step's outputs must match."""
    P = eo.make_params()
    S = oracle_state_from_golden(g)
    n_steps = int(g["meta.n_steps"])
    n_resets = n_timeouts = n_push = 0
    for t in range(n_steps):
        p = f"step{t:03d}."
        noise = g.group(p + "noise.")
        # E1/E2
        eo.pre_physics(S, P, g.t(p + "actions_in"), noise["u_delay"], noise["z_act"])
        _cmp("actions(pre)", S["actions"], g[p + "pre.actions"], t)
        # E3 on the dof state the last decimation sub-step saw
        S["dof_pos"], S["dof_vel"] = g.t(p + "torque_in.dof_pos"), g.t(p + "torque_in.dof_vel")
        eo.compute_torques(S, P)
        _cmp("torques", S["torques"], g[p + "pre.torques"], t)
        # physics stand-in output
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state"):
            S[k] = g.t(p + "pre." + k).clone()
        obs, priv, rew, reset = eo.post_physics(S, P, noise)
        post = g.group(p + "post.")
        for k in ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_last_actions",
                  "last_dof_vel", "last_root_vel", "commands", "episode_length_buf", "reset_buf",
                  "time_out_buf", "base_lin_vel", "base_ang_vel", "projected_gravity", "base_euler_xyz",
                  "feet_air_time", "last_contacts", "feet_height", "last_feet_z", "ref_dof_pos",
                  "rand_push_force", "rand_push_torque", "rew_buf", "episode_sums"):
            _cmp(k, S[k], post[k], t)
        _cmp("obs_frame", obs[:, -47:], post["obs_frame"], t)
        _cmp("priv_frame", priv[:, -73:], post["priv_frame"], t)
        if "obs_buf" in post:
            _cmp("obs_buf", obs, post["obs_buf"], t)
            _cmp("privileged_obs_buf", priv, post["privileged_obs_buf"], t)
        _cmp("extras_time_outs", S["extras_time_outs"], post["extras_time_outs"], t)
        _cmp("episode_means", S["episode_means"], post["episode_means"], t)
        n_resets += int(reset.sum())
        n_timeouts += int(S["time_out_buf"].sum())
        n_push += int(S["common_step_counter"] % 400 == 0)
    # the fixture must actually exercise the rare branches
    assert n_resets > 10 and n_timeouts >= 3 and n_push == 1


def test_command_curriculum_replay_matches_reference():
    """update_command_curriculum (legged_robot.py:178-180,422-431) inside chained steps: tests/golden/env_cmd_curriculum.npz
    (unmodified reference with commands.curriculum on; the range widens twice -- the second time into the max_curriculum clip
    -- and stays put when the resetting envs tracked badly or the counter is off the multiple of max_episode_length)."""
    g = Golden("env_cmd_curriculum.npz")
    P = eo.make_params()
    P["cmd_curriculum"], P["max_curriculum"] = True, float(g["meta.max_curriculum"])
    P["cmd_x"] = tuple(g["meta.init_range_x"])
    S = oracle_state_from_golden(g)
    widened = 0
    for t in range(int(g["meta.n_steps"])):
        p = f"step{t:03d}."
        noise = g.group(p + "noise.")
        eo.pre_physics(S, P, g.t(p + "actions_in"), noise["u_delay"], noise["z_act"])
        S["dof_pos"], S["dof_vel"] = g.t(p + "torque_in.dof_pos"), g.t(p + "torque_in.dof_vel")
        eo.compute_torques(S, P)
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state", "episode_length_buf", "episode_sums"):
            S[k] = g.t(p + "pre." + k).clone()
        S["common_step_counter"] = int(g[p + "pre.common_step_counter"])
        before = P["cmd_x"]
        obs, priv, rew, reset = eo.post_physics(S, P, noise)
        post = g.group(p + "post.")
        assert tuple(float(x) for x in P["cmd_x"]) == tuple(float(x) for x in g[p + "post.range_x"]), t
        widened += P["cmd_x"] != before
        for k in ("commands", "root_states", "reset_buf", "episode_sums", "rew_buf", "episode_length_buf"):
            _cmp(k, S[k], post[k], t)
        _cmp("obs_frame", obs[:, -47:], post["obs_frame"], t)
        _cmp("episode_means", S["episode_means"], post["episode_means"], t)
        assert S["common_step_counter"] == int(g[p + "post.common_step_counter"])
    assert widened == 2 and float(P["cmd_x"][1]) == 1.5
