"""Rough-terrain path, CPU side: the product's Terrain / HumanoidTerrain classes and the oracle's
terrain functions against tests/golden/env_terrain.npz, which tests/golden/make_golden.py captured from the UNMODIFIED
reference (utils/terrain.py HumanoidTerrain, legged_robot.py _get_heights / _update_terrain_curriculum / custom-origin
resets, humanoid_env.py height-augmented critic frames).  The Isaac Gym terrain primitives underneath are third-party
and parity-unpinned (see utils/terrain.py)."""
import numpy as np
import pytest
import torch

from golden_io import Golden, oracle_state_from_golden
from oracle import env_oracle as eo
from parity_utils import terrain_cfg_from_golden, terrain_params_from_golden

RTOL, ATOL = 1e-6, 1e-7


@pytest.fixture(scope="module")
def g():
    return Golden("env_terrain.npz")


def test_humanoid_terrain_matches_reference(g):
    """Same numpy seed, same cfg -> the height field, the platform origins and the triangle mesh the reference built."""
    from humanoid.utils.terrain import HumanoidTerrain
    cfg = terrain_cfg_from_golden(g)
    np.random.seed(int(g["meta.np_seed"]))
    t = HumanoidTerrain(cfg.terrain, int(g["meta.n_envs"]))
    np.testing.assert_array_equal(t.heightsamples, g["meta.height_samples"])
    np.testing.assert_array_equal(t.env_origins, g["meta.terrain_env_origins_f64"])
    assert t.vertices.shape[0] == int(g["meta.n_vertices"]) and t.triangles.shape[0] == int(g["meta.n_triangles"])
    np.testing.assert_array_equal(t.vertices[:4096], g["meta.vertices_head"])
    np.testing.assert_array_equal(t.triangles[:4096], g["meta.triangles_head"])
    assert float(t.vertices.astype(np.float64).sum()) == float(g["meta.vertices_sum"])
    assert t.height_field_raw.dtype == np.int16 and (t.tot_rows, t.tot_cols) == t.height_field_raw.shape


def test_terrain_variants_build():
    """Every sub-terrain family of Terrain.make_terrain / HumanoidTerrain.make_terrain, all three population modes."""
    from humanoid.envs.base.legged_robot_config import LeggedRobotCfg
    from humanoid.utils.terrain import Terrain, HumanoidTerrain

    def cfg(**kw):
        class T(LeggedRobotCfg.terrain):
            pass
        for k, v in kw.items():
            setattr(T, k, v)
        return T
    np.random.seed(1)
    base = dict(num_rows=3, num_cols=8, border_size=2, terrain_proportions=[0.1, 0.1, 0.2, 0.2, 0.1, 0.1, 0.1, 0.1])
    t = Terrain(cfg(mesh_type="heightfield", curriculum=True, **base), 8)
    assert t.height_field_raw.min() <= -200 and t.height_field_raw.max() > 50      # gaps / pits and stairs are there
    assert not hasattr(t, "vertices")
    t = Terrain(cfg(mesh_type="trimesh", curriculum=False, **base), 8)
    assert t.triangles.shape == (2 * (t.tot_rows - 1) * (t.tot_cols - 1), 3) and t.triangles.dtype == np.uint32
    assert t.vertices.shape == (t.tot_rows * t.tot_cols, 3)
    t = Terrain(cfg(mesh_type="heightfield", curriculum=False, selected=True,
                    terrain_kwargs=dict(type="pyramid_stairs_terrain", step_width=0.3, step_height=0.1, platform_size=2.0), **base), 8)
    assert t.env_origins[:, :, 2].max() > 0.5                                       # spawn on top of the stairs
    t = HumanoidTerrain(cfg(mesh_type="trimesh", curriculum=True, num_rows=3, num_cols=7, border_size=2,
                            terrain_proportions=[0.1, 0.1, 0.2, 0.2, 0.2, 0.1, 0.1]), 8)
    assert t.height_field_raw[:20, :].max() == 0 and t.height_field_raw.max() > 0   # flat border, rough inside
    assert Terrain(cfg(mesh_type="plane"), 8).type == "plane"


def _cmp(name, a, b, t):
    a = a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    if a.dtype == np.bool_ or b.dtype == np.bool_ or a.dtype.kind in "iu":
        assert np.array_equal(a.reshape(b.shape).astype(np.int64), b.astype(np.int64)), f"{name} @ step {t}"
    else:
        np.testing.assert_allclose(a.reshape(b.shape), b, rtol=RTOL, atol=ATOL, err_msg=f"{name} @ step {t}")


def test_rollout_replay_matches_reference(g):
    """Chain all recorded rough-terrain steps on the oracle: heights, curriculum levels / origins, spawn positions and
    the 3 x 892 critic history must follow the reference step for step."""
    P = eo.make_params()
    P["terrain"] = terrain_params_from_golden(g)
    S = oracle_state_from_golden(g)
    S["terrain_levels"], S["terrain_types"] = g.t("init.terrain_levels"), g.t("meta.terrain_types")
    S["measured_heights"] = g.t("init.measured_heights")
    n_steps = int(g["meta.n_steps"])
    ups = downs = wraps = 0
    for t in range(n_steps):
        p = f"step{t:03d}."
        noise = g.group(p + "noise.")
        eo.pre_physics(S, P, g.t(p + "actions_in"), noise["u_delay"], noise["z_act"])
        S["dof_pos"], S["dof_vel"] = g.t(p + "torque_in.dof_pos"), g.t(p + "torque_in.dof_vel")
        eo.compute_torques(S, P)
        _cmp("torques", S["torques"], g[p + "pre.torques"], t)
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state", "episode_length_buf", "terrain_levels"):
            S[k] = g.t(p + "pre." + k).clone()          # the stand-in physics moved robots / the script forced time-outs
        lv0 = S["terrain_levels"].clone()
        obs, priv, rew, reset = eo.post_physics(S, P, noise)
        post = g.group(p + "post.")
        for k in ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "commands", "episode_length_buf", "reset_buf",
                  "time_out_buf", "base_lin_vel", "base_euler_xyz", "feet_air_time", "rew_buf", "episode_sums",
                  "env_origins", "terrain_levels", "measured_heights"):
            _cmp(k, S[k], post[k], t)
        _cmp("obs_frame", obs[:, -47:], post["obs_frame"], t)
        _cmp("priv_frame", priv[:, -892:], post["priv_frame"], t)
        if "privileged_obs_buf" in post:
            _cmp("privileged_obs_buf", priv, post["privileged_obs_buf"], t)
        assert abs(float(S["terrain_levels"].float().mean()) - float(g[p + "post.extras_terrain_level"])) < 1e-6
        d = S["terrain_levels"] - lv0
        r = reset.bool()
        wraps += int((r & (lv0 == P["terrain"]["max_terrain_level"] - 1) & (noise["r_level"] == S["terrain_levels"]) & (d != -1)).sum())
        ups += int((d == 1).sum())
        downs += int((d == -1).sum())
    assert ups >= 3 and downs >= 3 and wraps >= 1, (ups, downs, wraps)       # the fixture exercises every curriculum branch
    assert float(g["meta.height_samples"].max()) > 0 and float(np.abs(g["step010.post.measured_heights"]).max()) > 0
