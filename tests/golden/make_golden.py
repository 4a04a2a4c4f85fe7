#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  The
reference's Python is imported as-is; `isaacgym` is the test-only fake in
tests/golden/fake_isaacgym and `matplotlib` is stubbed.
Random draws made by the reference are RECORDED (not replaced) and densified
to per-env tensors so that the oracle / CUDA kernels can be fed the same
numbers.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Outputs
  env_rollout.npz        : initial state + per-step inputs/outputs of env.step()
  ppo_learning.npz       : ActorCritic fwd, GAE, one PPO.update() incl. gradients
  policy_example_kat.npz : weights + known answers of the reference's only
                           shipped fixture (logs/XBot_ppo/exported/policies/policy_example.pt)
  env_cmd_curriculum.npz : the same env with commands.curriculum on: steps on which common_step_counter hits a multiple of
                           max_episode_length with envs resetting (range widened twice, clipped at max_curriculum, then a
                           step whose resetting envs tracked badly -> unchanged)
  env_terrain.npz        : the same env on rough terrain (mesh_type 'trimesh', terrain curriculum, measured heights
                           in the critic frames): HumanoidTerrain's height field + chained env.step() calls
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HG_REFERENCE_ROOT", "/root/reference")
_REAL = dict(rand=torch.rand, randn_like=torch.randn_like, randint_like=torch.randint_like)   # instrument_env patches these


def _install_shims():
    sys.path.insert(0, os.path.join(HERE, "fake_isaacgym"))
    sys.path.insert(0, REF)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    os.environ.setdefault("WANDB_MODE", "disabled")


# --------------------------------------------------------------------------
# env goldens
# --------------------------------------------------------------------------
STATE_KEYS = ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state", "actions",
              "last_actions", "last_last_actions", "torques", "last_dof_vel", "last_root_vel",
              "commands", "episode_length_buf", "reset_buf", "time_out_buf", "base_lin_vel",
              "base_ang_vel", "projected_gravity", "base_euler_xyz", "feet_air_time",
              "last_contacts", "feet_height", "ref_dof_pos", "rand_push_force", "rand_push_torque",
              "env_frictions", "body_mass", "rew_buf", "env_origins")


def snap(env):
    out = {}
    for k in STATE_KEYS:
        v = getattr(env, k)
        out[k] = v.detach().clone().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    n = env.num_envs
    lf = env.last_feet_z
    out["last_feet_z"] = lf.clone().numpy() if isinstance(lf, torch.Tensor) else np.full((n, 2), lf, np.float32)
    out["episode_sums"] = np.stack([env.episode_sums[k].numpy().copy() for k in env.reward_names])
    out["obs_hist"] = torch.stack(list(env.obs_history), dim=1).numpy().copy()
    out["critic_hist"] = torch.stack(list(env.critic_history), dim=1).numpy().copy()
    out["obs_buf"] = env.obs_buf.numpy().copy()
    out["privileged_obs_buf"] = env.privileged_obs_buf.numpy().copy()
    out["common_step_counter"] = np.int64(env.common_step_counter)
    return out


class DrawRecorder:
    """Records the reference's random draws and densifies them per env."""

    def __init__(self, n):
        self.n = n
        self.ctx = None
        self.in_reset = False
        self.new_step()

    def new_step(self):
        n = self.n
        self.u_cmd_cb = np.zeros((n, 3), np.float32)
        self.u_cmd_rs = np.zeros((n, 3), np.float32)
        self.u_dof = np.zeros((n, 12), np.float32)
        self.u_push = np.zeros((n, 5), np.float32)
        self.z_obs = np.zeros((n, 47), np.float32)
        self.u_delay = np.zeros((n, 1), np.float32)
        self.z_act = np.zeros((n, 12), np.float32)
        self.u_root = np.zeros((n, 2), np.float32)       # rough terrain: spawn jitter (legged_robot.py:384)
        self.r_level = np.zeros((n,), np.int64)          # rough terrain: random level after the last one (:418)
        self._cmd_call = 0

    def noise(self, terrain=False):
        keys = ("u_cmd_cb", "u_cmd_rs", "u_dof", "u_push", "z_obs", "u_delay", "z_act") + (("u_root", "r_level") if terrain else ())
        return {k: getattr(self, k).copy() for k in keys}

    # replacement for isaacgym.torch_utils.torch_rand_float inside the reference modules
    def torch_rand_float(self, lower, upper, shape, device):
        u = torch.rand(*shape, device=device)
        kind, ids = self.ctx
        if kind == "cmd":
            dst = self.u_cmd_rs if self.in_reset else self.u_cmd_cb
            if len(ids):
                dst[ids.numpy(), self._cmd_call] = u[:, 0].numpy()
            self._cmd_call += 1
        elif kind == "dof":
            self.u_dof[ids.numpy()] = u.numpy()
        elif kind == "push":
            if shape[1] == 2:
                self.u_push[:, 0:2] = u.numpy()
            else:
                self.u_push[:, 2:5] = u.numpy()
        elif kind == "root":
            self.u_root[ids.numpy()] = u.numpy()
        else:
            raise RuntimeError(f"unexpected draw in ctx {self.ctx}")
        return (upper - lower) * u + lower


def uninstrument():
    """Undo instrument_env's module-level patches (a second env built in the same process must draw un-recorded)."""
    import humanoid.envs.base.legged_robot as lr
    import humanoid.envs.custom.humanoid_env as he
    from isaacgym.torch_utils import torch_rand_float
    lr.torch_rand_float = he.torch_rand_float = torch_rand_float
    torch.rand, torch.randn_like, torch.randint_like = _REAL["rand"], _REAL["randn_like"], _REAL["randint_like"]


def instrument_env(env, rec):
    import humanoid.envs.base.legged_robot as lr
    import humanoid.envs.custom.humanoid_env as he
    lr.torch_rand_float = rec.torch_rand_float
    he.torch_rand_float = rec.torch_rand_float

    orig_resample = env._resample_commands
    orig_reset_dofs = env._reset_dofs
    orig_push = env._push_robots
    orig_reset_idx = env.reset_idx

    def resample(env_ids):
        rec.ctx = ("cmd", env_ids.clone())
        rec._cmd_call = 0
        orig_resample(env_ids)
        rec.ctx = None

    def reset_dofs(env_ids):
        rec.ctx = ("dof", env_ids.clone())
        orig_reset_dofs(env_ids)
        rec.ctx = None

    def push():
        rec.ctx = ("push", None)
        orig_push()
        rec.ctx = None

    def reset_idx(env_ids):
        rec.in_reset = True
        orig_reset_idx(env_ids)
        rec.in_reset = False

    env._resample_commands = resample
    env._reset_dofs = reset_dofs
    env._push_robots = push
    env.reset_idx = reset_idx

    if env.custom_origins:                               # rough terrain: two more draws inside reset_idx
        orig_root, orig_curr = env._reset_root_states, env._update_terrain_curriculum

        def reset_root_states(env_ids):
            rec.ctx = ("root", env_ids.clone())
            orig_root(env_ids)
            rec.ctx = None

        def update_terrain_curriculum(env_ids):
            def randint_like(t, high, **k):
                r = _REAL["randint_like"](t, high, **k)
                rec.r_level[env_ids.numpy()] = r.numpy()
                return r
            torch.randint_like = randint_like
            try:
                orig_curr(env_ids)
            finally:
                torch.randint_like = _REAL["randint_like"]

        env._reset_root_states = reset_root_states
        env._update_terrain_curriculum = update_terrain_curriculum

    real_rand, real_randn_like = _REAL["rand"], _REAL["randn_like"]

    def rand(*a, **k):
        u = real_rand(*a, **k)
        if rec.ctx is None and u.shape == (env.num_envs, 1):
            rec.u_delay[:] = u.numpy()
        return u

    def randn_like(t, **k):
        z = real_randn_like(t, **k)
        if z.shape == (env.num_envs, 47):
            rec.z_obs[:] = z.numpy()
        elif z.shape == (env.num_envs, 12):
            rec.z_act[:] = z.numpy()
        return z

    torch.rand = rand
    torch.randn_like = randn_like


def make_env_golden(out_path, n_envs=24, n_steps=56):
    from humanoid.envs import XBotLFreeEnv  # noqa: F401  (registers humanoid_ppo)
    from humanoid.utils import task_registry

    args = argparse.Namespace(
        task="humanoid_ppo", resume=False, experiment_name=None, run_name=None, load_run=None,
        checkpoint=None, headless=True, horovod=False, rl_device="cpu", num_envs=n_envs, seed=5,
        max_iterations=None, physics_engine=1, use_gpu=False, use_gpu_pipeline=False, subscenes=0,
        num_threads=0, sim_device="cpu", sim_device_type="cpu", compute_device_id=0, sim_device_id=0,
        device="cpu")
    env, cfg = task_registry.make_env(name="humanoid_ppo", args=args)
    assert env.dt == 10 * float(np.float32(0.001))
    meta = dict(
        dt=np.float64(env.dt), max_episode_length=np.float64(env.max_episode_length),
        push_interval=np.float64(env.cfg.domain_rand.push_interval),
        resample_period=np.int64(int(env.cfg.commands.resampling_time / env.dt)),
        feet_indices=env.feet_indices.numpy(), knee_indices=env.knee_indices.numpy(),
        termination_contact_indices=env.termination_contact_indices.numpy(),
        penalised_contact_indices=env.penalised_contact_indices.numpy(),
        p_gains=env.p_gains[0].numpy(), d_gains=env.d_gains[0].numpy(),
        torque_limits=env.torque_limits.numpy(), default_dof_pos=env.default_dof_pos[0].numpy(),
        noise_scale_vec=env.noise_scale_vec.numpy(),
        reward_names=np.array(env.reward_names),
        reward_scales=np.array([env.reward_scales[k] for k in env.reward_names], np.float64),
        dof_names=np.array(env.dof_names), base_init_state=env.base_init_state.numpy(),
    )

    rec = DrawRecorder(n_envs)
    instrument_env(env, rec)

    # make rare branches reachable inside a short rollout
    g = torch.Generator().manual_seed(77)
    ep = torch.randint(0, 2390, (n_envs,), generator=g)
    ep[0], ep[1], ep[2] = 2399, 2400, 2398          # time-outs (ep_len > 2400 after increment)
    ep[3], ep[4], ep[5] = 796, 1597, 2396           # command resampling at multiples of 799
    env.episode_length_buf[:] = ep
    env.common_step_counter = 390                    # push at the 10th step (counter % 400 == 0)
    env.gym_sim = env.sim

    last_torque_in = {}
    orig_ct = env._compute_torques

    def compute_torques(actions):
        last_torque_in["dof_pos"] = env.dof_pos.clone().numpy()
        last_torque_in["dof_vel"] = env.dof_vel.clone().numpy()
        return orig_ct(actions)

    env._compute_torques = compute_torques

    data = {f"meta.{k}": v for k, v in meta.items()}
    for k, v in snap(env).items():
        data[f"init.{k}"] = v

    orig_pps = env.post_physics_step
    pre_post = {}

    def post_physics_step():
        pre_post["pre"] = snap(env)
        orig_pps()

    env.post_physics_step = post_physics_step

    ag = torch.Generator().manual_seed(99)
    n_reset = 0
    for t in range(n_steps):
        rec.new_step()
        act_in = 3.0 * torch.randn(n_envs, 12, generator=ag)
        if t % 7 == 3:
            act_in[::5] *= 20.0                      # exercise the +-18 action clip
        obs, priv, rew, reset, extras = env.step(act_in.clone())
        noise = rec.noise()
        pre = pre_post["pre"]
        post = snap(env)
        p = f"step{t:03d}."
        data[p + "actions_in"] = act_in.numpy()
        for k, v in noise.items():
            data[p + "noise." + k] = v
        data[p + "torque_in.dof_pos"] = last_torque_in["dof_pos"]
        data[p + "torque_in.dof_vel"] = last_torque_in["dof_vel"]
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state", "actions", "torques"):
            data[p + "pre." + k] = pre[k]
        for k in STATE_KEYS + ("last_feet_z", "episode_sums"):
            if k in ("contact_forces", "rigid_state", "env_frictions", "body_mass", "env_origins"):
                continue
            data[p + "post." + k] = post[k]
        data[p + "post.obs_frame"] = obs[:, -47:].numpy().copy()
        data[p + "post.priv_frame"] = priv[:, -73:].numpy().copy()
        if t % 8 == 7 or t == n_steps - 1:
            data[p + "post.obs_buf"] = obs.numpy().copy()
            data[p + "post.privileged_obs_buf"] = priv.numpy().copy()
        data[p + "post.extras_time_outs"] = extras["time_outs"].numpy().copy()
        data[p + "post.episode_means"] = np.array(
            [float(extras["episode"]["rew_" + k]) for k in env.reward_names], np.float32)
        n_reset += int(reset.sum())
    data["meta.n_steps"] = np.int64(n_steps)
    data["meta.n_envs"] = np.int64(n_envs)
    np.savez_compressed(out_path, **data)
    print(f"wrote {out_path}: {n_steps} steps x {n_envs} envs, {n_reset} resets,"
          f" {os.path.getsize(out_path) / 1e6:.2f} MB")


def make_terrain_golden(out_path, n_envs=20, n_steps=24):
    """XBotLFreeEnv on rough terrain: HumanoidTerrain (reference utils/terrain.py, UNMODIFIED) over the restated Isaac Gym
    primitives, `_get_heights`, `_update_terrain_curriculum`, the custom-origin spawn and the height-augmented critic frames
    (humanoid_env.py:246-248; the cfg sizes the critic for them, otherwise the reference fails in its first matmul)."""
    from humanoid.envs import XBotLCfg, XBotLFreeEnv  # noqa: F401
    from humanoid.utils import task_registry
    uninstrument()

    class TerrainCfg(XBotLCfg):
        class env(XBotLCfg.env):
            single_num_privileged_obs = XBotLCfg.env.num_observations + 17 * 11
            num_privileged_obs = int(XBotLCfg.env.c_frame_stack * single_num_privileged_obs)

        class terrain(XBotLCfg.terrain):
            mesh_type, curriculum, measure_heights = "trimesh", True, True
            num_rows, num_cols, border_size, max_init_terrain_level = 4, 5, 5, 3
            terrain_proportions = [0.1, 0.25, 0.25, 0.1, 0.1, 0.1, 0.1]

    args = argparse.Namespace(
        task="humanoid_ppo", resume=False, experiment_name=None, run_name=None, load_run=None,
        checkpoint=None, headless=True, horovod=False, rl_device="cpu", num_envs=n_envs, seed=5,
        max_iterations=None, physics_engine=1, use_gpu=False, use_gpu_pipeline=False, subscenes=0,
        num_threads=0, sim_device="cpu", sim_device_type="cpu", compute_device_id=0, sim_device_id=0,
        device="cpu")
    env_cfg = TerrainCfg()
    env_cfg.seed = 5                                 # task_registry.get_cfgs copies the runner's seed (task_registry.py:74-76)
    env, cfg = task_registry.make_env(name="humanoid_ppo", args=args, env_cfg=env_cfg)
    tc = env.cfg.terrain
    data = {
        "meta.height_samples": env.height_samples.numpy().copy(), "meta.terrain_origins": env.terrain_origins.numpy().copy(),
        "meta.terrain_env_origins_f64": env.terrain.env_origins.copy(),
        "meta.vertices_head": env.terrain.vertices[:4096].copy(), "meta.triangles_head": env.terrain.triangles[:4096].copy(),
        "meta.n_vertices": np.int64(env.terrain.vertices.shape[0]), "meta.n_triangles": np.int64(env.terrain.triangles.shape[0]),
        "meta.vertices_sum": np.float64(env.terrain.vertices.astype(np.float64).sum()),
        "meta.terrain_types": env.terrain_types.numpy().copy(), "meta.init_terrain_levels": env.terrain_levels.numpy().copy(),
        "meta.border_size": np.float64(tc.border_size), "meta.horizontal_scale": np.float64(tc.horizontal_scale),
        "meta.vertical_scale": np.float64(tc.vertical_scale), "meta.env_length": np.float64(env.terrain.env_length),
        "meta.num_rows": np.int64(tc.num_rows), "meta.num_cols": np.int64(tc.num_cols),
        "meta.max_init_terrain_level": np.int64(tc.max_init_terrain_level),
        "meta.terrain_proportions": np.array(tc.terrain_proportions, np.float64),
        "meta.measured_points_x": np.array(tc.measured_points_x, np.float64),
        "meta.measured_points_y": np.array(tc.measured_points_y, np.float64),
        "meta.height_scale": np.float64(env.obs_scales.height_measurements),
        "meta.np_seed": np.int64(cfg.seed), "meta.n_steps": np.int64(n_steps), "meta.n_envs": np.int64(n_envs),
    }
    rec = DrawRecorder(n_envs)
    instrument_env(env, rec)

    def tsnap():
        out = snap_terrain(env)
        return out

    def snap_terrain(env):
        out = {}
        for k in STATE_KEYS:
            v = getattr(env, k)
            out[k] = v.detach().clone().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        lf = env.last_feet_z
        out["last_feet_z"] = lf.clone().numpy() if isinstance(lf, torch.Tensor) else np.full((n_envs, 2), lf, np.float32)
        out["episode_sums"] = np.stack([env.episode_sums[k].numpy().copy() for k in env.reward_names])
        out["obs_hist"] = torch.stack(list(env.obs_history), dim=1).numpy().copy()
        out["critic_hist"] = torch.stack(list(env.critic_history), dim=1).numpy().copy()
        out["obs_buf"] = env.obs_buf.numpy().copy()
        out["privileged_obs_buf"] = env.privileged_obs_buf.numpy().copy()
        out["common_step_counter"] = np.int64(env.common_step_counter)
        out["terrain_levels"] = env.terrain_levels.numpy().copy()
        mh = env.measured_heights
        out["measured_heights"] = mh.numpy().copy() if isinstance(mh, torch.Tensor) else np.zeros((n_envs, 187), np.float32)
        return out

    g = torch.Generator().manual_seed(78)
    ep = torch.randint(0, 2300, (n_envs,), generator=g)
    env.episode_length_buf[:] = ep
    # The reference's FIRST critic frame is malformed in this mode: __init__ calls compute_observations() while
    # measured_heights is still the scalar 0 (legged_robot.py:483), so the frame is 705 + 1 wide and privileged_obs_buf only
    # reaches its nominal 3 x 892 columns once c_frame_stack steps have pushed it out.  The fixture starts from there.
    for _ in range(env.cfg.env.c_frame_stack):
        env.step(torch.zeros(n_envs, 12))
    assert env.privileged_obs_buf.shape == (n_envs, 3 * 892)
    env.common_step_counter = 395                    # a push inside the rollout
    for k, v in tsnap().items():
        data[f"init.{k}"] = v

    last_torque_in = {}
    orig_ct = env._compute_torques

    def compute_torques(actions):
        last_torque_in["dof_pos"] = env.dof_pos.clone().numpy()
        last_torque_in["dof_vel"] = env.dof_vel.clone().numpy()
        return orig_ct(actions)

    env._compute_torques = compute_torques
    orig_pps = env.post_physics_step
    pre_post = {}
    plan = {}                                        # step -> what the stand-in physics did to a few robots

    def post_physics_step():
        t = plan["t"]
        r = env.root_states
        # walk robots around so that the height samples change and the curriculum sees every case:
        #  far from the origin (> env_length / 2) -> level up (wraps to a random level from the last one);
        #  close to it with a non-zero command -> level down (floor 0); zero command -> unchanged
        r[:, 0:2] += 0.35 * torch.randn(n_envs, 2, generator=g)
        if t % 4 == 1:
            far = torch.tensor([(3 * t) % n_envs, (3 * t + 7) % n_envs])
            r[far, 0] += 4.6
            env.episode_length_buf[far] = 2400       # they time out this step -> reset -> curriculum
            near = torch.tensor([(3 * t + 11) % n_envs])
            env.episode_length_buf[near] = 2400
        if t == 9:
            env.terrain_levels[:] = tc.num_rows - 1   # next promotions wrap around
        pre_post["pre"] = tsnap()
        orig_pps()

    env.post_physics_step = post_physics_step

    ag = torch.Generator().manual_seed(98)
    n_reset = 0
    for t in range(n_steps):
        rec.new_step()
        plan["t"] = t
        act_in = 2.0 * torch.randn(n_envs, 12, generator=ag)
        obs, priv, rew, reset, extras = env.step(act_in.clone())
        noise = rec.noise(terrain=True)
        pre, post = pre_post["pre"], tsnap()
        p = f"step{t:03d}."
        data[p + "actions_in"] = act_in.numpy()
        for k, v in noise.items():
            data[p + "noise." + k] = v
        data[p + "torque_in.dof_pos"] = last_torque_in["dof_pos"]
        data[p + "torque_in.dof_vel"] = last_torque_in["dof_vel"]
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state", "actions", "torques",
                  "episode_length_buf", "terrain_levels"):
            data[p + "pre." + k] = pre[k]
        for k in STATE_KEYS + ("last_feet_z", "episode_sums", "terrain_levels", "measured_heights"):
            if k in ("contact_forces", "rigid_state", "env_frictions", "body_mass"):
                continue
            data[p + "post." + k] = post[k]
        W = 705 + 187
        data[p + "post.obs_frame"] = obs[:, -47:].numpy().copy()
        data[p + "post.priv_frame"] = priv[:, -W:].numpy().copy()
        if t % 8 == 7 or t == n_steps - 1:
            data[p + "post.obs_buf"] = obs.numpy().copy()
            data[p + "post.privileged_obs_buf"] = priv.numpy().copy()
        data[p + "post.extras_time_outs"] = extras["time_outs"].numpy().copy()
        data[p + "post.episode_means"] = np.array(
            [float(extras["episode"]["rew_" + k]) for k in env.reward_names], np.float32)
        data[p + "post.extras_terrain_level"] = np.float32(float(extras["episode"]["terrain_level"]))
        n_reset += int(reset.sum())
    np.savez_compressed(out_path, **data)
    print(f"wrote {out_path}: {n_steps} steps x {n_envs} envs, {n_reset} resets,"
          f" height field {env.height_samples.shape}, {os.path.getsize(out_path) / 1e6:.2f} MB")


def make_cmd_curriculum_golden(out_path, n_envs=12, n_steps=8):
    """update_command_curriculum (legged_robot.py:178-180,422-431) inside chained env.step() calls."""
    from humanoid.envs import XBotLCfg, XBotLFreeEnv  # noqa: F401
    from humanoid.utils import task_registry
    uninstrument()

    class CurCfg(XBotLCfg):
        class commands(XBotLCfg.commands):
            curriculum, max_curriculum = True, 1.5

            class ranges(XBotLCfg.commands.ranges):
                lin_vel_x = [-0.3, 0.6]            # own list: the reference widens it in place

    args = argparse.Namespace(
        task="humanoid_ppo", resume=False, experiment_name=None, run_name=None, load_run=None,
        checkpoint=None, headless=True, horovod=False, rl_device="cpu", num_envs=n_envs, seed=5,
        max_iterations=None, physics_engine=1, use_gpu=False, use_gpu_pipeline=False, subscenes=0,
        num_threads=0, sim_device="cpu", sim_device_type="cpu", compute_device_id=0, sim_device_id=0,
        device="cpu")
    env_cfg = CurCfg()
    env_cfg.seed = 5
    env, cfg = task_registry.make_env(name="humanoid_ppo", args=args, env_cfg=env_cfg)
    rec = DrawRecorder(n_envs)
    instrument_env(env, rec)
    data = {"meta.n_envs": np.int64(n_envs), "meta.n_steps": np.int64(n_steps),
            "meta.max_curriculum": np.float64(env.cfg.commands.max_curriculum),
            "meta.init_range_x": np.array(env.command_ranges["lin_vel_x"], np.float64)}
    g = torch.Generator().manual_seed(79)
    env.episode_length_buf[:] = torch.randint(0, 2300, (n_envs,), generator=g)
    env.common_step_counter = 2398                   # step 1 lands on 2400
    k_track = env.reward_names.index("tracking_lin_vel")
    for k, v in snap(env).items():
        data[f"init.{k}"] = v

    last_torque_in, pre_post, plan = {}, {}, {}
    orig_ct, orig_pps = env._compute_torques, env.post_physics_step

    def compute_torques(actions):
        last_torque_in["dof_pos"] = env.dof_pos.clone().numpy()
        last_torque_in["dof_vel"] = env.dof_vel.clone().numpy()
        return orig_ct(actions)

    def post_physics_step():
        t = plan["t"]
        # t = 1, 3: counter -> 2400, 4800 with well-tracking envs timing out (range widens, the second time into the clip);
        # t = 5: counter -> 7200, the resetting envs tracked badly (unchanged); t = 6: a reset off the multiple (no check)
        if t in (1, 3, 5, 6):
            ids = torch.tensor([(2 * t) % n_envs, (2 * t + 5) % n_envs])
            env.episode_length_buf[ids] = 2400
            env.episode_sums["tracking_lin_vel"][:] = 3.0 if t == 5 else 40.0 + t
        if t == 3:
            env.common_step_counter = 4799
        if t == 5:
            env.common_step_counter = 7199
        pre_post["pre"] = snap(env)
        orig_pps()

    env._compute_torques = compute_torques
    env.post_physics_step = post_physics_step
    ag = torch.Generator().manual_seed(97)
    for t in range(n_steps):
        rec.new_step()
        plan["t"] = t
        act_in = 2.0 * torch.randn(n_envs, 12, generator=ag)
        obs, priv, rew, reset, extras = env.step(act_in.clone())
        noise, pre, post = rec.noise(), pre_post["pre"], snap(env)
        p = f"step{t:03d}."
        data[p + "actions_in"] = act_in.numpy()
        for k, v in noise.items():
            data[p + "noise." + k] = v
        data[p + "torque_in.dof_pos"], data[p + "torque_in.dof_vel"] = last_torque_in["dof_pos"], last_torque_in["dof_vel"]
        for k in ("root_states", "dof_pos", "dof_vel", "contact_forces", "rigid_state", "actions", "torques", "episode_length_buf",
                  "episode_sums", "common_step_counter"):
            data[p + "pre." + k] = pre[k]
        for k in STATE_KEYS + ("last_feet_z", "episode_sums"):
            if k in ("contact_forces", "rigid_state", "env_frictions", "body_mass", "env_origins"):
                continue
            data[p + "post." + k] = post[k]
        data[p + "post.obs_frame"] = obs[:, -47:].numpy().copy()
        data[p + "post.priv_frame"] = priv[:, -73:].numpy().copy()
        data[p + "post.extras_time_outs"] = extras["time_outs"].numpy().copy()
        data[p + "post.episode_means"] = np.array([float(extras["episode"]["rew_" + k]) for k in env.reward_names], np.float32)
        data[p + "post.range_x"] = np.array(env.command_ranges["lin_vel_x"], np.float64)
        data[p + "post.max_command_x"] = np.float64(extras["episode"]["max_command_x"])
        data[p + "post.common_step_counter"] = np.int64(env.common_step_counter)
    np.savez_compressed(out_path, **data)
    print(f"wrote {out_path}: {n_steps} steps x {n_envs} envs, ranges {[tuple(data[f'step{t:03d}.post.range_x']) for t in range(n_steps)]},"
          f" {os.path.getsize(out_path) / 1e6:.2f} MB")


# --------------------------------------------------------------------------
# learning-side goldens
# --------------------------------------------------------------------------
def make_ppo_golden(out_path):
    from humanoid.algo import ActorCritic, PPO, RolloutStorage

    torch.manual_seed(11)
    data = {}
    # A reduced architecture keeps the fixture small; the arithmetic being pinned (ELU MLP,
    # diag-Normal, GAE, clipped PPO loss, adaptive-KL lr, grad clip, Adam) is shape-independent.
    # The full 705-512-256-128-12 actor is pinned by policy_example_kat.npz.
    NA, NC = 60, 40
    ac = ActorCritic(NA, NC, 12, actor_hidden_dims=[32, 24, 16], critic_hidden_dims=[48, 24, 16],
                     init_noise_std=1.0)
    with torch.no_grad():
        ac.std.copy_(0.6 + 0.8 * torch.rand(12))      # non-trivial sigma
    for k, v in ac.state_dict().items():
        data["w0." + k] = v.numpy().copy()

    # --- A1/A2: forward, sample statistics, log-prob, entropy
    M = 96
    obs = torch.randn(M, NA).clamp(-18, 18)
    cobs = torch.randn(M, NC).clamp(-18, 18)
    acts = torch.randn(M, 12)
    with torch.no_grad():
        ac.update_distribution(obs)
        data["fwd.obs"], data["fwd.cobs"], data["fwd.actions"] = obs.numpy(), cobs.numpy(), acts.numpy()
        data["fwd.mean"] = ac.action_mean.numpy().copy()
        data["fwd.std"] = ac.action_std.numpy().copy()
        data["fwd.logp"] = ac.get_actions_log_prob(acts).numpy().copy()
        data["fwd.entropy"] = ac.entropy.numpy().copy()
        data["fwd.value"] = ac.evaluate(cobs).numpy().copy()

    # --- A5: GAE + advantage normalisation
    T, N = 24, 40
    st = RolloutStorage(N, T, [NA], [NC], [12], "cpu")
    st.rewards.copy_(torch.rand(T, N, 1))
    st.values.copy_(torch.randn(T, N, 1))
    st.dones.copy_((torch.rand(T, N, 1) < 0.08).byte())
    last_values = torch.randn(N, 1)
    st.compute_returns(last_values, 0.994, 0.9)
    for k in ("rewards", "values", "dones", "returns", "advantages"):
        data["gae." + k] = getattr(st, k).numpy().copy()
    data["gae.last_values"] = last_values.numpy()

    # --- A3/A6/A7: a full PPO.update() on a small storage
    T, N = 8, 48
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9,
              value_loss_coef=1.0, entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0,
              use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01, device="cpu")
    alg.init_storage(N, T, [NA], [NC], [12])
    alg.actor_critic.train()
    obs = torch.randn(N, NA).clamp(-18, 18)
    cobs = torch.randn(N, NC).clamp(-18, 18)
    with torch.inference_mode():
        for t in range(T):
            a = alg.act(obs, cobs)
            rew = torch.rand(N)
            dones = torch.rand(N) < 0.1
            infos = {"time_outs": dones & (torch.rand(N) < 0.5)}
            alg.process_env_step(rew, dones, infos)
            obs = (0.7 * obs + 0.6 * torch.randn(N, NA)).clamp(-18, 18)
            cobs = (0.7 * cobs + 0.6 * torch.randn(N, NC)).clamp(-18, 18)
        alg.compute_returns(cobs)
    s = alg.storage
    for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values",
              "returns", "advantages", "actions_log_prob", "mu", "sigma"):
        data["upd.storage." + k] = getattr(s, k).numpy().copy()
    data["upd.last_cobs"] = cobs.numpy().copy()

    perms, grads, lrs = [], [], []
    real_randperm = torch.randperm
    real_clip = torch.nn.utils.clip_grad_norm_

    def randperm(n, **kw):
        p = real_randperm(n, **kw)
        perms.append(p.numpy().copy())
        return p

    def clip_grad_norm_(params, max_norm, *a, **kw):
        params = list(params)
        grads.append(np.concatenate([p.grad.reshape(-1).numpy().copy() for p in params]))
        lrs.append(alg.learning_rate)
        return real_clip(params, max_norm, *a, **kw)

    torch.randperm = randperm
    torch.nn.utils.clip_grad_norm_ = clip_grad_norm_
    try:
        mv, ms = alg.update()
    finally:
        torch.randperm = real_randperm
        torch.nn.utils.clip_grad_norm_ = real_clip
    data["upd.perm"] = perms[0]
    data["upd.grads"] = np.stack(grads)                # (8, n_params) pre-clip gradients per optimizer step
    data["upd.lrs"] = np.array(lrs, np.float64)        # lr in force at each optimizer step
    data["upd.mean_value_loss"] = np.float64(mv)
    data["upd.mean_surrogate_loss"] = np.float64(ms)
    data["upd.final_lr"] = np.float64(alg.learning_rate)
    for k, v in ac.state_dict().items():
        data["w1." + k] = v.numpy().copy()
    data["param_order"] = np.array([k for k, _ in ac.named_parameters()])
    np.savez_compressed(out_path, **data)
    print(f"wrote {out_path}: {os.path.getsize(out_path) / 1e6:.2f} MB")


def make_policy_kat(out_path):
    path = os.path.join(REF, "logs/XBot_ppo/exported/policies/policy_example.pt")
    pol = torch.jit.load(path, map_location="cpu")
    data = {"w." + k: v.numpy().copy() for k, v in pol.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.cat((torch.zeros(1, 705), torch.randn(7, 705, generator=g).clamp(-18, 18)))
    with torch.no_grad():
        data["x"], data["y"] = x.numpy(), pol(x).numpy()
    np.savez_compressed(out_path, **data)
    print(f"wrote {out_path}: {os.path.getsize(out_path) / 1e6:.2f} MB; y[0]={np.round(data['y'][0], 4)}")


def make_cfg_golden(out_path):
    """class_to_dict() dumps of the reference's registered configs (the config tree IS the API)."""
    import json
    from humanoid.envs import XBotLCfg, XBotLCfgPPO
    from humanoid.utils.helpers import class_to_dict
    json.dump({"XBotLCfg": class_to_dict(XBotLCfg()), "XBotLCfgPPO": class_to_dict(XBotLCfgPPO())}, open(out_path, "w"),
              indent=1, sort_keys=True)
    print(f"wrote {out_path}")


if __name__ == "__main__":
    _install_shims()
    torch.set_num_threads(1)
    which = sys.argv[1:] or ["env", "ppo", "kat", "cfg", "terrain", "cmdcur"]
    if "env" in which:
        make_env_golden(os.path.join(HERE, "env_rollout.npz"))
    if "ppo" in which:
        make_ppo_golden(os.path.join(HERE, "ppo_learning.npz"))
    if "cfg" in which:
        make_cfg_golden(os.path.join(HERE, "cfg_dump.json"))
    if "kat" in which:
        make_policy_kat(os.path.join(HERE, "policy_example_kat.npz"))
    if "terrain" in which:
        make_terrain_golden(os.path.join(HERE, "env_terrain.npz"))
    if "cmdcur" in which:
        make_cmd_curriculum_golden(os.path.join(HERE, "env_cmd_curriculum.npz"))
