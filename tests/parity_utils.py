"""Shared helpers for the GPU parity tests and smoke(): build the product env, feed it a state, run the
oracle on the same state, compare."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import env_oracle as eo  # noqa: E402

# env attribute <- oracle/golden state key
_DIRECT = ("actions", "last_actions", "last_last_actions", "torques", "last_dof_vel", "last_root_vel", "commands",
           "base_lin_vel", "base_ang_vel", "projected_gravity", "base_euler_xyz", "feet_air_time", "last_contacts",
           "feet_height", "last_feet_z", "ref_dof_pos", "rand_push_force", "rand_push_torque", "env_frictions",
           "body_mass", "rew_buf", "env_origins", "reset_buf", "time_out_buf")


def make_args(num_envs, device="cuda:0"):
    return argparse.Namespace(
        task="humanoid_ppo", resume=False, experiment_name=None, run_name=None, load_run=None, checkpoint=None,
        headless=True, horovod=False, rl_device=device, num_envs=num_envs, seed=5, max_iterations=None,
        physics_engine=1, use_gpu=True, use_gpu_pipeline=True, subscenes=0, num_threads=0, sim_device=device,
        sim_device_type="cuda", compute_device_id=0, sim_device_id=0)


def make_env(num_envs, physics="external", device="cuda:0", cfg=None):
    from humanoid.envs import XBotLCfg, XBotLFreeEnv
    from humanoid.utils.helpers import class_to_dict, parse_sim_params
    cfg = XBotLCfg() if cfg is None else cfg
    cfg.env.num_envs = num_envs
    cfg.seed = 5
    cfg.physics_backend = physics
    args = make_args(num_envs, device)
    sim_params = parse_sim_params(args, {"sim": class_to_dict(cfg.sim)})
    return XBotLFreeEnv(cfg, sim_params, args.physics_engine, device, True)


def terrain_cfg_from_golden(g, num_envs=None):
    """XBotLCfg as tests/golden/make_golden.py::make_terrain_golden configured the reference (rough terrain, curriculum,
    measured heights in the critic frames)."""
    from humanoid.envs import XBotLCfg

    class TerrainCfg(XBotLCfg):
        class env(XBotLCfg.env):
            single_num_privileged_obs = XBotLCfg.env.num_observations + 17 * 11
            num_privileged_obs = int(XBotLCfg.env.c_frame_stack * single_num_privileged_obs)

        class terrain(XBotLCfg.terrain):
            mesh_type, curriculum, measure_heights = "trimesh", True, True
            num_rows, num_cols = int(g["meta.num_rows"]), int(g["meta.num_cols"])
            border_size, max_init_terrain_level = int(g["meta.border_size"]), int(g["meta.max_init_terrain_level"])
            terrain_proportions = [float(x) for x in g["meta.terrain_proportions"]]
    cfg = TerrainCfg()
    cfg.seed = 5
    if num_envs is not None:
        cfg.env.num_envs = num_envs
    return cfg


def terrain_params_from_golden(g):
    return eo.make_terrain_params(
        g["meta.height_samples"], g["meta.terrain_origins"], float(g["meta.border_size"]), float(g["meta.horizontal_scale"]),
        float(g["meta.vertical_scale"]), float(g["meta.env_length"]), curriculum=True, measure_heights=True,
        points_x=[float(x) for x in g["meta.measured_points_x"]], points_y=[float(x) for x in g["meta.measured_points_y"]],
        height_scale=float(g["meta.height_scale"]))


def random_state(N, g, p_base_contact=0.03, ep_max=2400):
    """A plausible, branch-rich env state in oracle layout (CPU tensors)."""
    def rn(*s):
        return torch.randn(*s, generator=g)

    def ru(*s):
        return torch.rand(*s, generator=g)
    S = eo.new_state(N, env_origins=eo.grid_origins(N), env_frictions=0.1 + 1.9 * ru(N, 1), body_mass=5 + 10 * ru(N, 1) - 5)
    rpy = 0.15 * rn(N, 3)
    rpy[:, 2] = (2 * ru(N) - 1) * np.pi
    cr, sr, cp, sp = torch.cos(rpy[:, 0] / 2), torch.sin(rpy[:, 0] / 2), torch.cos(rpy[:, 1] / 2), torch.sin(rpy[:, 1] / 2)
    cy, sy = torch.cos(rpy[:, 2] / 2), torch.sin(rpy[:, 2] / 2)
    r = S["root_states"]
    r[:, 0:2] = S["env_origins"][:, 0:2] + rn(N, 2)
    r[:, 2] = 0.95 + 0.03 * rn(N)
    r[:, 3], r[:, 4] = sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy
    r[:, 5], r[:, 6] = cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy
    r[:, 7:13] = 0.4 * rn(N, 6)
    S["dof_pos"], S["dof_vel"] = 0.25 * rn(N, 12), 1.5 * rn(N, 12)
    c = S["contact_forces"]
    in_contact = ru(N, 2) < 0.5
    for j, b in enumerate((6, 12)):
        c[:, b, 2] = (50 + 800 * ru(N)) * in_contact[:, j]
        c[:, b, 0:2] = 30 * rn(N, 2) * in_contact[:, j:j + 1]
    c[:, 0, :] = (ru(N) < p_base_contact).unsqueeze(1) * (1.5 + 4 * ru(N, 3))
    c[:, 0, :] += (ru(N) < 0.05).unsqueeze(1) * 0.15 * ru(N, 3)
    rg = S["rigid_state"]
    for j, b in enumerate((6, 12)):
        rg[:, b, 0] = r[:, 0] + 0.1 * rn(N)
        rg[:, b, 1] = r[:, 1] + (0.15 if j == 0 else -0.15) + 0.1 * rn(N)
        rg[:, b, 2] = 0.05 + 0.07 * ru(N)
        rg[:, b, 7:9] = 0.4 * rn(N, 2)
    for j, b in enumerate((4, 10)):
        rg[:, b, 0] = r[:, 0] + 0.05 * rn(N)
        rg[:, b, 1] = r[:, 1] + (0.12 if j == 0 else -0.12) + 0.06 * rn(N)
    for k in ("actions", "last_actions", "last_last_actions"):
        S[k] = 1.5 * rn(N, 12)
    S["torques"] = 30 * rn(N, 12)
    S["last_dof_vel"], S["last_root_vel"] = 1.5 * rn(N, 12), 0.4 * rn(N, 6)
    S["commands"] = torch.stack((-0.3 + 0.9 * ru(N), -0.3 + 0.6 * ru(N), 2 * ru(N) - 1, 3.14 * (2 * ru(N) - 1)), 1)
    S["commands"][ru(N) < 0.2, :2] = 0
    ep = torch.randint(0, ep_max, (N,), generator=g)
    ep[::97] = 2400                                   # time-outs
    ep[5::101] = 798                                  # command resampling after the increment
    S["episode_length_buf"] = ep
    S["feet_air_time"] = 0.6 * ru(N, 2) * (ru(N, 2) < 0.7)
    S["last_contacts"] = ru(N, 2) < 0.5
    S["feet_height"] = 0.08 * ru(N, 2)
    S["last_feet_z"] = 0.07 * ru(N, 2)
    S["ref_dof_pos"] = 0.17 * rn(N, 12) * (ru(N, 1) < 0.8)
    S["rand_push_force"][:, :2] = 0.2 * (2 * ru(N, 2) - 1)
    S["rand_push_torque"] = 0.4 * (2 * ru(N, 3) - 1)
    S["episode_sums"] = ru(22, N)
    S["obs_hist"] = rn(N, 15, 47).clamp(-18, 18)
    S["critic_hist"] = rn(N, 3, 73).clamp(-18, 18)
    S["obs_buf"] = S["obs_hist"].reshape(N, -1).clone()
    S["privileged_obs_buf"] = S["critic_hist"].reshape(N, -1).clone()
    S["base_lin_vel"], S["base_ang_vel"] = 0.3 * rn(N, 3), 0.3 * rn(N, 3)
    S["common_step_counter"] = 399                    # the step under test pushes (counter % 400 == 0)
    return S


def random_noise(N, g):
    return dict(u_cmd_cb=torch.rand(N, 3, generator=g), u_cmd_rs=torch.rand(N, 3, generator=g),
                u_dof=torch.rand(N, 12, generator=g), u_push=torch.rand(N, 5, generator=g),
                z_obs=torch.randn(N, 47, generator=g), u_delay=torch.rand(N, 1, generator=g),
                z_act=torch.randn(N, 12, generator=g))


def load_state(env, S):
    """Copy an oracle-layout state dict into the live tensors of the product env."""
    dev = env.device
    N = env.num_envs
    for k in _DIRECT:
        getattr(env, k).copy_(S[k].to(dev).reshape(getattr(env, k).shape))
    env.root_states.copy_(S["root_states"].to(dev))
    ds = env.dof_state.view(N, 12, 2)
    ds[..., 0] = S["dof_pos"].to(dev)
    ds[..., 1] = S["dof_vel"].to(dev)
    env.contact_forces.copy_(S["contact_forces"].to(dev))
    env.rigid_state.copy_(S["rigid_state"].to(dev))
    env.episode_length_buf = S["episode_length_buf"].to(dev)
    env._episode_sums.copy_(S["episode_sums"].to(dev))
    env.obs_buf.copy_(S["obs_hist"].reshape(N, -1).to(dev))
    env.privileged_obs_buf.copy_(S["critic_hist"].reshape(N, -1).to(dev))
    env.common_step_counter = int(S["common_step_counter"])
    if "episode_means" in S:
        env._episode_means.copy_(S["episode_means"].to(dev))
    if "extras_time_outs" in S:
        env.extras_time_outs.copy_(S["extras_time_outs"].to(dev))
    if "terrain_levels" in S and getattr(env, "terrain", None) is not None:
        env.terrain_levels.copy_(S["terrain_levels"].to(dev))


def oracle_step(S, noise, actions, physics_after=None):
    """Full env.step() on the oracle: E1/E2, E3 (last sub-step), post_physics.  Returns the mutated copy."""
    P = eo.make_params()
    S = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in S.items()}
    eo.pre_physics(S, P, actions, noise["u_delay"], noise["z_act"])
    eo.compute_torques(S, P)
    if physics_after is not None:
        for k, v in physics_after.items():
            S[k] = v.clone()
    eo.post_physics(S, P, noise)
    return S


CHECK_KEYS = ("root_states", "actions", "last_actions", "last_last_actions", "torques", "last_dof_vel", "last_root_vel",
              "commands", "episode_length_buf", "reset_buf", "time_out_buf", "base_lin_vel", "base_ang_vel",
              "projected_gravity", "base_euler_xyz", "feet_air_time", "last_contacts", "feet_height", "last_feet_z",
              "ref_dof_pos", "rand_push_force", "rand_push_torque", "rew_buf", "episode_sums", "obs_buf",
              "privileged_obs_buf", "dof_pos", "dof_vel", "extras_time_outs", "episode_means")
# fields that sit behind a threshold (sin>=0, |sin|<0.1, F>5, |h-0.06|<0.01, rew clip ...): a value within
# rounding distance of the threshold may legitimately land on the other side (SURVEY.md section 8c hazard 7)
THRESHOLDED = {"rew_buf", "episode_sums", "feet_air_time", "feet_height", "privileged_obs_buf", "ref_dof_pos",
               "episode_means", "obs_buf"}


def near_threshold_envs(S_pre, ref, noise=None, margin=1e-6, sin_margin=1e-7, wrap_margin=1e-5):
    """Envs whose step sits within rounding distance of one of the path's discontinuities (SURVEY.md section 8c hazard 7),
    computed from the oracle's inputs / outputs.  ONLY these envs may differ beyond tolerance in the THRESHOLDED
    quantities: a kernel that is off anywhere else fails.  Discontinuities covered:
      gait clock   sin(2 pi phase) at 0 (stance swap, ref-pose sign) and |sin| at 0.1 (double stance / zeroed ref pose);
                   (both sides evaluate the SAME fp32 argument, so only the last-ulp difference of sin itself matters -> sin_margin)
      clearance    | |feet_height - 0.06| - 0.01 |  (feet_clearance hit)
      rewards      the total at the >= 0 clip
      contacts     |F_base| at 1.0 (termination) and 0.1 (collision)
      low_speed    |v_x| vs 0.5 / 1.2 |c_x|, sign(v_x), |c_x| vs 0.1     (envs not reset this step)
      heading      wrap_to_pi of (heading command - heading) at +-pi, Euler angles at +-pi
      commands     |c_xy| at 0.2 on resampling
    """
    P = eo.make_params()
    N = S_pre["root_states"].shape[0]
    near = torch.zeros(N, dtype=torch.bool)
    ep1 = S_pre["episode_length_buf"] + 1
    phase = ep1 * P["dt"] / P["cycle_time"]
    s = torch.sin(2 * torch.pi * phase)
    near |= s.abs() < sin_margin
    near |= (s.abs() - 0.1).abs() < sin_margin
    fz = S_pre["rigid_state"][:, list(P["feet"]), 2] - 0.05
    h = S_pre["feet_height"] + (fz - S_pre["last_feet_z"])
    near |= (((h - P["target_feet_height"]).abs() - 0.01).abs() < margin).any(1)
    if "rew_terms" in ref:
        near |= ref["rew_terms"].sum(0).abs() < margin
    fb = torch.norm(S_pre["contact_forces"][:, 0, :], dim=-1)
    near |= ((fb - 1.0).abs() < margin * 10) | ((fb - 0.1).abs() < margin)
    reset = ref["reset_buf"].bool()
    v, c = ref["base_lin_vel"][:, 0], ref["commands"][:, 0]
    av, ac = v.abs(), c.abs()
    ls = ((av - 0.5 * ac).abs() < margin) | ((av - 1.2 * ac).abs() < margin) | (av < margin * 0.1) | ((ac - 0.1).abs() < margin)
    near |= ls & ~reset
    q = S_pre["root_states"][:, 3:7]
    fwd = eo.quat_apply(q, torch.tensor([1., 0., 0.]).repeat(N, 1))
    heading = torch.atan2(fwd[:, 1], fwd[:, 0])
    d = (ref["commands"][:, 3] - heading) % (2 * torch.pi)
    near |= ((d - torch.pi).abs() < wrap_margin) & ~reset
    near |= ((ref["base_euler_xyz"].abs() - torch.pi).abs() < wrap_margin).any(1)
    if noise is not None:
        for u, mask in ((noise["u_cmd_cb"], (ep1 % P["resample_period"] == 0)), (noise["u_cmd_rs"], reset)):
            cx = eo._uniform(*P["cmd_x"], u[:, 0])
            cy = eo._uniform(*P["cmd_y"], u[:, 1])
            near |= ((torch.sqrt(cx * cx + cy * cy) - 0.2).abs() < margin) & mask
    return near


def env_value(env, k):
    N = env.num_envs
    if k == "dof_pos":
        return env.dof_state.view(N, 12, 2)[..., 0]
    if k == "dof_vel":
        return env.dof_state.view(N, 12, 2)[..., 1]
    if k == "episode_sums":
        return env._episode_sums
    if k == "episode_means":
        return env._episode_means
    return getattr(env, k)


def compare_step(env, ref, rtol=1e-5, atol=1e-6, max_outlier_frac=0.0, keys=CHECK_KEYS, near=None):
    """Returns a list of human-readable mismatches (empty == parity).

    near: optional (N,) bool mask from near_threshold_envs(): out-of-tolerance elements of THRESHOLDED quantities are
    accepted only in those envs (and `max_outlier_frac` is then ignored); without it the legacy fractional allowance
    applies (used by smoke() only)."""
    bad = []
    N = env.num_envs
    for k in keys:
        got = env_value(env, k).detach().cpu()
        want = ref[k]
        if got.dtype == torch.bool or want.dtype == torch.bool:
            n = int((got.bool().reshape(want.shape) != want.bool()).sum())
            if n:
                bad.append(f"{k}: {n} boolean mismatches")
            continue
        got = got.reshape(want.shape).double()
        want = want.double()
        err = (got - want).abs()
        tol = atol + rtol * want.abs()
        viol = err > tol
        n = int(viol.sum())
        allowed = int(max_outlier_frac * want.numel()) if k in THRESHOLDED else 0
        if near is not None:
            allowed = 0
            if k in THRESHOLDED and n:
                if k == "episode_means":          # a mean over the reset envs: tainted if any of them is near a threshold
                    ok = bool((near & ref["reset_buf"].bool()).any())
                    n = 0 if ok else n
                else:
                    v_env = viol.reshape(22, N).any(0) if k == "episode_sums" else viol.reshape(N, -1).any(1)
                    n = int((v_env & ~near).sum())
        if n > allowed:
            i = int(torch.argmax((err - tol).flatten()))
            bad.append(f"{k}: {n}/{want.numel()} out of tolerance (allowed {allowed}); worst got={got.flatten()[i].item():.8g} "
                       f"want={want.flatten()[i].item():.8g}")
    return bad
