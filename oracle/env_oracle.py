"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the
reference's vectorised XBot-L environment step, in plain fp32 torch ops.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path never routes
through it.

Pinning: tests/test_oracle_env.py checks every function here against golden
vectors captured from the UNMODIFIED reference (tests/golden/make_golden.py
runs /root/reference over a test-only fake isaacgym).  The Isaac Gym math
helpers (quat_rotate_inverse, quat_apply, get_euler_xyz, torch_rand_float)
are third-party, not vendored in /root/reference and not pinned by any
reference test: they follow Isaac Gym's public definitions -> that part is
"parity unpinned" (DESIGN.md section 3).

All `file:line` citations are relative to /root/reference/humanoid/.
State lives in a plain dict of tensors `S`; constants in the dict `P`
(make_params).  Random draws are INJECTED as dense per-env tensors (uniform
[0,1) or standard normal) because the reference's draws have data-dependent
shapes (SURVEY.md section 8c hazard 12).
"""
import math

import numpy as np
import torch

REWARD_NAMES = (  # alphabetical == dir() order, utils/helpers.py:44-59, envs/base/legged_robot.py:518-541
    "action_smoothness", "base_acc", "base_height", "collision", "default_joint_pos", "dof_acc",
    "dof_vel", "feet_air_time", "feet_clearance", "feet_contact_forces", "feet_contact_number",
    "feet_distance", "foot_slip", "joint_pos", "knee_distance", "low_speed", "orientation",
    "torques", "track_vel_hard", "tracking_ang_vel", "tracking_lin_vel", "vel_mismatch_exp")

REWARD_SCALES = dict(  # envs/custom/humanoid_config.py:188-216
    joint_pos=1.6, feet_clearance=1.0, feet_contact_number=1.2, feet_air_time=1.0, foot_slip=-0.05,
    feet_distance=0.2, knee_distance=0.2, feet_contact_forces=-0.01, tracking_lin_vel=1.2,
    tracking_ang_vel=1.1, vel_mismatch_exp=0.5, low_speed=0.2, track_vel_hard=0.5,
    default_joint_pos=0.5, orientation=1.0, base_height=0.2, base_acc=0.2, action_smoothness=-0.002,
    torques=-1e-5, dof_vel=-5e-4, dof_acc=-1e-7, collision=-1.0)

N_DOF, N_BODY, N_OBS1, N_PRIV1, N_FRAMES, N_CFRAMES = 12, 13, 47, 73, 15, 3


def make_params(sim_dt=0.001, decimation=10):
    """Constants of XBotLCfg (envs/custom/humanoid_config.py) as the reference derives them.

    `sim_dt` goes through a C float inside gymapi.SimParams, hence the float32
    round trip (SURVEY.md section 8c hazard 3): dt = 0.010000000475, command
    resample period int(8/dt) = 799."""
    dt = decimation * float(np.float32(sim_dt))          # legged_robot.py:711
    P = dict(
        dt=dt,
        cycle_time=0.64, target_joint_pos_scale=0.17, target_feet_height=0.06,
        base_height_target=0.89, min_dist=0.2, max_dist=0.5, max_contact_force=700.0,
        tracking_sigma=5.0, only_positive_rewards=True,
        max_episode_length=float(np.ceil(24.0 / dt)),        # legged_robot.py:718
        max_episode_length_s=24.0,
        resample_period=int(8.0 / dt),                      # legged_robot.py:309
        push_interval=float(np.ceil(4.0 / dt)),             # legged_robot.py:720
        max_push_vel_xy=0.2, max_push_ang_vel=0.4,
        action_delay=0.5, action_noise=0.02,
        clip_actions=18.0, clip_obs=18.0, action_scale=0.25,
        cmd_x=(-0.3, 0.6), cmd_y=(-0.3, 0.3), cmd_heading=(-3.14, 3.14),
        obs_scale_lin_vel=2.0, obs_scale_ang_vel=1.0, obs_scale_dof_pos=1.0,
        obs_scale_dof_vel=0.05, obs_scale_quat=1.0,
        noise_level=0.6, add_noise=True,
        feet=(6, 12), knees=(4, 10), term_bodies=(0,), pen_bodies=(0,),
        base_init_state=(0.0, 0.0, 0.95, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0),
    )
    # reward scale * dt is formed in double, then applied to fp32 tensors (legged_robot.py:528)
    P["reward_scales"] = [REWARD_SCALES[k] * dt for k in REWARD_NAMES]
    # PD gains by joint-name match (legged_robot.py:487-501, humanoid_config.py:120-123)
    P["p_gains"] = torch.tensor([200., 200., 350., 350., 15., 15.] * 2)
    P["d_gains"] = torch.full((12,), 10.0)
    # 0.85 * URDF effort (legged_robot.py:293, humanoid_config.py:55)
    P["torque_limits"] = torch.tensor([100., 100., 250., 250., 100., 100.] * 2) * 0.85
    P["default_dof_pos"] = torch.zeros(12)
    nv = torch.zeros(N_OBS1)                               # humanoid_env.py:166-186
    nv[5:17] = 0.05 * 1.0
    nv[17:29] = 0.5 * 0.05
    nv[41:44] = 0.1 * 1.0
    nv[44:47] = 0.03 * 1.0
    P["noise_scale_vec"] = nv
    return P


# ----------------------------------------------------------------------------
# Isaac Gym torch_utils, public definitions (xyzw).  Parity unpinned.
# ----------------------------------------------------------------------------
def quat_rotate_inverse(q, v):
    w = q[:, 3]
    u = q[:, :3]
    a = v * (2.0 * w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(u, v, dim=-1) * w.unsqueeze(-1) * 2.0
    c = u * (u * v).sum(-1, keepdim=True) * 2.0
    return a - b + c


def quat_apply(q, v):
    u = q[:, :3]
    t = torch.cross(u, v, dim=-1) * 2
    return v + q[:, 3:] * t + torch.cross(u, t, dim=-1)


def euler_xyz_wrapped(q):
    """get_euler_xyz followed by the (pi, 2pi) -> negative fold, legged_robot.py:50-55."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2.0 * (w * x + y * z), w * w - x * x - y * y + z * z)
    sinp = 2.0 * (w * y - z * x)
    half_pi = torch.full_like(sinp, np.pi / 2.0)
    pitch = torch.where(sinp.abs() >= 1, half_pi.abs() * torch.sign(sinp), torch.asin(sinp))
    yaw = torch.atan2(2.0 * (w * z + x * y), w * w + x * x - y * y - z * z)
    e = torch.stack((roll % (2 * np.pi), pitch % (2 * np.pi), yaw % (2 * np.pi)), dim=1)
    e[e > np.pi] -= 2 * np.pi
    return e


def wrap_to_pi(a):                                          # utils/math.py:47-50
    a = a % (2 * np.pi)
    return a - 2 * np.pi * (a > np.pi)


def _uniform(lo, hi, u):                                    # isaacgym torch_rand_float
    return (hi - lo) * u + lo


def quat_apply_yaw(q, v):                                   # utils/math.py:38-43 (isaacgym normalize: x / |x|.clamp(min=1e-9))
    qy = q.clone().view(-1, 4)
    qy[:, :2] = 0.
    qy = qy / qy.norm(p=2, dim=-1).clamp(min=1e-9).unsqueeze(-1)
    return quat_apply(qy, v.reshape(-1, 3)).view(v.shape)       # isaacgym quat_apply flattens to (-1, 4) / (-1, 3)


# ----------------------------------------------------------------------------
# rough terrain: P["terrain"] is None on a plane, else a dict
# ----------------------------------------------------------------------------
def make_terrain_params(height_samples, terrain_origins, border_size, horizontal_scale, vertical_scale, env_length,
                        curriculum, measure_heights, points_x=None, points_y=None, height_scale=5.0):
    """height_samples: int16 (rows, cols) = Terrain.heightsamples; terrain_origins: (levels, types, 3) fp32
    (legged_robot.py:570,586,695-696); height points as _init_height_points builds them (:743-757)."""
    T = dict(height_samples=torch.as_tensor(height_samples), terrain_origins=torch.as_tensor(terrain_origins).float(),
             border_size=border_size, horizontal_scale=horizontal_scale, vertical_scale=vertical_scale,
             env_length=env_length, max_terrain_level=int(terrain_origins.shape[0]), curriculum=bool(curriculum),
             measure_heights=bool(measure_heights), height_scale=height_scale)
    if measure_heights:
        gx, gy = torch.meshgrid(torch.tensor(points_x), torch.tensor(points_y), indexing="ij")
        pts = torch.zeros(gx.numel(), 3)
        pts[:, 0], pts[:, 1] = gx.flatten(), gy.flatten()
        T["height_points"] = pts
    return T


def get_heights(S, P):                                      # legged_robot.py:759-795
    T = P["terrain"]
    N = S["root_states"].shape[0]
    pts = T["height_points"].unsqueeze(0).repeat(N, 1, 1)
    npts = pts.shape[1]
    q = S["root_states"][:, 3:7]
    points = quat_apply_yaw(q.repeat(1, npts), pts) + (S["root_states"][:, :3]).unsqueeze(1)
    points += T["border_size"]
    points = (points / T["horizontal_scale"]).long()
    hs = T["height_samples"]
    px = torch.clip(points[:, :, 0].view(-1), 0, hs.shape[0] - 2)
    py = torch.clip(points[:, :, 1].view(-1), 0, hs.shape[1] - 2)
    h = torch.min(torch.min(hs[px, py], hs[px + 1, py]), hs[px, py + 1])
    return h.view(N, -1) * T["vertical_scale"]


def update_terrain_curriculum(S, P, ids, r_level):          # legged_robot.py:400-420
    """r_level (N,) int64: the torch.randint_like draw, densified per env."""
    T = P["terrain"]
    distance = torch.norm(S["root_states"][ids, :2] - S["env_origins"][ids, :2], dim=1)
    move_up = distance > T["env_length"] / 2
    move_down = (distance < torch.norm(S["commands"][ids, :2], dim=1) * P["max_episode_length_s"] * 0.5) * ~move_up
    S["terrain_levels"][ids] += 1 * move_up - 1 * move_down
    S["terrain_levels"][ids] = torch.where(S["terrain_levels"][ids] >= T["max_terrain_level"], r_level[ids],
                                           torch.clip(S["terrain_levels"][ids], 0))
    S["env_origins"][ids] = T["terrain_origins"][S["terrain_levels"][ids], S["terrain_types"][ids]]


# ----------------------------------------------------------------------------
# state
# ----------------------------------------------------------------------------
def new_state(N, env_origins=None, env_frictions=None, body_mass=None):
    """Buffers as allocated by base_task.py:62-94, legged_robot.py:434-516, humanoid_env.py:78-79."""
    f = torch.zeros
    S = dict(
        root_states=f(N, 13), dof_pos=f(N, 12), dof_vel=f(N, 12),
        contact_forces=f(N, 13, 3), rigid_state=f(N, 13, 13),
        actions=f(N, 12), last_actions=f(N, 12), last_last_actions=f(N, 12), torques=f(N, 12),
        last_dof_vel=f(N, 12), last_root_vel=f(N, 6), commands=f(N, 4),
        episode_length_buf=f(N, dtype=torch.long), reset_buf=torch.ones(N, dtype=torch.bool),
        time_out_buf=f(N, dtype=torch.bool), extras_time_outs=f(N, dtype=torch.bool),
        base_lin_vel=f(N, 3), base_ang_vel=f(N, 3), projected_gravity=f(N, 3), base_euler_xyz=f(N, 3),
        feet_air_time=f(N, 2), last_contacts=f(N, 2, dtype=torch.bool),
        feet_height=f(N, 2), last_feet_z=torch.full((N, 2), 0.05),
        ref_dof_pos=f(N, 12), rand_push_force=f(N, 3), rand_push_torque=f(N, 3),
        env_frictions=f(N, 1) if env_frictions is None else env_frictions.clone(),
        body_mass=f(N, 1) if body_mass is None else body_mass.clone(),
        episode_sums=f(len(REWARD_NAMES), N), episode_means=f(len(REWARD_NAMES)),
        obs_hist=f(N, N_FRAMES, N_OBS1), critic_hist=f(N, N_CFRAMES, N_PRIV1),
        obs_buf=f(N, N_FRAMES * N_OBS1), privileged_obs_buf=f(N, N_CFRAMES * N_PRIV1), rew_buf=f(N),
        env_origins=f(N, 3) if env_origins is None else env_origins.clone(),
        common_step_counter=0,
    )
    S["root_states"][:, 6] = 1.0
    return S


def grid_origins(N, spacing=3.0):                            # legged_robot.py:698-708
    cols = np.floor(np.sqrt(N))
    rows = np.ceil(N / cols)
    xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
    o = torch.zeros(N, 3)
    o[:, 0] = spacing * xx.flatten()[:N]
    o[:, 1] = spacing * yy.flatten()[:N]
    return o


# ----------------------------------------------------------------------------
# E1 + E2 (clip) + E3
# ----------------------------------------------------------------------------
def pre_physics(S, P, actions, u_delay, z_act):
    """humanoid_env.py:189-197 then legged_robot.py:90-91.  u_delay (N,1) U[0,1), z_act (N,12) N(0,1)."""
    c = P["clip_actions"]
    a = torch.clip(actions, -c, c)
    delay = u_delay * P["action_delay"]
    a = (1 - delay) * a + delay * S["actions"]
    a = a + P["action_noise"] * z_act * a
    S["actions"] = torch.clip(a, -c, c)
    return S["actions"]


def compute_torques(S, P):
    """legged_robot.py:340-356 (PD on position targets)."""
    scaled = S["actions"] * P["action_scale"]
    t = P["p_gains"] * (scaled + P["default_dof_pos"] - S["dof_pos"]) - P["d_gains"] * S["dof_vel"]
    S["torques"] = torch.clip(t, -P["torque_limits"], P["torque_limits"])
    return S["torques"]


# ----------------------------------------------------------------------------
# gait clock
# ----------------------------------------------------------------------------
def _sin_phase(S, P):
    phase = S["episode_length_buf"] * P["dt"] / P["cycle_time"]     # humanoid_env.py:100-103
    return phase, torch.sin(2 * torch.pi * phase)


def _stance_mask(S, P):                                           # humanoid_env.py:105-118
    _, s = _sin_phase(S, P)
    m = torch.zeros(s.shape[0], 2)
    m[:, 0] = s >= 0
    m[:, 1] = s < 0
    m[torch.abs(s) < 0.1] = 1
    return m


def _feet_contact(S, P):
    return S["contact_forces"][:, list(P["feet"]), 2] > 5.0


# ----------------------------------------------------------------------------
# E7.k -- the 22 reward terms (humanoid_env.py:272-540), each returns (N,)
# ----------------------------------------------------------------------------
def _two_point_distance_reward(p, lo, hi):
    d = torch.norm(p[:, 0, :] - p[:, 1, :], dim=1)
    d_min = torch.clamp(d - lo, -0.5, 0.0)
    d_max = torch.clamp(d - hi, 0, 0.5)
    return (torch.exp(-torch.abs(d_min) * 100) + torch.exp(-torch.abs(d_max) * 100)) / 2


def r_action_smoothness(S, P):                                    # :530-540
    t1 = torch.sum(torch.square(S["last_actions"] - S["actions"]), dim=1)
    t2 = torch.sum(torch.square(S["actions"] + S["last_last_actions"] - 2 * S["last_actions"]), dim=1)
    t3 = 0.05 * torch.sum(torch.abs(S["actions"]), dim=1)
    return t1 + t2 + t3


def r_base_acc(S, P):                                             # :386-393
    return torch.exp(-torch.norm(S["last_root_vel"] - S["root_states"][:, 7:13], dim=1) * 3)


def r_base_height(S, P):                                          # :374-384
    st = _stance_mask(S, P)
    fz = S["rigid_state"][:, list(P["feet"]), 2]
    measured = torch.sum(fz * st, dim=1) / torch.sum(st, dim=1)
    h = S["root_states"][:, 2] - (measured - 0.05)
    return torch.exp(-torch.abs(h - P["base_height_target"]) * 100)


def r_collision(S, P):                                            # :523-528
    f = S["contact_forces"][:, list(P["pen_bodies"]), :]
    return torch.sum(1. * (torch.norm(f, dim=-1) > 0.1), dim=1)


def r_default_joint_pos(S, P):                                    # :362-372
    d = S["dof_pos"] - P["default_dof_pos"]
    yr = torch.norm(d[:, :2], dim=1) + torch.norm(d[:, 6:8], dim=1)
    yr = torch.clamp(yr - 0.1, 0, 50)
    return torch.exp(-yr * 100) - 0.01 * torch.norm(d, dim=1)


def r_dof_acc(S, P):                                              # :516-521
    return torch.sum(torch.square((S["last_dof_vel"] - S["dof_vel"]) / P["dt"]), dim=1)


def r_dof_vel(S, P):                                              # :509-514
    return torch.sum(torch.square(S["dof_vel"]), dim=1)


def r_feet_air_time(S, P):                                        # :320-334  (stateful)
    contact = _feet_contact(S, P)
    st = _stance_mask(S, P)
    filt = torch.logical_or(torch.logical_or(contact, st), S["last_contacts"])
    S["last_contacts"] = contact
    first = (S["feet_air_time"] > 0.) * filt
    S["feet_air_time"] = S["feet_air_time"] + P["dt"]
    air = S["feet_air_time"].clamp(0, 0.5) * first
    S["feet_air_time"] = S["feet_air_time"] * ~filt
    return air.sum(dim=1)


def r_feet_clearance(S, P):                                       # :446-467  (stateful)
    contact = _feet_contact(S, P)
    fz = S["rigid_state"][:, list(P["feet"]), 2] - 0.05
    S["feet_height"] = S["feet_height"] + (fz - S["last_feet_z"])
    S["last_feet_z"] = fz
    swing = 1 - _stance_mask(S, P)
    hit = torch.abs(S["feet_height"] - P["target_feet_height"]) < 0.01
    r = torch.sum(hit * swing, dim=1)
    S["feet_height"] = S["feet_height"] * ~contact
    return r


def r_feet_contact_forces(S, P):                                  # :355-360
    n = torch.norm(S["contact_forces"][:, list(P["feet"]), :], dim=-1)
    return torch.sum((n - P["max_contact_force"]).clip(0, 400), dim=1)


def r_feet_contact_number(S, P):                                  # :336-344
    r = torch.where(_feet_contact(S, P) == _stance_mask(S, P), 1.0, -0.3)
    return torch.mean(r, dim=1)


def r_feet_distance(S, P):                                        # :282-292
    return _two_point_distance_reward(S["rigid_state"][:, list(P["feet"]), :2], P["min_dist"], P["max_dist"])


def r_foot_slip(S, P):                                            # :308-318
    sp = torch.sqrt(torch.norm(S["rigid_state"][:, list(P["feet"]), 7:9], dim=2))
    return torch.sum(sp * _feet_contact(S, P), dim=1)


def r_joint_pos(S, P):                                            # :272-280 (uses the STALE ref_dof_pos)
    e = torch.norm(S["dof_pos"] - S["ref_dof_pos"], dim=1)
    return torch.exp(-2 * e) - 0.2 * e.clamp(0, 0.5)


def r_knee_distance(S, P):                                        # :295-305
    return _two_point_distance_reward(S["rigid_state"][:, list(P["knees"]), :2], P["min_dist"], P["max_dist"] / 2)


def r_low_speed(S, P):                                            # :469-500
    v, c = S["base_lin_vel"][:, 0], S["commands"][:, 0]
    av, ac = torch.abs(v), torch.abs(c)
    low = av < 0.5 * ac
    high = av > 1.2 * ac
    r = torch.zeros_like(v)
    r[low] = -1.0
    r[high] = 0.
    r[~(low | high)] = 1.2
    r[torch.sign(v) != torch.sign(c)] = -2.0
    return r * (ac > 0.1)


def r_orientation(S, P):                                          # :346-353
    a = torch.exp(-torch.sum(torch.abs(S["base_euler_xyz"][:, :2]), dim=1) * 10)
    b = torch.exp(-torch.norm(S["projected_gravity"][:, :2], dim=1) * 20)
    return (a + b) / 2.


def r_torques(S, P):                                              # :502-507
    return torch.sum(torch.square(S["torques"]), dim=1)


def r_track_vel_hard(S, P):                                       # :408-425
    le = torch.norm(S["commands"][:, :2] - S["base_lin_vel"][:, :2], dim=1)
    ae = torch.abs(S["commands"][:, 2] - S["base_ang_vel"][:, 2])
    return (torch.exp(-le * 10) + torch.exp(-ae * 10)) / 2. - 0.2 * (le + ae)


def r_tracking_ang_vel(S, P):                                     # :436-444
    return torch.exp(-torch.square(S["commands"][:, 2] - S["base_ang_vel"][:, 2]) * P["tracking_sigma"])


def r_tracking_lin_vel(S, P):                                     # :427-434
    e = torch.sum(torch.square(S["commands"][:, :2] - S["base_lin_vel"][:, :2]), dim=1)
    return torch.exp(-e * P["tracking_sigma"])


def r_vel_mismatch_exp(S, P):                                     # :396-406
    a = torch.exp(-torch.square(S["base_lin_vel"][:, 2]) * 10)
    b = torch.exp(-torch.norm(S["base_ang_vel"][:, :2], dim=1) * 5.)
    return (a + b) / 2.


REWARD_FNS = tuple(globals()["r_" + k] for k in REWARD_NAMES)


# ----------------------------------------------------------------------------
# E5 / E6 / E7 / E8 / E9
# ----------------------------------------------------------------------------
def resample_commands(S, P, ids, u):
    """legged_robot.py:322-336.  u: dense (N,3) U[0,1) for (x, y, heading)."""
    if len(ids) == 0:
        return
    c = S["commands"]
    c[ids, 0] = _uniform(*P["cmd_x"], u[ids, 0])
    c[ids, 1] = _uniform(*P["cmd_y"], u[ids, 1])
    c[ids, 3] = _uniform(*P["cmd_heading"], u[ids, 2])
    c[ids, :2] *= (torch.norm(c[ids, :2], dim=1) > 0.2).unsqueeze(1)


def step_callback(S, P, noise):
    """legged_robot.py:304-320 + humanoid_env.py:83-98."""
    ids = (S["episode_length_buf"] % P["resample_period"] == 0).nonzero(as_tuple=False).flatten()
    resample_commands(S, P, ids, noise["u_cmd_cb"])
    q = S["root_states"][:, 3:7]
    fwd = quat_apply(q, torch.tensor([1., 0., 0.]).repeat(q.shape[0], 1))
    heading = torch.atan2(fwd[:, 1], fwd[:, 0])
    S["commands"][:, 2] = torch.clip(0.5 * wrap_to_pi(S["commands"][:, 3] - heading), -1., 1.)
    T = P.get("terrain")
    if T is not None and T["measure_heights"]:                  # :316-317
        S["measured_heights"] = get_heights(S, P)
    if S["common_step_counter"] % P["push_interval"] == 0:
        u = noise["u_push"]
        mv, ma = P["max_push_vel_xy"], P["max_push_ang_vel"]
        S["rand_push_force"][:, :2] = _uniform(-mv, mv, u[:, 0:2])
        S["root_states"][:, 7:9] = S["rand_push_force"][:, :2]
        S["rand_push_torque"] = _uniform(-ma, ma, u[:, 2:5])
        S["root_states"][:, 10:13] = S["rand_push_torque"]


def check_termination(S, P):                                      # legged_robot.py:156-161
    f = S["contact_forces"][:, list(P["term_bodies"]), :]
    S["reset_buf"] = torch.any(torch.norm(f, dim=-1) > 1., dim=1)
    S["time_out_buf"] = S["episode_length_buf"] > P["max_episode_length"]
    S["reset_buf"] = S["reset_buf"] | S["time_out_buf"]


def compute_reward(S, P):                                         # legged_robot.py:217-235
    S["rew_buf"] = torch.zeros_like(S["rew_buf"])
    S["rew_terms"] = torch.zeros_like(S["episode_sums"])
    for k, fn in enumerate(REWARD_FNS):
        r = fn(S, P) * P["reward_scales"][k]
        S["rew_buf"] = S["rew_buf"] + r
        S["episode_sums"][k] += r
        S["rew_terms"][k] = r
    if P["only_positive_rewards"]:
        S["rew_buf"] = torch.clip(S["rew_buf"], min=0.)


def update_command_curriculum(S, P, ids):                       # legged_robot.py:422-431
    """Widens P["cmd_x"] in place (the reference mutates self.command_ranges)."""
    k = REWARD_NAMES.index("tracking_lin_vel")
    if torch.mean(S["episode_sums"][k][ids]) / P["max_episode_length"] > 0.8 * P["reward_scales"][k]:
        lo, hi = P["cmd_x"]
        m = P["max_curriculum"]
        P["cmd_x"] = (np.clip(lo - 0.5, -m, 0.), np.clip(hi + 0.5, 0., m))


def reset_idx(S, P, ids, noise):
    """legged_robot.py:163-215 + :359-397 + humanoid_env.py:264-269."""
    if len(ids) == 0:
        return
    T = P.get("terrain")
    if T is not None and T["curriculum"] and S.get("init_done", True):       # :175-177, :407-409
        update_terrain_curriculum(S, P, ids, noise["r_level"])
    if P.get("cmd_curriculum") and S["common_step_counter"] % P["max_episode_length"] == 0:      # :178-180
        update_command_curriculum(S, P, ids)
    S["dof_pos"][ids] = P["default_dof_pos"] + _uniform(-0.1, 0.1, noise["u_dof"][ids])
    S["dof_vel"][ids] = 0.
    S["root_states"][ids] = torch.tensor(P["base_init_state"])
    S["root_states"][ids, :3] += S["env_origins"][ids]
    if T is not None:                                                         # custom origins :381-384
        S["root_states"][ids, :2] += _uniform(-1., 1., noise["u_root"][ids])
    resample_commands(S, P, ids, noise["u_cmd_rs"])
    for k in ("last_last_actions", "actions", "last_actions", "last_dof_vel", "feet_air_time"):
        S[k][ids] = 0.
    S["episode_length_buf"][ids] = 0
    S["reset_buf"][ids] = True
    for k in range(len(REWARD_NAMES)):
        S["episode_means"][k] = torch.mean(S["episode_sums"][k][ids]) / P["max_episode_length_s"]
        S["episode_sums"][k][ids] = 0.
    S["extras_time_outs"] = S["time_out_buf"].clone()           # only refreshed when something reset
    q = S["root_states"][:, 3:7]
    S["base_euler_xyz"] = euler_xyz_wrapped(q)
    g = torch.tensor([0., 0., -1.]).repeat(len(ids), 1)
    S["projected_gravity"][ids] = quat_rotate_inverse(q[ids], g)
    S["obs_hist"][ids] *= 0
    S["critic_hist"][ids] *= 0


def compute_observations(S, P, z_obs):
    """humanoid_env.py:200-262 incl. compute_ref_state :121-142.  z_obs (N,47) N(0,1)."""
    phase, s = _sin_phase(S, P)
    sl, sr = s.clone(), s.clone()
    ref = torch.zeros_like(S["dof_pos"])
    k1 = P["target_joint_pos_scale"]
    k2 = 2 * k1
    sl[sl > 0] = 0
    ref[:, 2], ref[:, 3], ref[:, 4] = sl * k1, sl * k2, sl * k1
    sr[sr < 0] = 0
    ref[:, 8], ref[:, 9], ref[:, 10] = sr * k1, sr * k2, sr * k1
    ref[torch.abs(s) < 0.1] = 0
    S["ref_dof_pos"] = ref

    sin_pos = torch.sin(2 * torch.pi * phase).unsqueeze(1)
    cos_pos = torch.cos(2 * torch.pi * phase).unsqueeze(1)
    stance = _stance_mask(S, P)
    contact = _feet_contact(S, P)
    cmd_scale = torch.tensor([P["obs_scale_lin_vel"], P["obs_scale_lin_vel"], P["obs_scale_ang_vel"]])
    cmd_in = torch.cat((sin_pos, cos_pos, S["commands"][:, :3] * cmd_scale), dim=1)
    q = (S["dof_pos"] - P["default_dof_pos"]) * P["obs_scale_dof_pos"]
    dq = S["dof_vel"] * P["obs_scale_dof_vel"]
    diff = S["dof_pos"] - S["ref_dof_pos"]
    priv = torch.cat((
        cmd_in, q, dq, S["actions"], diff,
        S["base_lin_vel"] * P["obs_scale_lin_vel"], S["base_ang_vel"] * P["obs_scale_ang_vel"],
        S["base_euler_xyz"] * P["obs_scale_quat"], S["rand_push_force"][:, :2], S["rand_push_torque"],
        S["env_frictions"], S["body_mass"] / 30., stance, contact), dim=-1)
    obs = torch.cat((cmd_in, q, dq, S["actions"], S["base_ang_vel"] * P["obs_scale_ang_vel"],
                     S["base_euler_xyz"] * P["obs_scale_quat"]), dim=-1)
    T = P.get("terrain")
    if T is not None and T["measure_heights"]:                  # humanoid_env.py:246-248: [STACKED obs of the last step | heights]
        heights = torch.clip(S["root_states"][:, 2].unsqueeze(1) - 0.5 - S["measured_heights"], -1, 1.) * T["height_scale"]
        priv = torch.cat((S["obs_buf"], heights), dim=-1)
    if P["add_noise"]:
        obs = obs + z_obs * P["noise_scale_vec"] * P["noise_level"]
    S["obs_hist"] = torch.cat((S["obs_hist"][:, 1:], obs.unsqueeze(1)), dim=1)
    S["critic_hist"] = torch.cat((S["critic_hist"][:, 1:], priv.unsqueeze(1)), dim=1)
    S["obs_buf"] = S["obs_hist"].reshape(obs.shape[0], -1)
    S["privileged_obs_buf"] = S["critic_hist"].reshape(obs.shape[0], -1)


def post_physics(S, P, noise):
    """legged_robot.py:119-154 followed by the +-18 clip of legged_robot.py:104-108.

    noise: dict of dense tensors u_cmd_cb (N,3), u_push (N,5), u_dof (N,12),
    u_cmd_rs (N,3), z_obs (N,47)."""
    S["episode_length_buf"] = S["episode_length_buf"] + 1
    S["common_step_counter"] += 1
    q = S["root_states"][:, 3:7]
    S["base_lin_vel"] = quat_rotate_inverse(q, S["root_states"][:, 7:10])
    S["base_ang_vel"] = quat_rotate_inverse(q, S["root_states"][:, 10:13])
    g = torch.tensor([0., 0., -1.]).repeat(q.shape[0], 1)
    S["projected_gravity"] = quat_rotate_inverse(q, g)
    S["base_euler_xyz"] = euler_xyz_wrapped(q)
    step_callback(S, P, noise)
    check_termination(S, P)
    compute_reward(S, P)
    ids = S["reset_buf"].nonzero(as_tuple=False).flatten()
    reset_idx(S, P, ids, noise)
    compute_observations(S, P, noise["z_obs"])
    S["last_last_actions"] = S["last_actions"].clone()
    S["last_actions"] = S["actions"].clone()
    S["last_dof_vel"] = S["dof_vel"].clone()
    S["last_root_vel"] = S["root_states"][:, 7:13].clone()
    c = P["clip_obs"]
    S["obs_buf"] = torch.clip(S["obs_buf"], -c, c)
    S["privileged_obs_buf"] = torch.clip(S["privileged_obs_buf"], -c, c)
    return S["obs_buf"], S["privileged_obs_buf"], S["rew_buf"], S["reset_buf"]
