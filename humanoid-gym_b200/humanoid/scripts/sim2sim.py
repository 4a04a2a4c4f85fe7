"""Deployment loop of an exported policy: python humanoid/scripts/sim2sim.py --load_model policy_1.pt
(reference scripts/sim2sim.py:87-164).

CPU plumbing around the TorchScript actor that `export_policy_as_jit` writes: at 100 Hz assemble the 47-wide
observation frame from the robot state (gait clock, commanded velocity, joint positions / velocities, last action,
base angular velocity, Euler angles), push it through the 15-frame history, call the policy, clip, turn the action
into PD position targets; at 1 kHz evaluate the PD law.  The rigid-body step itself is MuJoCo's (`mujoco==2.3.6`,
not installable in this image): when `import mujoco` fails the state comes from a seeded synthetic source and the
script says so -- everything the policy sees is still built exactly as the reference builds it, which is what the
C1 "sim2sim plumbing" baseline of BASELINE.json times.

This file deliberately imports nothing of the native library: deployment runs on a CPU-only box.
"""
import argparse
import math
import time
from collections import deque

import numpy as np
import torch


class cmd:
    vx, vy, dyaw = 0.4, 0.0, 0.0


class Sim2simCfg:
    """The constants of reference sim2sim.py:170-192 / XBotLCfg that the loop reads."""
    num_actions, num_single_obs, frame_stack = 12, 47, 15
    num_observations = 47 * 15
    clip_observations = clip_actions = 18.0
    action_scale = 0.25
    obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_dof_pos, obs_scale_dof_vel = 2.0, 1.0, 1.0, 0.05
    sim_duration, dt, decimation = 60.0, 0.001, 10
    kps = np.array([200, 200, 350, 350, 15, 15, 200, 200, 350, 350, 15, 15], dtype=np.double)
    kds = np.full(12, 10.0)
    tau_limit = np.full(12, 200.0)


def quaternion_to_euler_array(quat):                     # reference sim2sim.py:46-66 (xyzw)
    x, y, z, w = quat
    roll = np.arctan2(2.0 * (w * x + y * z), 1.0 - 2.0 * (x * x + y * y))
    pitch = np.arcsin(np.clip(2.0 * (w * y - z * x), -1.0, 1.0))
    yaw = np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z))
    return np.array([roll, pitch, yaw])


def pd_control(target_q, q, kp, target_dq, dq, kd):      # :79-82
    return (target_q - q) * kp + (target_dq - dq) * kd


def assemble_obs(cfg, count_lowlevel, q, dq, action, omega, quat):
    """One 47-wide frame, reference sim2sim.py:124-139."""
    obs = np.zeros([1, cfg.num_single_obs], dtype=np.float32)
    eu = quaternion_to_euler_array(quat)
    eu[eu > math.pi] -= 2 * math.pi
    obs[0, 0] = math.sin(2 * math.pi * count_lowlevel * cfg.dt / 0.64)
    obs[0, 1] = math.cos(2 * math.pi * count_lowlevel * cfg.dt / 0.64)
    obs[0, 2] = cmd.vx * cfg.obs_scale_lin_vel
    obs[0, 3] = cmd.vy * cfg.obs_scale_lin_vel
    obs[0, 4] = cmd.dyaw * cfg.obs_scale_ang_vel
    obs[0, 5:17] = q * cfg.obs_scale_dof_pos
    obs[0, 17:29] = dq * cfg.obs_scale_dof_vel
    obs[0, 29:41] = action
    obs[0, 41:44] = omega
    obs[0, 44:47] = eu
    return np.clip(obs, -cfg.clip_observations, cfg.clip_observations)


class SyntheticRobot:
    """Seeded stand-in for mujoco.MjData when MuJoCo is unavailable: plausible joint / base state, no dynamics."""

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)
        self.q, self.dq = np.zeros(12), np.zeros(12)
        self.quat = np.array([0.0, 0.0, 0.0, 1.0])
        self.omega = np.zeros(3)

    def step(self, tau):
        r = self.rng
        self.dq = 0.98 * self.dq + 0.002 * tau / 10.0 + 0.01 * r.standard_normal(12)
        self.q = np.clip(self.q + 0.001 * self.dq, -1.0, 1.0)
        self.omega = 0.95 * self.omega + 0.02 * r.standard_normal(3)
        rpy = 0.02 * r.standard_normal(3)
        cr, sr, cp, sp, cy, sy = (math.cos(rpy[0] / 2), math.sin(rpy[0] / 2), math.cos(rpy[1] / 2), math.sin(rpy[1] / 2),
                                  math.cos(rpy[2] / 2), math.sin(rpy[2] / 2))
        self.quat = np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                              cr * cp * cy + sr * sp * sy])


def run(policy, cfg=Sim2simCfg, robot=None, low_level_steps=None, record=None):
    """The loop of reference sim2sim.py:113-160 without the viewer.  Returns (policy calls, seconds)."""
    robot = robot or SyntheticRobot()
    steps = int(cfg.sim_duration / cfg.dt) if low_level_steps is None else low_level_steps
    target_q = np.zeros(cfg.num_actions)
    action = np.zeros(cfg.num_actions)
    hist = deque(np.zeros([1, cfg.num_single_obs], dtype=np.double) for _ in range(cfg.frame_stack))
    calls = 0
    t0 = time.time()
    with torch.no_grad():
        for count in range(steps):
            q, dq = robot.q[-cfg.num_actions:], robot.dq[-cfg.num_actions:]
            if count % cfg.decimation == 0:                               # 1000 Hz -> 100 Hz
                obs = assemble_obs(cfg, count, q, dq, action, robot.omega, robot.quat)
                hist.append(obs)
                hist.popleft()
                policy_input = np.zeros([1, cfg.num_observations], dtype=np.float32)
                for i in range(cfg.frame_stack):
                    policy_input[0, i * cfg.num_single_obs:(i + 1) * cfg.num_single_obs] = hist[i][0, :]
                action[:] = policy(torch.tensor(policy_input))[0].detach().numpy()
                action = np.clip(action, -cfg.clip_actions, cfg.clip_actions)
                target_q = action * cfg.action_scale
                calls += 1
                if record is not None:
                    record.append((policy_input.copy(), action.copy()))
            tau = pd_control(target_q, q, cfg.kps, np.zeros(cfg.num_actions), dq, cfg.kds)
            tau = np.clip(tau, -cfg.tau_limit, cfg.tau_limit)
            robot.step(tau)
    return calls, time.time() - t0


def main():
    ap = argparse.ArgumentParser(description="Deployment script.")
    ap.add_argument("--load_model", type=str, required=True, help="TorchScript policy (export_policy_as_jit -> policy_1.pt)")
    ap.add_argument("--terrain", action="store_true", help="(MuJoCo only)")
    ap.add_argument("--duration", type=float, default=Sim2simCfg.sim_duration)
    args = ap.parse_args()
    try:
        import mujoco  # noqa: F401
        raise SystemExit("MuJoCo found: wire mujoco.MjData into run(robot=...) as reference sim2sim.py:98-102,154-160 does")
    except ImportError:
        print("MuJoCo unavailable: synthetic robot state, policy + observation assembly + PD law only")
    policy = torch.jit.load(args.load_model, map_location="cpu")
    calls, sec = run(policy, low_level_steps=int(args.duration / Sim2simCfg.dt))
    print(f"{calls} policy calls in {sec:.2f} s -> {calls / sec:.0f} calls/s ({1e3 * sec / calls:.3f} ms per 100 Hz control tick)")


if __name__ == "__main__":
    main()
