"""XBotLFreeEnv: the humanoid_ppo task (reference envs/custom/humanoid_env.py).

The gait clock, reference trajectory, 47/73-wide observation frames with 15/3-frame histories,
domain-randomisation noise, pushes and the 22 reward terms all live inside the fused kernel
(csrc/hg_env.cu); this class supplies the constants (HgEnvParams) and the step prologue."""
import numpy as np
import torch

from humanoid import _native as nat
from humanoid.envs.base.legged_robot import LeggedRobot
from humanoid.utils.terrain import HumanoidTerrain


class XBotLFreeEnv(LeggedRobot):
    terrain_class = HumanoidTerrain                               # humanoid_env.py:153

    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.reset_idx(torch.arange(self.num_envs, device=self.device))     # humanoid_env.py:80-81
        self.compute_observations()

    def _get_noise_scale_vec(self, cfg):                          # humanoid_env.py:166-186
        v = torch.zeros(self.cfg.env.num_single_obs, device=self.device)
        self.add_noise = self.cfg.noise.add_noise
        ns, sc = self.cfg.noise.noise_scales, self.obs_scales
        v[5:17] = ns.dof_pos * sc.dof_pos
        v[17:29] = ns.dof_vel * sc.dof_vel
        v[41:44] = ns.ang_vel * sc.ang_vel
        v[44:47] = ns.quat * sc.quat
        return v

    def _native_params(self):
        cfg, P = self.cfg, nat.EnvParams()
        e = cfg.env
        priv1 = nat.PRIV1
        if cfg.terrain.measure_heights:
            # humanoid_env.py:246-248 replaces the 73-wide critic frame by [obs_buf | heights]: the cfg has to size the critic
            # for it (the reference otherwise fails in the critic's first matmul)
            priv1 = e.num_observations + self.num_height_points
            if e.num_privileged_obs != e.c_frame_stack * priv1:
                raise nat.NativeError(
                    f"measure_heights: the critic frame is num_observations + {self.num_height_points} height points = {priv1} "
                    f"wide; set env.single_num_privileged_obs = {priv1} and env.num_privileged_obs = {e.c_frame_stack * priv1}")
        if (e.num_single_obs, e.frame_stack, e.single_num_privileged_obs, e.c_frame_stack, e.num_actions) != \
                (nat.OBS1, nat.OBS_FRAMES, priv1, nat.PRIV_FRAMES, nat.NUM_DOF):
            raise nat.NativeError("the fused env kernel is specialised to 15x47 / 3x73 observations and 12 actions")
        if e.use_ref_actions:
            raise NotImplementedError("use_ref_actions is not part of the humanoid_ppo hot path")
        if len(self.reward_names) != nat.NUM_REWARDS:
            raise nat.NativeError(f"expected the 22 XBot-L reward terms, got {self.reward_names}")
        P.dt = self.dt
        r, n, dr, c = cfg.rewards, cfg.normalization, cfg.domain_rand, cfg.commands
        P.cycle_time = r.cycle_time
        P.clip_actions, P.clip_obs, P.action_scale = n.clip_actions, n.clip_observations, cfg.control.action_scale
        P.action_delay, P.action_noise = dr.action_delay, dr.action_noise

        def lo_span(rng):                                          # span formed in double, like Python does
            return rng[0], rng[1] - rng[0]
        P.cmd_x_lo, P.cmd_x_span = lo_span(self.command_ranges["lin_vel_x"])
        P.cmd_y_lo, P.cmd_y_span = lo_span(self.command_ranges["lin_vel_y"])
        P.cmd_heading_lo, P.cmd_heading_span = lo_span(self.command_ranges["heading"])
        P.push_vel_lo, P.push_vel_span = lo_span([-dr.max_push_vel_xy, dr.max_push_vel_xy])
        P.push_ang_lo, P.push_ang_span = lo_span([-dr.max_push_ang_vel, dr.max_push_ang_vel])
        P.dof_reset_lo, P.dof_reset_span = lo_span([-0.1, 0.1])
        P.target_joint_pos_scale, P.target_feet_height = r.target_joint_pos_scale, r.target_feet_height
        P.base_height_target, P.min_dist, P.max_dist = r.base_height_target, r.min_dist, r.max_dist
        P.max_contact_force, P.tracking_sigma = r.max_contact_force, r.tracking_sigma
        s = self.obs_scales
        P.obs_scale_lin_vel, P.obs_scale_ang_vel, P.obs_scale_dof_pos = s.lin_vel, s.ang_vel, s.dof_pos
        P.obs_scale_dof_vel, P.obs_scale_quat = s.dof_vel, s.quat
        P.noise_level = cfg.noise.noise_level
        P.max_episode_length_s = self.max_episode_length_s
        P.add_noise, P.only_positive_rewards = int(cfg.noise.add_noise), int(r.only_positive_rewards)
        P.heading_command, P.push_robots = int(c.heading_command), int(dr.push_robots)
        if not c.heading_command:
            raise NotImplementedError("ang_vel_yaw command mode is not part of the humanoid_ppo hot path")
        P.resample_period = int(c.resampling_time / self.dt)
        P.push_interval = int(dr.push_interval)
        P.max_episode_length = int(self.max_episode_length)
        P.num_bodies = self.num_bodies
        P.feet[:] = self.feet_indices.tolist()
        P.knees[:] = self.knee_indices.tolist()
        term, pen = self.termination_contact_indices.tolist(), self.penalised_contact_indices.tolist()
        P.n_term, P.n_pen = len(term), len(pen)
        for i, b in enumerate(term):
            P.term_bodies[i] = b
        for i, b in enumerate(pen):
            P.pen_bodies[i] = b
        for k, name in enumerate(self.reward_names):
            P.reward_scales[k] = self.reward_scales[name]
        P.p_gains[:] = self.p_gains[0].tolist()
        P.d_gains[:] = self.d_gains[0].tolist()
        P.torque_limits[:] = self.torque_limits.tolist()
        P.default_dof_pos[:] = self.default_dof_pos[0].tolist()
        P.noise_scale_vec[:] = self.noise_scale_vec.tolist()
        P.base_init_state[:] = self.base_init_state.tolist()
        return P

    def step(self, actions):                                      # humanoid_env.py:189-197
        inj = self._injected
        actions = actions.to(self.device, torch.float32).contiguous()
        nat.check(nat.lib.hg_env_pre_physics(
            self._B, self._P, nat.ptr(actions), nat.ptr(inj.get("u_delay")), nat.ptr(inj.get("z_act")),
            self._Z.seed, nat.STEP_FROM_DEVICE if self._Z.use_device_counters else self._noise_step, self.num_envs,
            nat.stream_ptr(self._dev_index)), "hg_env_pre_physics")
        return super().step(self.actions)

    def reset_idx(self, env_ids):                                 # humanoid_env.py:264-269
        super().reset_idx(env_ids)
        if self._privh_pp is not None and len(env_ids):           # the height-augmented critic history is zeroed as well
            self.privileged_obs_buf[env_ids] = 0.0

    # conveniences kept from the reference API ------------------------------------------------
    def _get_phase(self):                                         # :100-103
        return self.episode_length_buf * self.dt / self.cfg.rewards.cycle_time

    @property
    def obs_history(self):
        return list(self.obs_buf.view(self.num_envs, self.cfg.env.frame_stack, -1).unbind(1))

    @property
    def critic_history(self):
        return list(self.privileged_obs_buf.view(self.num_envs, self.cfg.env.c_frame_stack, -1).unbind(1))
