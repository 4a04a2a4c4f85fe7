"""LeggedRobot: host side of the vectorised env step.

Same public surface as reference envs/base/legged_robot.py (ctor, step, reset, post_physics_step,
check_termination, compute_reward, reset_idx, _compute_torques, the live state tensors), but every
tensor op between the physics refresh and the return of step() is ONE launch of the fused
sm_100a kernel `hg_env_post_physics`; this class only owns the buffers, the physics seam and the
argument marshalling.

Rough terrain (cfg.terrain.mesh_type 'heightfield' / 'trimesh'; XBotLCfg ships 'plane', humanoid_config.py:72-76):
the height field is built on the host by utils/terrain.py; per step the env samples it around every robot
(`_get_heights`), moves terminated envs through the terrain curriculum (`_update_terrain_curriculum`) and spawns
them within 1 m of their terrain origin -- three small kernels of csrc/hg_terrain.cu around the fused kernel, which
then runs as two launches (see post_physics_step).  The viewer is out of scope.
"""
import os

import numpy as np
import torch

from humanoid import _native as nat
from humanoid.envs.base.base_task import BaseTask
from humanoid.utils.helpers import class_to_dict
from humanoid.utils.terrain import Terrain
from humanoid import physics as phys

# the two halves of a rough-terrain step: the curriculum sits between termination and reset (reference :175-186)
_PHASES_BEFORE_RESET = nat.PHASE_COUNTERS | nat.PHASE_CALLBACK | nat.PHASE_TERMINATE | nat.PHASE_REWARD
_PHASES_FROM_RESET = nat.PHASE_RESET | nat.PHASE_OBS | nat.PHASE_LAST


class LeggedRobot(BaseTask):
    terrain_class = Terrain                                      # XBotLFreeEnv: HumanoidTerrain (humanoid_env.py:153)

    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless):
        self.cfg = cfg
        self.sim_params = sim_params
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        self._parse_cfg(self.cfg)
        super().__init__(self.cfg, sim_params, physics_engine, sim_device, headless)
        self._init_buffers()
        self._prepare_reward_function()
        self._bind_native()
        self.init_done = True

    # ------------------------------------------------------------------------------------------
    # configuration
    # ------------------------------------------------------------------------------------------
    def _parse_cfg(self, cfg):                                   # reference legged_robot.py:710-720
        self.dt = self.cfg.control.decimation * self.sim_params.dt
        self.obs_scales = self.cfg.normalization.obs_scales
        self.reward_scales = class_to_dict(self.cfg.rewards.scales)
        self.command_ranges = {k: (list(v) if isinstance(v, (list, tuple)) else v)          # own lists: the command curriculum
                               for k, v in class_to_dict(self.cfg.commands.ranges).items()}   # widens them in place
        if self.cfg.terrain.mesh_type not in ("heightfield", "trimesh"):
            self.cfg.terrain.curriculum = False
        self.max_episode_length_s = self.cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.cfg.domain_rand.push_interval = np.ceil(self.cfg.domain_rand.push_interval_s / self.dt)

    def create_sim(self):
        """Physics seam (reference humanoid_env.py:145-163 + legged_robot.py:588-681)."""
        self.up_axis_idx = 2
        mesh_type = self.cfg.terrain.mesh_type                    # humanoid_env.py:151-162
        if mesh_type not in (None, "plane", "heightfield", "trimesh"):
            raise ValueError("Terrain mesh type not recognised. Allowed types are [None, plane, heightfield, trimesh]")
        self.terrain = self.terrain_class(self.cfg.terrain, self.num_envs) if mesh_type in ("heightfield", "trimesh") else None
        if self.terrain is not None:                              # _create_heightfield / _create_trimesh :570,586
            self.height_samples = torch.tensor(self.terrain.heightsamples).view(
                self.terrain.tot_rows, self.terrain.tot_cols).to(self.device).contiguous()
        self._get_env_origins()
        kind = os.environ.get("HG_PHYSICS", getattr(self.cfg, "physics_backend", "auto"))
        seed = getattr(self.cfg, "seed", 0)
        rank = int(os.environ.get("RANK", "0"))
        self.gym = phys.make_physics(kind, self.num_envs, self.device, self.cfg, self.env_origins, seed=seed, rank=rank,
                                     sim_params=self.sim_params, physics_engine=self.physics_engine,
                                     sim_device_id=self.sim_device_id, custom_origins=self.custom_origins)
        self.sim = self.gym
        if self.terrain is not None:
            self.gym.add_terrain(self.terrain, mesh_type)
        real_sim = hasattr(self.gym, "create_actors")             # a simulator: actors carry the randomised properties
        if real_sim:
            sim_frictions, sim_masses = self.gym.create_actors()  # _create_envs :638-665
            self.gym.prepare()                                    # base_task.py:96 + tensor acquisition :438-457
        self.num_dof = self.num_dofs = self.gym.num_dof
        self.num_bodies = self.gym.num_bodies
        self.dof_names = list(self.gym.dof_names)
        body_names = list(self.gym.body_names)
        dev = self.device

        def indices(pattern_list):
            names = []
            for pat in pattern_list:
                names.extend(s for s in body_names if pat in s)
            find = self.gym.body_index if real_sim else body_names.index       # find_actor_rigid_body_handle :667-681
            return torch.tensor([find(n) for n in names], dtype=torch.long, device=dev)

        self.feet_indices = indices([self.cfg.asset.foot_name])
        self.knee_indices = indices([self.cfg.asset.knee_name])
        self.penalised_contact_indices = indices(self.cfg.asset.penalize_contacts_on)
        self.termination_contact_indices = indices(self.cfg.asset.terminate_after_contacts_on)

        props = self.gym.dof_properties()                         # _process_dof_props :276-293
        s = self.cfg.safety
        self.dof_pos_limits = torch.tensor(list(zip(props["lower"], props["upper"])), dtype=torch.float, device=dev) * s.pos_limit
        self.dof_vel_limits = torch.tensor(props["velocity"], dtype=torch.float, device=dev) * s.vel_limit
        self.torque_limits = torch.tensor(props["effort"], dtype=torch.float, device=dev) * s.torque_limit

        init = self.cfg.init_state
        self.base_init_state = torch.tensor(init.pos + init.rot + init.lin_vel + init.ang_vel, dtype=torch.float, device=dev)

        N = self.num_envs
        dr = self.cfg.domain_rand
        if real_sim:                                              # drawn by the actor loop, per env
            self.env_frictions = sim_frictions.to(dev)
            self.body_mass = sim_masses.to(dev)
            if self.gym.friction_coeffs is not None:
                self.friction_coeffs = self.gym.friction_coeffs
            return
        # one-time domain randomisation, drawn on the CPU like the reference (:257-270, :296-302)
        self.env_frictions = torch.zeros(N, 1, dtype=torch.float32, device=dev)
        if dr.randomize_friction:
            buckets = 256
            ids = torch.randint(0, buckets, (N, 1))
            lo, hi = dr.friction_range
            vals = (hi - lo) * torch.rand(buckets, 1) + lo
            self.friction_coeffs = vals[ids]
            self.env_frictions[:] = self.friction_coeffs.view(N, 1).to(dev)
        self.body_mass = torch.full((N, 1), phys.robot.BASE_LINK_MASS, dtype=torch.float32, device=dev)
        if dr.randomize_base_mass:
            lo, hi = dr.added_mass_range
            self.body_mass += torch.from_numpy(np.random.uniform(lo, hi, size=(N, 1)).astype(np.float32)).to(dev)

    def _get_env_origins(self):                                  # :683-708
        N = self.num_envs
        if self.terrain is not None:                              # origins are the terrain platforms :687-697
            tc = self.cfg.terrain
            self.custom_origins = True
            max_init_level = tc.max_init_terrain_level if tc.curriculum else tc.num_rows - 1
            self.terrain_levels = torch.randint(0, max_init_level + 1, (N,)).to(self.device)
            self.terrain_types = torch.div(torch.arange(N), (N / tc.num_cols), rounding_mode="floor").to(torch.long).to(self.device)
            self.max_terrain_level = tc.num_rows
            self.terrain_origins = torch.from_numpy(self.terrain.env_origins).to(self.device).to(torch.float).contiguous()
            self.env_origins = self.terrain_origins[self.terrain_levels, self.terrain_types].contiguous()
            return
        self.custom_origins = False                               # flat ground: a grid :698-708
        cols = np.floor(np.sqrt(N))
        rows = np.ceil(N / cols)
        xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
        sp = self.cfg.env.env_spacing
        o = torch.zeros(N, 3)
        o[:, 0] = sp * xx.flatten()[:N]
        o[:, 1] = sp * yy.flatten()[:N]
        self.env_origins = o.to(self.device)

    # ------------------------------------------------------------------------------------------
    # buffers
    # ------------------------------------------------------------------------------------------
    def _init_buffers(self):                                     # :434-516
        N, dev = self.num_envs, self.device
        g = self.gym
        self.root_states = g.root_states
        self.dof_state = g.dof_state
        self.dof_pos = self.dof_state.view(N, self.num_dof, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, self.num_dof, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.contact_forces = g.contact_forces.view(N, -1, 3)
        self.rigid_state = g.rigid_state.view(N, -1, 13)

        def zeros(*shape, dtype=torch.float):
            return torch.zeros(*shape, dtype=dtype, device=dev, requires_grad=False)

        self.common_step_counter = 0
        self.extras = {}
        self.noise_scale_vec = self._get_noise_scale_vec(self.cfg)
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0], device=dev).repeat((N, 1))
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0], device=dev).repeat((N, 1))
        A = self.num_actions
        self.torques = zeros(N, A)
        self.p_gains = zeros(N, A)
        self.d_gains = zeros(N, A)
        self.actions = zeros(N, A)
        self.last_actions = zeros(N, A)
        self.last_last_actions = zeros(N, A)
        self.last_rigid_state = torch.zeros_like(self.rigid_state)    # dead data in the reference (:151)
        self.last_dof_vel = zeros(N, self.num_dof)
        self.last_root_vel = zeros(N, 6)
        self.commands = zeros(N, self.cfg.commands.num_commands)
        self.commands_scale = torch.tensor([self.obs_scales.lin_vel, self.obs_scales.lin_vel, self.obs_scales.ang_vel], device=dev)
        self.feet_air_time = zeros(N, self.feet_indices.shape[0])
        self.last_contacts = zeros(N, len(self.feet_indices), dtype=torch.bool)
        self.base_lin_vel = zeros(N, 3)
        self.base_ang_vel = zeros(N, 3)
        self.projected_gravity = zeros(N, 3)
        self.projected_gravity[:, 2] = -1.0
        self.base_euler_xyz = zeros(N, 3)
        if self.cfg.terrain.measure_heights:                      # :481-483
            self.height_points = self._init_height_points()
        self.measured_heights = 0
        self.feet_height = zeros(N, 2)
        self.last_feet_z = torch.full((N, 2), 0.05, device=dev)
        self.ref_dof_pos = zeros(N, self.num_dof)
        self.rand_push_force = zeros(N, 3)
        self.rand_push_torque = zeros(N, 3)
        self.extras_time_outs = zeros(N, dtype=torch.bool)
        self.reset_ids = zeros(N, dtype=torch.int32)
        self._scratch = zeros(32, dtype=torch.int32)

        self.default_dof_pos = zeros(self.num_dof)
        for i, name in enumerate(self.dof_names):                # :487-501
            self.default_dof_pos[i] = self.cfg.init_state.default_joint_angles[name]
            for key in self.cfg.control.stiffness.keys():
                if key in name:
                    self.p_gains[:, i] = self.cfg.control.stiffness[key]
                    self.d_gains[:, i] = self.cfg.control.damping[key]
        self.default_dof_pos = self.default_dof_pos.unsqueeze(0)
        self.default_joint_pd_target = self.default_dof_pos.clone()

    def _get_noise_scale_vec(self, cfg):
        raise NotImplementedError

    def _init_height_points(self):                               # :743-757
        """Base-frame grid (num_envs, num_height_points, 3) around every robot at which the terrain is sampled."""
        tc = self.cfg.terrain
        y = torch.tensor(tc.measured_points_y, device=self.device)
        x = torch.tensor(tc.measured_points_x, device=self.device)
        grid_x, grid_y = torch.meshgrid(x, y, indexing="ij")
        self.num_height_points = grid_x.numel()
        self._height_points_xy = torch.stack((grid_x.flatten(), grid_y.flatten()), dim=1).contiguous()   # the kernel's (P, 2) view
        points = torch.zeros(self.num_envs, self.num_height_points, 3, device=self.device)
        points[:, :, 0] = grid_x.flatten()
        points[:, :, 1] = grid_y.flatten()
        return points

    def _prepare_reward_function(self):                          # :518-541
        for key in list(self.reward_scales.keys()):
            if self.reward_scales[key] == 0:
                self.reward_scales.pop(key)
            else:
                self.reward_scales[key] *= self.dt
        if "termination" in self.reward_scales:
            # reference legged_robot.py:231-235 adds a post-clip termination reward; the fused kernel has no such term
            # (XBotLCfg: scale -0.0, filtered above).  Refuse rather than silently drop it.
            raise NotImplementedError("rewards.scales.termination != 0 is not supported by the fused env kernel")
        self.reward_names = list(self.reward_scales.keys())
        K = len(self.reward_names)
        self._episode_sums = torch.zeros(K, self.num_envs, dtype=torch.float, device=self.device)
        self._episode_means = torch.zeros(K, dtype=torch.float, device=self.device)
        self.episode_sums = {n: self._episode_sums[k] for k, n in enumerate(self.reward_names)}

    # ------------------------------------------------------------------------------------------
    # native binding
    # ------------------------------------------------------------------------------------------
    def _native_params(self):
        raise NotImplementedError

    def _bind_native(self):
        self._P = self._native_params()
        B = nat.EnvBuffers()
        tensors = dict(
            root_states=self.root_states, dof_state=self.dof_state, contact_forces=self.gym.contact_forces,
            rigid_state=self.gym.rigid_state, actions=self.actions, last_actions=self.last_actions,
            last_last_actions=self.last_last_actions, torques=self.torques, last_dof_vel=self.last_dof_vel,
            last_root_vel=self.last_root_vel, commands=self.commands, episode_length_buf=self._episode_length_buf,
            reset_buf=self.reset_buf, time_out_buf=self.time_out_buf, extras_time_outs=self.extras_time_outs,
            base_lin_vel=self.base_lin_vel, base_ang_vel=self.base_ang_vel, projected_gravity=self.projected_gravity,
            base_euler_xyz=self.base_euler_xyz, feet_air_time=self.feet_air_time, last_contacts=self.last_contacts,
            feet_height=self.feet_height, last_feet_z=self.last_feet_z, ref_dof_pos=self.ref_dof_pos,
            rand_push_force=self.rand_push_force, rand_push_torque=self.rand_push_torque,
            env_frictions=self.env_frictions, body_mass=self.body_mass, env_origins=self.env_origins,
            episode_sums=self._episode_sums, episode_means=self._episode_means, rew_terms=None,
            rew_buf=self.rew_buf, reset_ids=self.reset_ids, scratch=self._scratch)
        # observation histories are a ping-pong pair: the kernel reads frames 1..14 of one buffer and writes
        # frames 0..13 (+ the new frame) of the other -- like the reference, env.obs_buf is rebound every step
        def twin(v):        # same pitch, fresh storage
            return torch.zeros(v.shape[0], v.stride(0), dtype=v.dtype, device=v.device)[:, :v.shape[1]]
        self._obs_pp = [self.obs_buf, twin(self.obs_buf)]
        if self.cfg.terrain.measure_heights:
            # critic frames carry the terrain heights (humanoid_env.py:246-248): env.privileged_obs_buf is the
            # c_frame_stack x (num_obs + num_height_points) history written by hg_terrain_priv_frames; the fused kernel keeps
            # its 3 x 73 frames in a private pair
            self._privh_pp = [self.privileged_obs_buf, twin(self.privileged_obs_buf)]
            w = nat.PRIV1 * nat.PRIV_FRAMES
            first = torch.zeros(self.num_envs, (w + 31) // 32 * 32, device=self.device)[:, :w]
            self._priv_pp = [first, twin(first)]
        else:
            self._privh_pp = None
            self._priv_pp = [self.privileged_obs_buf, twin(self.privileged_obs_buf)]
        for k, t in tensors.items():
            setattr(B, k, nat.ptr(t))
        B.obs_buf, B.obs_out = self._obs_pp[0].data_ptr(), self._obs_pp[1].data_ptr()
        B.privileged_obs_buf, B.priv_out = self._priv_pp[0].data_ptr(), self._priv_pp[1].data_ptr()
        B.obs_pitch, B.priv_pitch = self.obs_buf.stride(0), self._priv_pp[0].stride(0)
        self._B = B
        self._keepalive = tensors
        self._bind_terrain()
        self._Z = nat.EnvNoise()
        self._Z.seed = int(getattr(self.cfg, "seed", 0)) * 0x9E3779B97F4A7C15 % (1 << 64) + int(os.environ.get("RANK", "0"))
        self._noise_step = 0
        self._injected = {}
        self._dev_index = torch.device(self.device).index

    def _bind_terrain(self):
        """Rough terrain: HgTerrain descriptor, the spawn-origin buffer the fused kernel's reset reads instead of
        env_origins (origin + the +-1 m jitter of :381-384), the heights buffer."""
        self._T = None
        self._launches_per_step = 1
        if self.terrain is None:
            if self.cfg.terrain.measure_heights:                  # plane: _get_heights returns zeros (:772-773)
                self._heights = torch.zeros(self.num_envs, self.num_height_points, device=self.device)
            return
        if float(self.base_init_state[0]) != 0.0 or float(self.base_init_state[1]) != 0.0:
            # spawn = (init + origin) + jitter in the reference; the kernels form init + (origin + jitter): equal iff init_xy = 0
            raise NotImplementedError("rough terrain needs init_state.pos[0:2] == 0 (XBotLCfg: [0, 0, 0.95])")
        tc, T = self.cfg.terrain, nat.Terrain()
        T.height_samples = self.height_samples.data_ptr()
        T.rows, T.cols = self.height_samples.shape
        T.border_size, T.horizontal_scale, T.vertical_scale = tc.border_size, tc.horizontal_scale, tc.vertical_scale
        T.terrain_origins = self.terrain_origins.data_ptr()
        T.num_levels, T.num_types = self.terrain_origins.shape[0], self.terrain_origins.shape[1]
        T.half_env_length = self.terrain.env_length / 2
        T.max_episode_length_s = self.max_episode_length_s
        T.curriculum = int(bool(tc.curriculum))
        self._T = T
        self._spawn = self.env_origins.clone()
        self._B.env_origins = self._spawn.data_ptr()
        self._launches_per_step = 2
        if tc.measure_heights:
            self._heights = torch.zeros(self.num_envs, self.num_height_points, device=self.device)

    # ---- CUDA-graph support: per-step counters move to device memory -----------------------------------
    def use_device_counters(self, on=True):
        """With device counters the launches of a step carry no per-step host value (common_step_counter and the
        Philox step live in scratch[4..7] and are bumped by the kernel), so a captured rollout can be replayed."""
        if on and not self._Z.use_device_counters:
            c = torch.tensor([self.common_step_counter, self._noise_step], dtype=torch.int64, device=self.device)
            self._scratch[4:8].copy_(c.view(torch.int32))
        elif not on and self._Z.use_device_counters:
            c = self._scratch[4:8].clone().view(torch.int64).tolist()
            self.common_step_counter, self._noise_step = int(c[0]), int(c[1])
        self._Z.use_device_counters = int(on)

    @property
    def noise_step_dev_ptr(self):
        """Device address of the Philox step counter (uint64 at scratch[6..7])."""
        return self._scratch.data_ptr() + 24

    def advance_host_counters(self, steps):
        """After replaying a captured rollout of `steps` env steps: keep the host mirrors in sync."""
        self.common_step_counter += steps
        self._noise_step += steps * self._launches_per_step
        if hasattr(self.gym, "substep"):
            self.gym.substep += steps * self.cfg.control.decimation

    def graph_safe(self, steps=None):
        """A rollout of `steps` env steps may be captured once and replayed: the synthetic source is a ring, so every
        replay must start at the same ring phase (device-resident frames and pinned host frames alike: the H2D
        copies become memcpy nodes of the graph)."""
        if not isinstance(self.gym, phys.SyntheticPhysics):
            return False
        if self.cfg.commands.curriculum:          # a host-side decision inside some steps (update_command_curriculum)
            return False
        return steps is None or steps % self.gym.ring == 0

    def inject_noise(self, **tensors):
        """Parity hook: dense per-env draws (u_cmd_cb, u_cmd_rs, u_dof, u_push, z_obs, u_delay, z_act; rough terrain:
        u_root (N,2), r_level (N) int64) used instead of in-kernel Philox for the NEXT kernel call(s) of this step."""
        self._injected = {k: v.to(self.device, torch.int64 if k == "r_level" else torch.float32).contiguous()
                          for k, v in tensors.items()}

    def _launch_post_physics(self, phases):
        Z = self._Z
        inj = self._injected
        for k in ("u_cmd_cb", "u_cmd_rs", "u_dof", "u_push", "z_obs"):
            setattr(Z, k, nat.ptr(inj.get(k)))
        Z.step = self._noise_step
        self._noise_step += 1
        B = self._B
        if self.obs_buf is self._obs_pp[0]:
            src, dst = 0, 1
        elif self.obs_buf is self._obs_pp[1]:
            src, dst = 1, 0
        else:       # user code rebound env.obs_buf: adopt its contents
            self._obs_pp[0].copy_(self.obs_buf)
            (self._privh_pp or self._priv_pp)[0].copy_(self.privileged_obs_buf)
            src, dst = 0, 1
        B.obs_buf, B.privileged_obs_buf = self._obs_pp[src].data_ptr(), self._priv_pp[src].data_ptr()
        B.obs_out, B.priv_out = self._obs_pp[dst].data_ptr(), self._priv_pp[dst].data_ptr()
        nat.check(nat.lib.hg_env_post_physics(B, self._P, Z, phases, int(self.common_step_counter),
                                              self.num_envs, nat.stream_ptr(self._dev_index)), "hg_env_post_physics")
        now = dst if phases & nat.PHASE_OBS else src
        self.obs_buf = self._obs_pp[now]
        if self._privh_pp is None:
            self.privileged_obs_buf = self._priv_pp[now]
        elif phases & nat.PHASE_OBS:
            # humanoid_env.py:246-248: the critic frame is [the obs_buf this step STARTED with | scaled heights]
            cfg = self.cfg
            nat.check(nat.lib.hg_terrain_priv_frames(
                self._obs_pp[src].data_ptr(), self.obs_buf.stride(0), self.num_obs, nat.ptr(self.root_states),
                nat.ptr(self._heights), self.num_height_points, self.obs_scales.height_measurements,
                cfg.normalization.clip_observations, nat.ptr(self.reset_buf) if phases & nat.PHASE_RESET else None,
                self._privh_pp[src].data_ptr(), self._privh_pp[dst].data_ptr(), self._privh_pp[0].stride(0),
                cfg.env.c_frame_stack, self.num_envs, nat.stream_ptr(self._dev_index)), "hg_terrain_priv_frames")
            self.privileged_obs_buf = self._privh_pp[dst]
        else:
            self.privileged_obs_buf = self._privh_pp[src]

    # ------------------------------------------------------------------------------------------
    # stepping
    # ------------------------------------------------------------------------------------------
    def step(self, actions):                                      # :84-109
        """`actions` are already the clipped / delayed / noised actions in self.actions when called
        from XBotLFreeEnv.step; a direct call clips and stores them first."""
        if actions is not self.actions:
            clip = self.cfg.normalization.clip_actions
            self.actions.copy_(torch.clip(actions, -clip, clip))
        g = self.gym
        st = nat.stream_ptr(self._dev_index)
        self._refreshed = bool(g.fused_decimation(self))          # synthetic source: the whole loop + refreshes, one launch
        if not self._refreshed:
            for _ in range(self.cfg.control.decimation):
                nat.check(nat.lib.hg_env_compute_torques(self._B, self._P, self.num_envs, st), "hg_env_compute_torques")
                g.set_dof_actuation_force_tensor(self.torques)
                g.simulate()
                g.refresh_dof_state_tensor()
        self.post_physics_step()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _compute_torques(self, actions):                          # :340-356
        if actions is not self.actions:
            self.actions.copy_(actions)
        nat.check(nat.lib.hg_env_compute_torques(self._B, self._P, self.num_envs, nat.stream_ptr(self._dev_index)),
                  "hg_env_compute_torques")
        return self.torques

    def post_physics_step(self):                                  # :119-154 + the clip of :104-108
        g = self.gym
        if not getattr(self, "_refreshed", False):
            g.refresh_actor_root_state_tensor()
            g.refresh_net_contact_force_tensor()
            g.refresh_rigid_body_state_tensor()
        self._refreshed = False
        self.common_step_counter += 1
        if self.cfg.terrain.measure_heights:                      # _post_physics_step_callback :316-317
            self.measured_heights = self._get_heights()
        # avoid updating the command curriculum at each step: the maximum command is common to all envs (:178-180)
        cmd_cur = bool(self.cfg.commands.curriculum) and self.common_step_counter % self.max_episode_length == 0
        if self._T is None and not cmd_cur:
            self._launch_post_physics(nat.PHASE_STEP_ALL)
        else:
            # rough terrain / command-curriculum steps: termination + rewards, then what reset_idx does BEFORE it re-initialises
            # the envs that terminated -- terrain curriculum and spawn origins (:175-177, :381-384), command curriculum
            # (:178-180; a host decision, hence a sync on those rare steps) --, then reset + observations + last_* copies
            self._launch_post_physics(_PHASES_BEFORE_RESET)
            if self._T is not None:
                self._terrain_reset_prepare(curriculum=True)
            if cmd_cur:
                env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
                if len(env_ids):
                    self.update_command_curriculum(env_ids)
            self._launch_post_physics(_PHASES_FROM_RESET)
        self._injected = {}
        pushed = self.cfg.domain_rand.push_robots and (self.common_step_counter % self.cfg.domain_rand.push_interval == 0)
        g.apply_env_writes(self.reset_ids, self._scratch, pushed)
        self._publish_extras()

    # ---- rough terrain -------------------------------------------------------------------------------------
    def _get_heights(self, env_ids=None):                         # :759-795
        """Terrain heights (num_envs, num_height_points) at the grid points around each robot: one launch of
        hg_terrain_get_heights; zeros on a plane.  `env_ids` selects rows of the result."""
        if self.cfg.terrain.mesh_type == "plane":
            h = self._heights
        elif self.cfg.terrain.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        else:
            nat.check(nat.lib.hg_terrain_get_heights(self._T, nat.ptr(self.root_states), nat.ptr(self._height_points_xy),
                                                     self.num_height_points, nat.ptr(self._heights), self.num_envs,
                                                     nat.stream_ptr(self._dev_index)), "hg_terrain_get_heights")
            h = self._heights
        return h if env_ids is None else h[env_ids]

    def _terrain_reset_prepare(self, curriculum):
        """For the envs flagged in reset_buf: terrain curriculum (if enabled) and the spawn origins of the reset."""
        T = self._T
        saved = T.curriculum
        T.curriculum = int(bool(curriculum and saved))
        inj, Z = self._injected, self._Z
        dev_ctr = Z.use_device_counters
        nat.check(nat.lib.hg_terrain_reset_prepare(
            T, nat.ptr(self.reset_buf), nat.ptr(self.root_states), nat.ptr(self.commands), nat.ptr(self.terrain_levels),
            nat.ptr(self.terrain_types), nat.ptr(self.env_origins), nat.ptr(self._spawn), nat.ptr(inj.get("r_level")),
            nat.ptr(inj.get("u_root")), Z.seed, self._noise_step, self.noise_step_dev_ptr if dev_ctr else None,
            self.num_envs, nat.stream_ptr(self._dev_index)), "hg_terrain_reset_prepare")
        T.curriculum = saved

    def _update_terrain_curriculum(self, env_ids):                # :400-420
        """Game-inspired curriculum for the envs being reset (stand-alone entry point; a step runs it between its two
        fused launches)."""
        if not self.init_done or self._T is None or not self.cfg.terrain.curriculum:
            return
        previous = self.reset_buf.clone()
        self.reset_buf.zero_()
        self.reset_buf[env_ids] = True
        self._terrain_reset_prepare(curriculum=True)
        self.reset_buf.copy_(previous)

    def update_command_curriculum(self, env_ids):                 # :422-431
        """Curriculum of increasing commands: when the envs being reset tracked their velocity command well (mean episode
        sum above 80 % of the maximum), widen the lin_vel_x range by 0.5 m/s on both sides up to commands.max_curriculum.
        The comparison is the reference's (fp32 mean / max_episode_length against a Python float) and, like there, a
        device->host read; the widened range reaches the kernels through the HgEnvParams block of the next launch."""
        k = self.reward_names.index("tracking_lin_vel")
        if torch.mean(self._episode_sums[k][env_ids]) / self.max_episode_length > 0.8 * self.reward_scales["tracking_lin_vel"]:
            r, m = self.command_ranges["lin_vel_x"], self.cfg.commands.max_curriculum
            r[0] = np.clip(r[0] - 0.5, -m, 0.0)
            r[1] = np.clip(r[1] + 0.5, 0.0, m)
            self._P.cmd_x_lo, self._P.cmd_x_span = r[0], r[1] - r[0]      # span formed in double, like torch_rand_float's

    def _publish_extras(self):
        if "episode" not in self.extras:
            self.extras["episode"] = {"rew_" + n: self._episode_means[k] for k, n in enumerate(self.reward_names)}
        if self.cfg.commands.curriculum:                           # :206-207
            self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        if self.cfg.terrain.mesh_type == "trimesh":                # :204-205 (the reference refreshes it on steps with a reset;
            self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())   # levels only change on those)
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.extras_time_outs

    # the individual stages stay callable, each as the same kernel with a phase mask
    def check_termination(self):                                  # :156-161
        self._launch_post_physics(nat.PHASE_TERMINATE)

    def compute_reward(self):                                     # :217-235
        self._launch_post_physics(nat.PHASE_REWARD)

    def compute_observations(self):
        self._launch_post_physics(nat.PHASE_OBS)

    def reset_idx(self, env_ids):                                 # :163-215
        if len(env_ids) == 0:
            return
        previous = self.reset_buf.clone()
        self.reset_buf.zero_()
        self.reset_buf[env_ids] = True                            # the kernel resets the masked envs
        if self._T is not None:
            self._terrain_reset_prepare(curriculum=self.init_done)     # "don't change on initial reset" :407-409
        if self.cfg.commands.curriculum and self.common_step_counter % self.max_episode_length == 0:
            self.update_command_curriculum(env_ids)
        self._launch_post_physics(nat.PHASE_RESET)
        self.reset_buf |= previous                                # reference only sets [env_ids] = 1 (:196)
        self._injected = {}
        self._publish_extras()

    @property
    def last_reset_count(self):
        """Number of envs reset by the most recent step (device->host read)."""
        return int(self._scratch[3].item())
