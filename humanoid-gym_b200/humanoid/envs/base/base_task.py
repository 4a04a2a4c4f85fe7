"""BaseTask: device selection, the VecEnv-facing buffers and getters
(mirrors reference envs/base/base_task.py:43-145; viewer / camera code is out of scope)."""
import os

import torch

from humanoid import _native


class BaseTask:
    def __init__(self, cfg, sim_params, physics_engine, sim_device, headless):
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        self.headless = headless
        dev = torch.device(sim_device)
        if dev.type != "cuda":
            raise _native.NativeError(
                f"sim_device={sim_device!r}: the humanoid_ppo hot path runs on sm_100a only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.sim_device_id = dev.index
        self.graphics_device_id = self.sim_device_id
        self.device = str(dev)

        self.num_envs = cfg.env.num_envs
        self.num_obs = cfg.env.num_observations
        self.num_privileged_obs = cfg.env.num_privileged_obs
        self.num_actions = cfg.env.num_actions

        z = dict(device=self.device)
        # observation rows are stored with a pitch rounded up to 32 floats (705 -> 736, 219 -> 224): every row starts on
        # a 128-byte line (the env kernel's warp-wide stores then write whole lines) and TMA can feed the rows to the
        # tensor-core actor / critic; obs_buf / privileged_obs_buf are the (N, 705) / (N, 219) views
        def pitched(width):
            q = int(os.environ.get("HG_OBS_PITCH_ALIGN", "32"))
            return torch.zeros(self.num_envs, (width + q - 1) // q * q, dtype=torch.float, **z)[:, :width]
        self.obs_buf = pitched(self.num_obs)
        self.rew_buf = torch.zeros(self.num_envs, dtype=torch.float, **z)
        # the reference allocates int64 ones but rebinds a bool tensor on every step
        # (legged_robot.py:159); the fused kernel always writes bool
        self.reset_buf = torch.ones(self.num_envs, dtype=torch.bool, **z)
        self._episode_length_buf = torch.zeros(self.num_envs, dtype=torch.long, **z)
        self.time_out_buf = torch.zeros(self.num_envs, dtype=torch.bool, **z)
        if self.num_privileged_obs is not None:
            self.privileged_obs_buf = pitched(self.num_privileged_obs)
        else:
            self.privileged_obs_buf = None
        self.extras = {}

        self.create_sim()
        self.enable_viewer_sync = True
        self.viewer = None

    # OnPolicyRunner *assigns* env.episode_length_buf (on_policy_runner.py:103-106); keep the
    # device buffer the kernels hold a pointer to and copy into it instead of rebinding.
    @property
    def episode_length_buf(self):
        return self._episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value):
        self._episode_length_buf.copy_(value)

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset_idx(self, env_ids):
        raise NotImplementedError

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(
            torch.zeros(self.num_envs, self.num_actions, device=self.device, requires_grad=False))
        return obs, privileged_obs

    def step(self, actions):
        raise NotImplementedError

    def render(self, sync_frame_time=True):
        return None        # headless only
