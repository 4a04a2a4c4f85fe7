"""OnPolicyRunner: rollout + update loop, logging, checkpoints
(reference algo/ppo/on_policy_runner.py:47-307; same checkpoint format and console/TensorBoard scalars).

Differences that matter for speed only: episode bookkeeping stays on the device (no per-step .cpu()),
and TensorBoard / wandb are optional imports."""
import os
import statistics
import time
from collections import deque
from datetime import datetime

import torch

from .ppo import PPO
from .actor_critic import ActorCritic
from humanoid.algo.vec_env import VecEnv  # noqa: F401

_CLASSES = {"PPO": PPO, "ActorCritic": ActorCritic}


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cpu"):
        self.cfg = train_cfg["runner"]
        self.alg_cfg = train_cfg["algorithm"]
        self.policy_cfg = train_cfg["policy"]
        self.all_cfg = train_cfg
        self.wandb_run_name = (datetime.now().strftime("%b%d_%H-%M-%S") + "_" + train_cfg["runner"]["experiment_name"]
                               + "_" + train_cfg["runner"]["run_name"])
        self.device = device
        self.env = env
        num_critic_obs = self.env.num_privileged_obs if self.env.num_privileged_obs is not None else self.env.num_obs
        actor_critic = _CLASSES[self.cfg["policy_class_name"]](
            self.env.num_obs, num_critic_obs, self.env.num_actions, **self.policy_cfg).to(self.device)
        self.alg = _CLASSES[self.cfg["algorithm_class_name"]](actor_critic, device=self.device, **self.alg_cfg)
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg["save_interval"]
        self.alg.init_storage(self.env.num_envs, self.num_steps_per_env, [self.env.num_obs],
                              [self.env.num_privileged_obs], [self.env.num_actions])
        self.log_dir = log_dir
        self.writer = None
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.last_perf = {}
        _, _ = self.env.reset()

    # ------------------------------------------------------------------------------------------
    def _init_writer(self):
        try:
            import wandb
            wandb.init(project="XBot", sync_tensorboard=True, name=self.wandb_run_name, config=self.all_cfg,
                       mode=os.environ.get("WANDB_MODE", "disabled"))
        except Exception as e:      # no network in most deployments
            print(f"wandb disabled: {e}")
        from torch.utils.tensorboard import SummaryWriter
        self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)

    def rollout(self, obs, critic_obs, book=None):
        """One collection phase: num_steps_per_env x (act, env.step, process_env_step) + compute_returns."""
        env, alg = self.env, self.alg
        step_dev = env.noise_step_dev_ptr if getattr(env, "_Z", None) is not None and env._Z.use_device_counters else None
        begin = getattr(getattr(env, "gym", None), "begin_rollout", None)
        if begin is not None:
            begin(self.num_steps_per_env)                 # host-resident frames: a self-contained prefetch schedule
        for t in range(self.num_steps_per_env):
            # the Philox step of the action noise is the env's noise-step counter: its device copy when the launches must
            # be graph-replayable, the host mirror otherwise -- the same value either way, so replay == eager bit for bit
            actions = alg.act(obs, critic_obs, step_dev=step_dev, step=getattr(env, "_noise_step", None))
            obs, privileged_obs, rewards, dones, infos = env.step(actions)
            critic_obs = privileged_obs if privileged_obs is not None else obs
            alg.process_env_step(rewards, dones, infos)
            if book is not None:
                book.step(t, rewards, dones, infos)
        alg.compute_returns(critic_obs)
        return obs, critic_obs

    # ---- CUDA graph of the whole collection phase -----------------------------------------------------
    def _graph_ok(self):
        if os.environ.get("HG_CUDA_GRAPH", "1") == "0" or self.num_steps_per_env % 2:
            return False
        fn = getattr(self.env, "graph_safe", None)
        return bool(fn(self.num_steps_per_env)) if fn is not None else False

    def collect(self, obs, critic_obs, book=None):
        """rollout(), replayed from a CUDA graph after the first (eager, warm-up) call when the physics source
        allows it: ~2000 launches of a 60-step rollout become one graph launch, nothing per-step crosses PCIe."""
        if not self._graph_ok():
            return self.rollout(obs, critic_obs, book)
        env, alg = self.env, self.alg
        key = (obs.data_ptr(), critic_obs.data_ptr(), book is not None)
        if getattr(self, "_graph_key", None) != key:
            self._graph, self._graph_key, self._graph_warm = None, key, 0
        if self._graph is None:
            if self._graph_warm < 1:                  # first call with these buffers: eager (JIT, attributes, caches)
                self._graph_warm += 1
                return self.rollout(obs, critic_obs, book)
            env.use_device_counters(True)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                host_before = (env.common_step_counter, env._noise_step, getattr(env.gym, "substep", 0))
                from humanoid import _native as nat
                n0 = nat.launch_count()
                with torch.cuda.graph(g, stream=side):
                    out = self.rollout(obs, critic_obs, book)
                self._graph_launches = nat.launch_count() - n0      # native kernels inside one replay
                # capture executed nothing on the device: rewind the host mirrors the Python loop advanced
                env.common_step_counter, env._noise_step = host_before[0], host_before[1]
                if hasattr(env.gym, "substep"):
                    env.gym.substep = host_before[2]
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph, self._graph_out = g, out
        self._graph.replay()
        self.replayed_launches = getattr(self, "replayed_launches", 0) + self._graph_launches
        env.advance_host_counters(self.num_steps_per_env)
        alg.storage.step = self.num_steps_per_env
        return self._graph_out

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None:
            self._init_writer()
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs = self.env.get_observations()
        privileged_obs = self.env.get_privileged_observations()
        critic_obs = privileged_obs if privileged_obs is not None else obs
        self.alg.actor_critic.train()
        book = (_EpisodeBook(self.env.num_envs, self.num_steps_per_env, len(self.env.extras.get("episode", {})), self.device)
                if self.log_dir is not None else None)

        tot_iter = self.current_learning_iteration + num_learning_iterations
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            with torch.inference_mode():
                obs, critic_obs = self.collect(obs, critic_obs, book)
                torch.cuda.synchronize(self.device)
                stop = time.time()
                collection_time = stop - start
                start = stop
                mean_value_loss, mean_surrogate_loss = self.alg.update()
            stop = time.time()
            learn_time = stop - start
            self.last_perf = dict(collection_time=collection_time, learn_time=learn_time,
                                  fps=self.num_steps_per_env * self.env.num_envs / (collection_time + learn_time))
            if self.log_dir is not None:
                ep_infos = book.drain_infos()
                rewbuffer, lenbuffer = book.rewbuffer, book.lenbuffer
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, "model_{}.pt".format(it)))
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, "model_{}.pt".format(self.current_learning_iteration)))

    # ------------------------------------------------------------------------------------------
    def log(self, locs, width=80, pad=35):
        self.tot_timesteps += self.num_steps_per_env * self.env.num_envs
        iteration_time = locs["collection_time"] + locs["learn_time"]
        self.tot_time += iteration_time
        it = locs["it"]
        ep_string = ""
        for key, value in locs["ep_infos"].items():
            self.writer.add_scalar("Episode/" + key, value, it)
            ep_string += f"""{f'Mean episode {key}:':>{pad}} {value:.4f}\n"""
        mean_std = self.alg.actor_critic.std.mean().item()
        fps = int(self.num_steps_per_env * self.env.num_envs / iteration_time)
        w = self.writer
        w.add_scalar("Loss/value_function", locs["mean_value_loss"], it)
        w.add_scalar("Loss/surrogate", locs["mean_surrogate_loss"], it)
        w.add_scalar("Loss/learning_rate", self.alg.learning_rate, it)
        w.add_scalar("Policy/mean_noise_std", mean_std, it)
        w.add_scalar("Perf/total_fps", fps, it)
        w.add_scalar("Perf/collection time", locs["collection_time"], it)
        w.add_scalar("Perf/learning_time", locs["learn_time"], it)
        have_eps = len(locs["rewbuffer"]) > 0
        if have_eps:
            mr, ml = statistics.mean(locs["rewbuffer"]), statistics.mean(locs["lenbuffer"])
            w.add_scalar("Train/mean_reward", mr, it)
            w.add_scalar("Train/mean_episode_length", ml, it)
            w.add_scalar("Train/mean_reward/time", mr, self.tot_time)
            w.add_scalar("Train/mean_episode_length/time", ml, self.tot_time)
        head = f" \033[1m Learning iteration {it}/{self.current_learning_iteration + locs['num_learning_iterations']} \033[0m "
        out = (f"""{'#' * width}\n{head.center(width, ' ')}\n\n"""
               f"""{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, learning {locs['learn_time']:.3f}s)\n"""
               f"""{'Value function loss:':>{pad}} {locs['mean_value_loss']:.4f}\n"""
               f"""{'Surrogate loss:':>{pad}} {locs['mean_surrogate_loss']:.4f}\n"""
               f"""{'Mean action noise std:':>{pad}} {mean_std:.2f}\n""")
        if have_eps:
            out += (f"""{'Mean reward:':>{pad}} {mr:.2f}\n"""
                    f"""{'Mean episode length:':>{pad}} {ml:.2f}\n""")
        out += ep_string
        out += (f"""{'-' * width}\n"""
                f"""{'Total timesteps:':>{pad}} {self.tot_timesteps}\n"""
                f"""{'Iteration time:':>{pad}} {iteration_time:.2f}s\n"""
                f"""{'Total time:':>{pad}} {self.tot_time:.2f}s\n"""
                f"""{'ETA:':>{pad}} {self.tot_time / (it + 1) * (locs['num_learning_iterations'] - it):.1f}s\n""")
        print(out)

    def save(self, path, infos=None):
        self.alg.sync_optimizer_container()
        torch.save({"model_state_dict": self.alg.actor_critic.state_dict(),
                    "optimizer_state_dict": self.alg.optimizer.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        loaded = torch.load(path, map_location=self.device)
        self.alg.actor_critic.load_state_dict(loaded["model_state_dict"])
        if load_optimizer:
            self.alg.optimizer.load_state_dict(loaded["optimizer_state_dict"])
            self.alg.load_optimizer_container()
        self.current_learning_iteration = loaded["iter"]
        return loaded["infos"]

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference

    def get_inference_critic(self, device=None):
        self.alg.actor_critic.eval()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.evaluate


class _EpisodeBook:
    """Per-episode reward / length bookkeeping of on_policy_runner.py:140-154: ONE native launch per env step
    (hg_episode_book_step) into preallocated (T, N) slabs, CUDA-graph capturable; finished-episode statistics are read
    back once per iteration instead of once per step."""

    def __init__(self, n, T, n_infos, device):
        z = dict(dtype=torch.float, device=device)
        self.n, self.n_infos = n, n_infos
        self.cur_reward_sum = torch.zeros(n, **z)
        self.cur_episode_length = torch.zeros(n, **z)
        self.done_rew = torch.full((T, n), float("nan"), **z)
        self.done_len = torch.full((T, n), float("nan"), **z)
        self.infos = torch.zeros(T, max(n_infos, 1), **z)
        self.rewbuffer, self.lenbuffer = deque(maxlen=100), deque(maxlen=100)
        self._info_keys = []
        self._dev_index = torch.device(device).index

    def step(self, t, rewards, dones, infos):
        from humanoid import _native as nat
        means = None
        if isinstance(infos, dict) and infos.get("episode"):
            self._info_keys = list(infos["episode"].keys())
            vals = list(infos["episode"].values())
            base = getattr(vals[0], "_base", None)
            # the env publishes the 22 means as views of ONE (22,) tensor: pass it straight through; anything else is stacked
            if base is not None and base.is_contiguous() and base.numel() == len(vals) and all(getattr(v, "_base", None) is base for v in vals):
                means = base
            else:
                means = torch.stack([v.reshape(()) for v in vals]).contiguous()
        d = dones if dones.dtype in (torch.bool, torch.uint8) else (dones > 0)
        nat.check(nat.lib.hg_episode_book_step(
            rewards.contiguous().data_ptr(), d.contiguous().data_ptr(), self.cur_reward_sum.data_ptr(), self.cur_episode_length.data_ptr(),
            self.done_rew[t].data_ptr(), self.done_len[t].data_ptr(), None if means is None else means.data_ptr(), self.infos[t].data_ptr(),
            0 if means is None else min(len(self._info_keys), self.infos.shape[1]), self.n, nat.stream_ptr(self._dev_index)),
            "hg_episode_book_step")

    def drain_infos(self):
        r, ln = self.done_rew.flatten(), self.done_len.flatten()
        keep = ~torch.isnan(r)
        self.rewbuffer.extend(r[keep].cpu().tolist())
        self.lenbuffer.extend(ln[keep].cpu().tolist())
        out = {}
        if self._info_keys:
            m = self.infos[:, :len(self._info_keys)].mean(dim=0).cpu().tolist()
            out = dict(zip(self._info_keys, m))
        return out
