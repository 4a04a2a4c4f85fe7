"""ORACLE (test infrastructure) -- the reference's training iteration (on_policy_runner.py:124-170) assembled
from the CPU restatements in env_oracle.py / ppo_oracle.py.  Used by bench.py for `cpu_baseline` and the
`--impl reference` arm (the reference's own PyTorch path, timed on the host cores), and by the multi-rank
gloo tests.  Never imported by the product package."""
import time

import torch

from . import env_oracle as eo
from . import ppo_oracle as po


class OracleTrainer:
    def __init__(self, num_envs, physics, T=60, seed=5, device="cpu", actor_hidden=(512, 256, 128),
                 critic_hidden=(768, 256, 128), lr=1e-5, epochs=2, mini_batches=4, gamma=0.994, lam=0.9):
        self.N, self.T, self.gamma, self.lam = num_envs, T, gamma, lam
        self.P = eo.make_params()
        self.gen = torch.Generator().manual_seed(seed)
        self.S = eo.new_state(num_envs, env_origins=eo.grid_origins(num_envs),
                              env_frictions=0.1 + 1.9 * torch.rand(num_envs, 1, generator=self.gen),
                              body_mass=5.0 + 10 * torch.rand(num_envs, 1, generator=self.gen) - 5)
        self.physics = physics
        params = po.init_params(705, 219, 12, actor_hidden, critic_hidden, 1.0, generator=self.gen)
        self.learner = po.Learner(params, lr=lr, num_learning_epochs=epochs, num_mini_batches=mini_batches)
        z = torch.zeros
        self.st = dict(observations=z(T, num_envs, 705), privileged_observations=z(T, num_envs, 219),
                       actions=z(T, num_envs, 12), rewards=z(T, num_envs, 1), dones=z(T, num_envs, 1).byte(),
                       values=z(T, num_envs, 1), actions_log_prob=z(T, num_envs, 1), mu=z(T, num_envs, 12),
                       sigma=z(T, num_envs, 12))
        all_ids = torch.arange(num_envs)
        eo.reset_idx(self.S, self.P, all_ids, self._noise())
        eo.compute_observations(self.S, self.P, self._noise()["z_obs"])
        self.S["episode_length_buf"] = torch.randint(0, 2400, (num_envs,), generator=self.gen)

    def _noise(self):
        N, g = self.N, self.gen
        return dict(u_cmd_cb=torch.rand(N, 3, generator=g), u_cmd_rs=torch.rand(N, 3, generator=g),
                    u_dof=torch.rand(N, 12, generator=g), u_push=torch.rand(N, 5, generator=g),
                    z_obs=torch.randn(N, 47, generator=g), u_delay=torch.rand(N, 1, generator=g),
                    z_act=torch.randn(N, 12, generator=g))

    def _pull_physics(self, dof_only=False):
        ph, S, N = self.physics, self.S, self.N
        d = ph.dof_state.view(N, 12, 2)
        S["dof_pos"], S["dof_vel"] = d[..., 0], d[..., 1]
        if not dof_only:
            S["root_states"] = ph.root_states.clone()
            S["contact_forces"] = ph.contact_forces.view(N, -1, 3)
            S["rigid_state"] = ph.rigid_state.view(N, -1, 13)

    def env_step(self, actions):
        S, P, ph = self.S, self.P, self.physics
        noise = self._noise()
        eo.pre_physics(S, P, actions, noise["u_delay"], noise["z_act"])
        for _ in range(10):
            eo.compute_torques(S, P)
            ph.simulate()
            ph.refresh_dof_state_tensor()
            self._pull_physics(dof_only=True)
        ph.refresh_actor_root_state_tensor()
        ph.refresh_net_contact_force_tensor()
        ph.refresh_rigid_body_state_tensor()
        self._pull_physics()
        return eo.post_physics(S, P, noise)

    def iteration(self):
        """One learning iteration; returns (collection_time, learn_time) like the reference runner."""
        S, st, p = self.S, self.st, self.learner.p
        t0 = time.time()
        with torch.inference_mode():
            obs, cobs = S["obs_buf"], S["privileged_obs_buf"]
            for t in range(self.T):
                eps = torch.randn(self.N, 12, generator=self.gen)
                a, v, lp, mu, sg = po.act(obs, cobs, p, eps)
                st["observations"][t], st["privileged_observations"][t] = obs, cobs
                st["actions"][t], st["values"][t], st["actions_log_prob"][t] = a, v, lp.unsqueeze(1)
                st["mu"][t], st["sigma"][t] = mu, sg
                obs, cobs, rew, dones = self.env_step(a)
                rew = po.bootstrap_timeouts(rew, v, S["extras_time_outs"], self.gamma)
                st["rewards"][t], st["dones"][t] = rew.unsqueeze(1), dones.unsqueeze(1).byte()
            last_v = po.mlp(cobs, p, "critic")
            st["returns"], st["advantages"] = po.gae(st["rewards"], st["values"], st["dones"], last_v, self.gamma, self.lam)
        t1 = time.time()
        perm = torch.randperm(self.N * self.T, generator=self.gen)
        self.learner.update({k: v.clone() for k, v in st.items()}, perm)
        t2 = time.time()
        return t1 - t0, t2 - t1


def episode_book_step(cur_reward_sum, cur_episode_length, rewards, dones):
    """ORACLE restatement of the runner's per-step episode bookkeeping (reference on_policy_runner.py:140-154):
    returns (finished episode rewards, finished episode lengths) of this step; mutates the two running tensors."""
    cur_reward_sum += rewards
    cur_episode_length += 1
    new_ids = (dones > 0).nonzero(as_tuple=False)
    rew = cur_reward_sum[new_ids][:, 0].clone()
    ln = cur_episode_length[new_ids][:, 0].clone()
    cur_reward_sum[new_ids] = 0
    cur_episode_length[new_ids] = 0
    return rew, ln
