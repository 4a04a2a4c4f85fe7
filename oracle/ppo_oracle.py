"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the
reference's PPO learning side in plain fp32 torch (autograd for gradients).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.

Pinned by tests/test_oracle_ppo.py against tests/golden/ppo_learning.npz
(forward, log-prob, entropy, GAE, a complete PPO.update() with per-step
gradients, adaptive learning rates and final weights, all produced by the
unmodified reference) and tests/golden/policy_example_kat.npz (the reference's
own shipped actor).  `file:line` citations are relative to
/root/reference/humanoid/algo/ppo/.

torch.nn.functional.elu / torch.optim.Adam / clip_grad_norm_ are the
reference's own third-party arithmetic (PyTorch), used here as-is.
"""
import math

import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def param_names(n_actor_layers=4, n_critic_layers=4):
    """nn.Module.named_parameters() order of the reference ActorCritic (actor_critic.py:36-83)."""
    names = ["std"]
    for i in range(n_actor_layers):
        names += [f"actor.{2 * i}.weight", f"actor.{2 * i}.bias"]
    for i in range(n_critic_layers):
        names += [f"critic.{2 * i}.weight", f"critic.{2 * i}.bias"]
    return names


def init_params(num_actor_obs, num_critic_obs, num_actions, actor_hidden, critic_hidden, init_noise_std=1.0,
                generator=None):
    """Fresh parameters with nn.Linear's default init (kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)))."""
    def lin(i, o):
        bound = 1.0 / math.sqrt(i)
        w = (torch.rand(o, i, generator=generator) * 2 - 1) * bound
        b = (torch.rand(o, generator=generator) * 2 - 1) * bound
        return w, b
    p = {"std": init_noise_std * torch.ones(num_actions)}
    dims = [num_actor_obs] + list(actor_hidden) + [num_actions]
    for k in range(len(dims) - 1):
        p[f"actor.{2 * k}.weight"], p[f"actor.{2 * k}.bias"] = lin(dims[k], dims[k + 1])
    dims = [num_critic_obs] + list(critic_hidden) + [1]
    for k in range(len(dims) - 1):
        p[f"critic.{2 * k}.weight"], p[f"critic.{2 * k}.bias"] = lin(dims[k], dims[k + 1])
    return p


def mlp(x, p, prefix):
    """nn.Sequential(Linear, ELU, ..., Linear) -- actor_critic.py:54-77."""
    k = 0
    while f"{prefix}.{2 * k}.weight" in p:
        x = F.linear(x, p[f"{prefix}.{2 * k}.weight"], p[f"{prefix}.{2 * k}.bias"])
        if f"{prefix}.{2 * (k + 1)}.weight" in p:
            x = F.elu(x)
        k += 1
    return x


def actor_dist(obs, p):
    """update_distribution, actor_critic.py:111-113: Normal(mean, mean*0 + std)."""
    mean = mlp(obs, p, "actor")
    return mean, mean * 0. + p["std"]


def log_prob(actions, mean, sigma):
    """Normal.log_prob(...).sum(-1), actor_critic.py:119-120."""
    var = sigma ** 2
    return (-((actions - mean) ** 2) / (2 * var) - sigma.log() - LOG_SQRT_2PI).sum(dim=-1)


def entropy(sigma):
    """Normal.entropy().sum(-1), actor_critic.py:107-109."""
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(dim=-1)


def act(obs, critic_obs, p, eps):
    """PPO.act, ppo.py:91-101.  eps: injected N(0,1) of shape (N, num_actions)."""
    with torch.no_grad():
        mean, sigma = actor_dist(obs, p)
        actions = mean + sigma * eps
        values = mlp(critic_obs, p, "critic")
        return actions, values, log_prob(actions, mean, sigma), mean, sigma


def bootstrap_timeouts(rewards, values, time_outs, gamma):
    """ppo.py:107-108: r += gamma * V * time_out."""
    return rewards + gamma * torch.squeeze(values * time_outs.unsqueeze(1), 1)


def gae(rewards, values, dones, last_values, gamma, lam):
    """RolloutStorage.compute_returns, rollout_storage.py:122-136.  Shapes (T,N,1); dones uint8."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for t in reversed(range(T)):
        nxt = last_values if t == T - 1 else values[t + 1]
        not_term = 1.0 - dones[t].float()
        delta = rewards[t] + not_term * gamma * nxt - values[t]
        adv = delta + not_term * gamma * lam * adv
        returns[t] = adv + values[t]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


def ppo_loss(p, batch, clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.001, use_clipped_value_loss=True):
    """ppo.py:133-168 for one minibatch.  Returns (loss, surrogate, value_loss, kl_mean)."""
    obs, cobs, actions, target_values, advantages, returns, old_logp, old_mu, old_sigma = batch
    mu, sigma = actor_dist(obs, p)
    logp = log_prob(actions, mu, sigma)
    value = mlp(cobs, p, "critic")
    ent = entropy(sigma)
    with torch.no_grad():
        kl = torch.sum(torch.log(sigma / old_sigma + 1.e-5)
                       + (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma)) - 0.5,
                       axis=-1)
        kl_mean = torch.mean(kl)
    ratio = torch.exp(logp - torch.squeeze(old_logp))
    a = torch.squeeze(advantages)
    surrogate = torch.max(-a * ratio, -a * torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param)).mean()
    if use_clipped_value_loss:
        v_clip = target_values + (value - target_values).clamp(-clip_param, clip_param)
        value_loss = torch.max((value - returns).pow(2), (v_clip - returns).pow(2)).mean()
    else:
        value_loss = (returns - value).pow(2).mean()
    loss = surrogate + value_loss_coef * value_loss - entropy_coef * ent.mean()
    return loss, surrogate, value_loss, kl_mean


def adapt_lr(lr, kl_mean, desired_kl=0.01):
    """ppo.py:142-145."""
    if kl_mean > desired_kl * 2.0:
        return max(1e-5, lr / 1.5)
    if kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
        return min(1e-2, lr * 1.5)
    return lr


class Learner:
    """PPO.update (ppo.py:119-184) + mini_batch_generator (rollout_storage.py:146-182)."""

    def __init__(self, params, lr=1e-5, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2,
                 value_loss_coef=1.0, entropy_coef=0.001, max_grad_norm=1.0, desired_kl=0.01,
                 schedule="adaptive", use_clipped_value_loss=True):
        self.names = list(params.keys())
        self.p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.lr = lr
        self.opt = torch.optim.Adam([self.p[k] for k in self.names], lr=lr)
        self.epochs, self.mbs = num_learning_epochs, num_mini_batches
        self.clip_param, self.vcoef, self.ecoef = clip_param, value_loss_coef, entropy_coef
        self.max_grad_norm, self.desired_kl, self.schedule = max_grad_norm, desired_kl, schedule
        self.use_clipped_value_loss = use_clipped_value_loss
        self.grad_log = None

    def flat_grad(self):
        return torch.cat([self.p[k].grad.reshape(-1) for k in self.names])

    def update(self, st, perm):
        """st: dict of (T,N,.) tensors named as RolloutStorage attributes; perm: the randperm to use."""
        flat = {k: v.flatten(0, 1) for k, v in st.items()}
        mb = perm.numel() // self.mbs
        mv = ms = 0.0
        for _ in range(self.epochs):
            for i in range(self.mbs):
                idx = perm[i * mb:(i + 1) * mb]
                batch = tuple(flat[k][idx] for k in (
                    "observations", "privileged_observations", "actions", "values", "advantages", "returns",
                    "actions_log_prob", "mu", "sigma"))
                loss, sur, vl, kl = ppo_loss(self.p, batch, self.clip_param, self.vcoef, self.ecoef,
                                             self.use_clipped_value_loss)
                if self.desired_kl is not None and self.schedule == "adaptive":
                    self.lr = adapt_lr(self.lr, kl, self.desired_kl)
                    for gp in self.opt.param_groups:
                        gp["lr"] = self.lr
                self.opt.zero_grad()
                loss.backward()
                if self.grad_log is not None:
                    self.grad_log.append((self.flat_grad().clone(), self.lr))
                torch.nn.utils.clip_grad_norm_([self.p[k] for k in self.names], self.max_grad_norm)
                self.opt.step()
                mv += vl.item()
                ms += sur.item()
        n = self.epochs * self.mbs
        return mv / n, ms / n
