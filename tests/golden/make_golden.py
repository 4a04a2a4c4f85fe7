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
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HG_REFERENCE_ROOT", "/root/reference")


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
        self._cmd_call = 0

    def noise(self):
        return {k: getattr(self, k).copy() for k in
                ("u_cmd_cb", "u_cmd_rs", "u_dof", "u_push", "z_obs", "u_delay", "z_act")}

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
        else:
            raise RuntimeError(f"unexpected draw in ctx {self.ctx}")
        return (upper - lower) * u + lower


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

    real_rand, real_randn_like = torch.rand, torch.randn_like

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
    which = sys.argv[1:] or ["env", "ppo", "kat", "cfg"]
    if "env" in which:
        make_env_golden(os.path.join(HERE, "env_rollout.npz"))
    if "ppo" in which:
        make_ppo_golden(os.path.join(HERE, "ppo_learning.npz"))
    if "cfg" in which:
        make_cfg_golden(os.path.join(HERE, "cfg_dump.json"))
    if "kat" in which:
        make_policy_kat(os.path.join(HERE, "policy_example_kat.npz"))
