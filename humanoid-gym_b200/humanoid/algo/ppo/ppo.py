"""PPO with the reference's constructor and methods (reference algo/ppo/ppo.py:41-184); act /
process_env_step / compute_returns / update are sequences of native launches with NO torch op and NO
host synchronisation inside (the adaptive-KL learning rate and the Adam step count live in HBM).

Env-sharded data parallelism: when torch.distributed is initialised, the flat
gradient buffer -- whose tail carries [surrogate, value loss, entropy, KL] partial means -- is
all-reduced (SUM) once per optimizer step; every mean already uses 1/B_global, so the summed buffer is
exactly the single-process gradient of the G*B-sample minibatch and all ranks take the same lr decision.
"""
import os

import torch
import torch.distributed as dist
import torch.optim as optim

from humanoid import _native as nat
from .actor_critic import ActorCritic
from .rollout_storage import RolloutStorage

_TAIL = 8   # [surrogate, value_loss, entropy, kl_mean, pad...]


class PPO:
    actor_critic: ActorCritic

    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cpu"):
        self.device = device
        if torch.device(device).type != "cuda":
            raise nat.NativeError(f"PPO(device={device!r}): the learning side runs on sm_100a only (no CPU fallback)")
        self.desired_kl, self.schedule = desired_kl, schedule
        self.actor_critic = actor_critic
        self.actor_critic.to(self.device)
        self.storage = None
        self.optimizer = optim.Adam(self.actor_critic.parameters(), lr=learning_rate)   # checkpoint container only
        self.transition = RolloutStorage.Transition()
        self.clip_param, self.num_learning_epochs, self.num_mini_batches = clip_param, num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.gamma, self.lam, self.max_grad_norm = gamma, lam, max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.global_advantage_norm = False      # True: one extra 3-double all-reduce per iteration (section 8e)

        flat = self.actor_critic.flat_params()
        n = self.actor_critic.num_params
        dev = flat.device
        self._dev_index = dev.index
        self._grad = torch.zeros(n + _TAIL, dtype=torch.float32, device=dev)
        self._scalars = self._grad[n:]
        self._exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self._exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._sqnorm = torch.zeros(1, dtype=torch.float64, device=dev)
        self._lr = torch.full((1,), learning_rate, dtype=torch.float64, device=dev)
        self._adam_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._loss_sums = torch.zeros(_TAIL, dtype=torch.float32, device=dev)
        self._sample_step = 0
        self._seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + 12345) % (1 << 64)
        # env-sharded data parallelism: every rank seeds torch identically (set_seed), and the Philox counter of
        # hg_policy_sample is (seed, LOCAL env index, step) -- without the rank in the key all shards would draw the
        # same exploration noise
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else int(__import__("os").environ.get("RANK", "0"))
        self._seed = self.rank_seed(self._seed, rank)
        self._alias_grads_and_state()
        self._mb_scratch = {}
        self._side = torch.cuda.Stream(dev)

    @staticmethod
    def rank_seed(seed, rank):
        """splitmix64 of the rank folded into the Philox key (rank 0 keeps a non-trivial offset too)."""
        z = (rank + 1) * 0x9E3779B97F4A7C15 % (1 << 64)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) % (1 << 64)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) % (1 << 64)
        return (seed ^ z ^ (z >> 31)) % (1 << 64)

    # ------------------------------------------------------------------------------------------
    def _alias_grads_and_state(self):
        """p.grad and the torch Adam state become views of the flat buffers (checkpoint compatibility)."""
        ac = self.actor_critic
        for name, p in ac.named_parameters():
            p.grad = ac.view_of(self._grad, name)
            self.optimizer.state[p] = dict(step=torch.tensor(0.0), exp_avg=ac.view_of(self._exp_avg, name),
                                           exp_avg_sq=ac.view_of(self._exp_avg_sq, name))

    def sync_optimizer_container(self):
        """Refresh the torch.optim.Adam container from the device-resident lr / step (before save)."""
        lr, step = float(self._lr.item()), float(self._adam_step.item())
        for gparam in self.optimizer.param_groups:
            gparam["lr"] = lr
        for st in self.optimizer.state.values():
            st["step"] = torch.tensor(step)

    def load_optimizer_container(self):
        """After optimizer.load_state_dict: pull exp_avg / exp_avg_sq / step / lr back into the flat buffers."""
        ac = self.actor_critic
        step = 0.0
        for name, p in ac.named_parameters():
            st = self.optimizer.state.get(p)
            if not st:
                continue
            ac.view_of(self._exp_avg, name).copy_(st["exp_avg"])
            ac.view_of(self._exp_avg_sq, name).copy_(st["exp_avg_sq"])
            step = float(st["step"])
        self._adam_step.fill_(int(step))
        self._lr.fill_(self.optimizer.param_groups[0]["lr"])
        self._alias_grads_and_state()
        self.sync_optimizer_container()

    @property
    def learning_rate(self):
        return float(self._lr.item())

    @learning_rate.setter
    def learning_rate(self, v):
        self._lr.fill_(float(v))

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    # ------------------------------------------------------------------------------------------
    # rollout
    # ------------------------------------------------------------------------------------------
    def act(self, obs, critic_obs, eps=None, step_dev=None, step=None):
        """ppo.py:91-101.  Everything lands directly in slab t of the storage: actor mean -> mu[t],
        value -> values[t], sampled actions / log-prob / sigma -> actions[t] / actions_log_prob[t] / sigma[t];
        obs and critic obs are copied now, because env.step() overwrites the env's buffers in place."""
        s, ac = self.storage, self.actor_critic
        t = s.step
        if t >= s.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        cur = torch.cuda.current_stream(self._dev_index)
        self._join_critic(cur)                             # (only if the caller skipped process_env_step)
        sample = dict(std=ac.std, eps=eps, actions=s.actions[t], log_prob=s.actions_log_prob[t], sigma=s.sigma[t], seed=self._seed,
                      step=self._sample_step if step is None else int(step), step_dev=step_dev)
        self._side.wait_stream(cur)
        # Actor + critic, all layers, with the sampling epilogue: ONE persistent launch (hg_actor_critic_forward).  The 15 MB
        # observation copy into slab t rides on the side stream: nothing before the update reads it.  (HG_ACT_STREAMS=2 puts
        # the critic into its own launch on the side stream; measured slower -- 184 vs 158 us per step -- because two
        # 200 KB-shared-memory grids fight for the same SMs.)
        fused = ac.chain_eligible(obs, critic_obs)
        split_streams = os.environ.get("HG_ACT_STREAMS", "1") == "2"
        if fused and not split_streams:
            ac.native_chain("both", (obs, critic_obs), (s.mu[t], s.values[t]), sample)
        with torch.cuda.stream(self._side):
            if fused and split_streams:
                ac.native_chain("critic", critic_obs, s.values[t])
            elif not fused:
                ac.native_forward("critic", critic_obs, s.values[t])       # fallback: per-layer launches
            s.add_native(t, obs=obs, priv_obs=critic_obs if s.privileged_observations is not None else None)
        self._critic_pending = True
        if fused and split_streams:
            ac.native_chain("actor", obs, s.mu[t], sample)
        elif not fused:
            ac.native_forward("actor", obs, s.mu[t], sample=sample)
        self._sample_step += 1
        # The value estimate is not needed before process_env_step (r += gamma * V * time_out), so the critic chain is
        # joined there: it overlaps the whole env step instead of sitting on the act -> step critical path.  The env
        # writes its next observations into the OTHER half of its ping-pong buffers, so critic_obs stays intact.
        tr = self.transition
        tr.actions, tr.values, tr.actions_log_prob = s.actions[t], s.values[t], s.actions_log_prob[t]
        tr.action_mean, tr.action_sigma = s.mu[t], s.sigma[t]
        tr.observations, tr.critic_observations = s.observations[t], (
            s.privileged_observations[t] if s.privileged_observations is not None else s.observations[t])
        return tr.actions

    def _join_critic(self, cur):
        if getattr(self, "_critic_pending", False):
            cur.wait_stream(self._side)
            self._critic_pending = False

    def process_env_step(self, rewards, dones, infos):
        """ppo.py:103-113: r += gamma * V * time_out, then store rewards and dones (one launch)."""
        s = self.storage
        t = s.step
        self._join_critic(torch.cuda.current_stream(self._dev_index))
        time_outs = infos.get("time_outs") if isinstance(infos, dict) else None
        d = dones if dones.dtype in (torch.bool, torch.uint8) else dones.to(torch.uint8)
        s.add_native(t, gamma=self.gamma, rewards=rewards.contiguous(), dones=d.contiguous(),
                     time_outs=None if time_outs is None else time_outs.contiguous())
        s.step += 1
        self.transition.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        """ppo.py:115-117 + rollout_storage.py:122-136."""
        s = self.storage
        if not hasattr(self, "_last_values") or self._last_values.shape[0] != s.num_envs:
            self._last_values = torch.empty(s.num_envs, 1, dtype=torch.float32, device=self.device)
        self.actor_critic.native_forward("critic", last_critic_obs, self._last_values)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world > 1 and self.global_advantage_norm:
            s.compute_returns(self._last_values, self.gamma, self.lam, normalise=False)
            dist.all_reduce(s._stats)
            s.normalise_advantages()
        else:
            s.compute_returns(self._last_values, self.gamma, self.lam)

    # ------------------------------------------------------------------------------------------
    # update
    # ------------------------------------------------------------------------------------------
    def _scratch(self, B):
        if B not in self._mb_scratch:
            ac, z = self.actor_critic, dict(dtype=torch.float32, device=self.device)
            A = ac.num_actions
            self._mb_scratch = {B: dict(
                mean=torch.empty(B, A, **z), value=torch.empty(B, 1, **z), d_mean=torch.empty(B, A, **z),
                d_value=torch.empty(B, 1, **z),
                hid_a=torch.empty(B * ac.hidden_width("actor"), **z), dhid_a=torch.empty(B * ac.hidden_width("actor"), **z),
                hid_c=torch.empty(B * ac.hidden_width("critic"), **z), dhid_c=torch.empty(B * ac.hidden_width("critic"), **z))}
        return self._mb_scratch[B]

    def use_split_path(self):
        """PPO.update runs on the split-precision (bf16x3) tensor-core path when the GEMM engine is 4 (default) and the
        architecture is eligible; the fp32-operand engines (0 exact CUDA-core, 1 3xTF32, 2 TF32) stay selectable."""
        return nat.lib.hg_set_gemm_mode(-1) == 4 and self.actor_critic.split_eligible()

    def _scratch_split(self, B):
        key = ("split", B)
        if key not in self._mb_scratch:
            ac, z = self.actor_critic, dict(dtype=torch.float32, device=self.device)
            A = ac.num_actions
            h = dict(dtype=torch.int16, device=self.device)
            self._mb_scratch = {key: dict(
                mean=torch.empty(B, A, **z), value=torch.empty(B, 1, **z), d_mean=torch.empty(B, A, **z),
                d_value=torch.empty(B, 1, **z),
                hid_a=torch.empty(2 * B * ac.hidden_width("actor"), **h), dhid_a=torch.empty(2 * B * ac.hidden_width("actor"), **h),
                hid_c=torch.empty(2 * B * ac.hidden_width("critic"), **h), dhid_c=torch.empty(2 * B * ac.hidden_width("critic"), **h))}
        return self._mb_scratch[key]

    def minibatch_step(self, mb, world=1):
        """Loss forward/backward + gradient (all-reduce) + lr rule + clip + Adam for one minibatch."""
        ac = self.actor_critic
        flat = ac.flat_params()
        st = nat.stream_ptr(self._dev_index)
        split = mb.get("obs_split") is not None and self.use_split_path()
        cur = torch.cuda.current_stream(self._dev_index)
        if split:
            xs_a = nat.Split.of(mb["obs_split"])
            xs_c = nat.Split.of(mb["priv_split"]) if mb.get("priv_split") is not None else xs_a
            B = mb["obs_split"].shape[1]
            w = self._scratch_split(B)
            if ac._wsplit_dirty:
                ac.refresh_split()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):               # critic chain overlaps the actor chain
                ac.native_forward_split("critic", xs_c, w["value"], w["hid_c"])
            ac.native_forward_split("actor", xs_a, w["mean"], w["hid_a"])
        else:
            obs = mb["obs"]
            cobs = mb["priv_obs"] if mb["priv_obs"] is not None else obs
            B = obs.shape[0]
            w = self._scratch(B)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):               # critic chain overlaps the actor chain
                ac.native_forward("critic", cobs, w["value"], hidden=w["hid_c"])
            ac.native_forward("actor", obs, w["mean"], hidden=w["hid_a"])
        cur.wait_stream(self._side)
        a = nat.PpoLossArgs()
        a.mean, a.value, a.std = w["mean"].data_ptr(), w["value"].data_ptr(), ac.std.data_ptr()
        a.actions, a.target_values = mb["actions"].data_ptr(), mb["values"].data_ptr()
        a.advantages, a.returns = mb["advantages"].data_ptr(), mb["returns"].data_ptr()
        a.old_log_prob, a.old_mu, a.old_sigma = mb["old_log_prob"].data_ptr(), mb["old_mu"].data_ptr(), mb["old_sigma"].data_ptr()
        a.d_mean, a.d_value = w["d_mean"].data_ptr(), w["d_value"].data_ptr()
        a.grad_std = self._grad.data_ptr() + 4 * ac._offsets["std"]
        a.scalars = self._scalars.data_ptr()
        a.clip_param, a.value_loss_coef, a.entropy_coef = self.clip_param, self.value_loss_coef, self.entropy_coef
        a.use_clipped_value_loss, a.num_actions = int(self.use_clipped_value_loss), ac.num_actions
        a.inv_B = 1.0 / (B * world)
        nat.check(nat.lib.hg_ppo_loss_fwd_bwd(a, B, st), "hg_ppo_loss_fwd_bwd")
        g = self._grad.data_ptr()
        self._side.wait_stream(cur)
        # The flat gradient buffer is [std | actor.* | critic.* | 8 loss statistics]; the critic's backward chain runs on the side
        # stream, and its range (+ the statistics, complete since the loss kernel) starts its all-reduce the moment that chain
        # ends -- behind the actor's backward -- while the actor's range follows its own chain: ONE logical reduction of the
        # buffer per optimizer step in two pieces, only the second of which is exposed.
        c0 = ac._offsets[next(k for k in ac._layout if k.startswith("critic."))]
        works = []
        with torch.cuda.stream(self._side):               # the two backward chains write disjoint gradient ranges
            if split:
                ac.native_backward_split("critic", xs_c, w["hid_c"], w["d_value"], w["dhid_c"], self._grad)
            else:
                nat.check(nat.lib.hg_mlp_backward(ac._desc["critic"], flat.data_ptr(), cobs.data_ptr(), cobs.stride(0), w["hid_c"].data_ptr(),
                                                  w["d_value"].data_ptr(), w["dhid_c"].data_ptr(), g, B,
                                                  nat.stream_ptr(self._dev_index)), "hg_mlp_backward(critic)")
            if world > 1:
                works.append(dist.all_reduce(self._grad[c0:], async_op=True))
        if split:
            ac.native_backward_split("actor", xs_a, w["hid_a"], w["d_mean"], w["dhid_a"], self._grad)
        else:
            nat.check(nat.lib.hg_mlp_backward(ac._desc["actor"], flat.data_ptr(), obs.data_ptr(), obs.stride(0), w["hid_a"].data_ptr(),
                                              w["d_mean"].data_ptr(), w["dhid_a"].data_ptr(), g, B, st), "hg_mlp_backward(actor)")
        if world > 1:
            works.append(dist.all_reduce(self._grad[:c0], async_op=True))
        cur.wait_stream(self._side)
        for wk in works:
            wk.wait()                                           # stream-side wait (no host block)
        if self.desired_kl is not None and self.schedule == "adaptive":
            nat.check(nat.lib.hg_adapt_lr(self._scalars.data_ptr() + 12, float(self.desired_kl), self._lr.data_ptr(), st), "hg_adapt_lr")
        n = ac.num_params
        nat.check(nat.lib.hg_grad_sqnorm(g, n, self._sqnorm.data_ptr(), st), "hg_grad_sqnorm")
        # the update's one-thread tail kernel also adds this step's 8 loss statistics to the sums update() reports
        nat.check(nat.lib.hg_clip_adam_step_stats(flat.data_ptr(), g, self._exp_avg.data_ptr(), self._exp_avg_sq.data_ptr(),
                                                  self._sqnorm.data_ptr(), float(self.max_grad_norm), self._lr.data_ptr(),
                                                  self._adam_step.data_ptr(), 0.9, 0.999, 1e-8, 1.0, n,
                                                  self._scalars.data_ptr(), self._loss_sums.data_ptr(), _TAIL, st),
                  "hg_clip_adam_step_stats")
        ac.invalidate_derived()
        if split:
            ac.refresh_split()                             # the next minibatch's GEMMs read the split image of the new weights

    def _permutation(self, n):
        """rollout_storage.py:155 `torch.randperm(n)`: one native launch (hg_randperm: keyed bijection + cycle walking, no
        sort), keyed by the rank-dependent seed and the number of updates so far.  HG_TORCH_RANDPERM=1 keeps torch's."""
        if os.environ.get("HG_TORCH_RANDPERM", "0") == "1":
            return torch.randperm(n, requires_grad=False, device=self.device)
        if getattr(self, "_perm", None) is None or self._perm.numel() != n:
            self._perm = torch.empty(n, dtype=torch.int64, device=self.device)
            self._perm_count = 0
        nat.check(nat.lib.hg_randperm(n, self._seed ^ 0x5045524D55544531, self._perm_count, self._perm.data_ptr(),
                                      nat.stream_ptr(self._dev_index)), "hg_randperm")
        self._perm_count += 1
        return self._perm

    def update(self):
        """ppo.py:119-184."""
        s = self.storage
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        batch_size = s.num_envs * s.num_transitions_per_env
        mini = batch_size // self.num_mini_batches
        indices = self._permutation(self.num_mini_batches * mini)
        self._loss_sums.zero_()
        split = self.use_split_path()
        if split:
            self.actor_critic.refresh_split()              # parameters may have been written from outside (load, tests)
        steps = [(ep, i) for ep in range(self.num_learning_epochs) for i in range(self.num_mini_batches)]
        if split:
            # software pipeline: minibatch k+1 is gathered (a pure HBM copy, 0.47 GB) on its own stream into the other buffer slot
            # while minibatch k runs its GEMM chains -- the gather's CTAs are small and share the SMs with the persistent GEMM CTAs
            cur = torch.cuda.current_stream(self._dev_index)
            if not hasattr(self, "_gather_stream"):
                self._gather_stream = torch.cuda.Stream(self._dev_index)
            gs = self._gather_stream
            gs.wait_stream(cur)
            done = [None, None]                                 # event: the minibatch step that last READ slot j has been enqueued

            def issue(k):
                i = steps[k][1]
                with torch.cuda.stream(gs):
                    if done[k % 2] is not None:
                        gs.wait_event(done[k % 2])
                    mb = s.gather(indices[i * mini:(i + 1) * mini], split=True, slot=k % 2)
                    ev = torch.cuda.Event()
                    ev.record(gs)
                return mb, ev
            nxt = issue(0)
            for k in range(len(steps)):
                mb, ready = nxt
                if k + 1 < len(steps):
                    nxt = issue(k + 1)
                cur.wait_event(ready)
                self.minibatch_step(mb, world)
                done[k % 2] = torch.cuda.Event()
                done[k % 2].record(cur)
            cur.wait_stream(gs)
        else:
            for _, i in steps:
                mb = s.gather(indices[i * mini:(i + 1) * mini], split=False)
                self.minibatch_step(mb, world)
        self.actor_critic.refresh_lo()                       # stand-alone 3xTF32 forwards (compute_returns) load weight-lo tiles by TMA
        if self.actor_critic.chain_f16:
            self.actor_critic.refresh_w16()                  # the next rollout's PPO.act chain reads the fp16x3 weight image
        num_updates = self.num_learning_epochs * self.num_mini_batches
        sums = self._loss_sums.tolist()                          # the only device->host read of the update
        s.clear()
        return sums[1] / num_updates, sums[0] / num_updates
