"""RolloutStorage: (T, N, .) transition slabs, GAE and minibatch gather
(reference algo/ppo/rollout_storage.py:35-182), each a single native launch."""
import torch

from humanoid import _native as nat


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device="cpu"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise nat.NativeError("RolloutStorage lives in HBM: pass a CUDA device (no CPU fallback)")
        self.device = device
        self.obs_shape, self.privileged_obs_shape, self.actions_shape = obs_shape, privileged_obs_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        z = dict(device=device)
        self.observations = torch.zeros(T, N, *obs_shape, **z)
        self.privileged_observations = torch.zeros(T, N, *privileged_obs_shape, **z) if privileged_obs_shape[0] is not None else None
        self.rewards = torch.zeros(T, N, 1, **z)
        self.actions = torch.zeros(T, N, *actions_shape, **z)
        self.dones = torch.zeros(T, N, 1, **z).byte()
        self.actions_log_prob = torch.zeros(T, N, 1, **z)
        self.values = torch.zeros(T, N, 1, **z)
        self.returns = torch.zeros(T, N, 1, **z)
        self.advantages = torch.zeros(T, N, 1, **z)
        self.mu = torch.zeros(T, N, *actions_shape, **z)
        self.sigma = torch.zeros(T, N, *actions_shape, **z)
        self.num_transitions_per_env, self.num_envs = T, N
        self.saved_hidden_states_a = self.saved_hidden_states_c = None
        self.step = 0
        self._stats = torch.zeros(4, dtype=torch.float64, device=device)
        self._dev_index = torch.device(device).index
        self._mb = None
        self._bind()

    def _bind(self):
        S = nat.Storage()
        for k in ("observations", "privileged_observations", "actions", "rewards", "dones", "values",
                  "actions_log_prob", "mu", "sigma", "returns", "advantages"):
            setattr(S, k, nat.ptr(getattr(self, k)))
        S.T = self.num_transitions_per_env
        S.num_obs = self.obs_shape[0]
        S.num_priv = self.privileged_obs_shape[0] or 0
        S.num_actions = self.actions_shape[0]
        self._S = S
        # compute_returns rebinds self.advantages in the reference; keep our buffer and remember its id
        self._bound_adv = self.advantages

    def _native(self):
        if self._bound_adv is not self.advantages:
            self._bind()
        return self._S

    # ---- add -----------------------------------------------------------------------------------
    def add_native(self, t, gamma=0.0, **tensors):
        """One launch: copy the given tensors into slab t (None / in-place ones are skipped)."""
        if t >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        tr = nat.Transition()
        for k, v in tensors.items():
            if v is not None:
                assert v.is_cuda
                if k in ("obs", "priv_obs"):            # row-pitched views (env buffers) are fine
                    assert v.dim() == 2 and v.stride(1) == 1
                    setattr(tr, "obs_pitch" if k == "obs" else "priv_pitch", v.stride(0))
                else:
                    assert v.is_contiguous()
                setattr(tr, k, v.data_ptr())
        nat.check(nat.lib.hg_storage_add(self._native(), tr, t, gamma, self.num_envs, nat.stream_ptr(self._dev_index)),
                  "hg_storage_add")

    def add_transitions(self, transition):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        f = lambda x: None if x is None else (x if (x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1)  # noqa: E731
                                              else x.to(torch.float32).contiguous())
        d = transition.dones
        self.add_native(
            self.step, obs=f(transition.observations),
            priv_obs=f(transition.critic_observations) if self.privileged_observations is not None else None,
            actions=f(transition.actions), rewards=f(transition.rewards.view(-1)),
            dones=d.view(-1).to(torch.uint8).contiguous() if d.dtype != torch.bool else d.view(-1).contiguous(),
            values=f(transition.values.view(-1)), log_prob=f(transition.actions_log_prob.view(-1)),
            mu=f(transition.action_mean), sigma=f(transition.action_sigma))
        self.step += 1

    def clear(self):
        self.step = 0

    # ---- returns -------------------------------------------------------------------------------
    def compute_returns(self, last_values, gamma, lam, normalise=True):
        lv = last_values.to(torch.float32).contiguous()
        nat.check(nat.lib.hg_gae(self._native(), lv.data_ptr(), gamma, lam, self._stats.data_ptr(), int(normalise),
                                 self.num_envs, nat.stream_ptr(self._dev_index)), "hg_gae")

    def normalise_advantages(self):
        nat.check(nat.lib.hg_adv_normalise(self._native(), self._stats.data_ptr(), self.num_envs,
                                           nat.stream_ptr(self._dev_index)), "hg_adv_normalise")

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        idx = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        return (idx[1:] - idx[:-1]).float().mean(), self.rewards.mean()

    # ---- minibatches -----------------------------------------------------------------------------
    def _minibatch_buffers(self, B):
        if self._mb is None or self._mb["obs"].shape[0] != B:
            z = dict(device=self.device, dtype=torch.float32)
            A = self.actions_shape[0]
            def padded(width):      # row pitch rounded up to 4 floats: TMA-addressable by the tensor-core MLP path
                return torch.zeros(B, (width + 3) // 4 * 4, **z)[:, :width]
            self._mb = dict(
                obs=padded(self.obs_shape[0]),
                priv_obs=padded(self.privileged_obs_shape[0]) if self.privileged_observations is not None else None,
                actions=torch.empty(B, A, **z), values=torch.empty(B, 1, **z), advantages=torch.empty(B, 1, **z),
                returns=torch.empty(B, 1, **z), old_log_prob=torch.empty(B, 1, **z), old_mu=torch.empty(B, A, **z),
                old_sigma=torch.empty(B, A, **z))
            m = nat.MiniBatch()
            for k, v in self._mb.items():
                setattr(m, k, None if v is None else v.data_ptr())
            m.ld_obs = self._mb["obs"].stride(0)
            m.ld_priv = self._mb["priv_obs"].stride(0) if self._mb["priv_obs"] is not None else 0
            self._mbs = m
        return self._mb

    def _split_buffers(self, B, slot=0):
        """Split (bf16 hi / lo planes, row pitch % 8) minibatch buffers for the bf16x3 update path: the observations are written
        ONLY in this form (same bytes as fp32, no extra traffic).  Two independent slots, so that PPO.update can gather
        minibatch i+1 on a side stream while minibatch i is being consumed."""
        key = (B, slot)
        if getattr(self, "_split_slots", None) is None:
            self._split_slots = {}
        if key not in self._split_slots:
            z = dict(device=self.device, dtype=torch.float32)
            A = self.actions_shape[0]

            def planes(width):
                return torch.zeros(2, B, (width + 7) // 8 * 8, dtype=torch.int16, device=self.device)
            t = dict(obs=None, priv_obs=None, obs_split=planes(self.obs_shape[0]),
                     priv_split=planes(self.privileged_obs_shape[0]) if self.privileged_observations is not None else None,
                     actions=torch.empty(B, A, **z), values=torch.empty(B, 1, **z), advantages=torch.empty(B, 1, **z),
                     returns=torch.empty(B, 1, **z), old_log_prob=torch.empty(B, 1, **z), old_mu=torch.empty(B, A, **z),
                     old_sigma=torch.empty(B, A, **z))
            m = nat.MiniBatch()
            for k in ("actions", "values", "advantages", "returns", "old_log_prob", "old_mu", "old_sigma"):
                setattr(m, k, t[k].data_ptr())
            m.obs = None
            m.priv_obs = None
            m.obs_split = nat.Split.of(t["obs_split"])
            if t["priv_split"] is not None:
                m.priv_split = nat.Split.of(t["priv_split"])
            if len(self._split_slots) >= 4:
                self._split_slots.clear()
            self._split_slots[key] = (t, m)
        return self._split_slots[key]

    def gather(self, batch_idx, split=False, slot=0):
        """Rows `batch_idx` of the flattened (T*N, .) storage -> contiguous minibatch tensors (split=True: the
        observations come out as split bf16 planes `obs_split` / `priv_split` instead of fp32), on the current stream."""
        B = batch_idx.numel()
        if split:
            mb, desc = self._split_buffers(B, slot)
        else:
            mb, desc = self._minibatch_buffers(B), None
            desc = self._mbs
        nat.check(nat.lib.hg_minibatch_gather(self._native(), batch_idx.data_ptr(), desc, B,
                                              nat.stream_ptr(self._dev_index)), "hg_minibatch_gather")
        return dict(mb)

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                mb = self.gather(indices[i * mini_batch_size:(i + 1) * mini_batch_size])
                cobs = mb["priv_obs"] if mb["priv_obs"] is not None else mb["obs"]
                yield (mb["obs"], cobs, mb["actions"], mb["values"], mb["advantages"], mb["returns"],
                       mb["old_log_prob"], mb["old_mu"], mb["old_sigma"], (None, None), None)
