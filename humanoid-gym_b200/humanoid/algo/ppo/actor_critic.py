"""ActorCritic with the reference's constructor, module tree and state-dict keys
(reference algo/ppo/actor_critic.py:36-128), whose forward / backward run in libhg_b200.

All parameters are views into ONE flat fp32 buffer (order == named_parameters(): std, actor.*,
critic.*) so that the gradient all-reduce, the norm clip and Adam each touch a single range.
`self.actor` / `self.critic` stay real nn.Sequential(Linear, ELU, ...) modules: checkpoints
(`model_state_dict`) and export_policy_as_jit keep working unchanged.
"""
import os

import torch
import torch.nn as nn
from torch.distributions import Normal

from humanoid import _native as nat


def _mlp(inp, hidden, out, activation):
    dims = [inp] + list(hidden) + [out]
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(activation)
    return nn.Sequential(*layers), dims


class ActorCritic(nn.Module):
    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=[256, 256, 256],
                 critic_hidden_dims=[256, 256, 256], init_noise_std=1.0, activation=nn.ELU(), **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs.keys())))
        super().__init__()
        if not (isinstance(activation, nn.ELU) and activation.alpha == 1.0):
            raise nat.NativeError("the native MLP kernels implement ELU(alpha=1) only (XBotLCfgPPO default)")
        self.actor, self._actor_dims = _mlp(num_actor_obs, actor_hidden_dims, num_actions, activation)
        self.critic, self._critic_dims = _mlp(num_critic_obs, critic_hidden_dims, 1, activation)
        print(f"Actor MLP: {self.actor}")
        print(f"Critic MLP: {self.critic}")
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.num_actions = num_actions
        self.distribution = None
        Normal.set_default_validate_args = False
        self._flat = None
        self._scratch = {}

    # ------------------------------------------------------------------------------------------
    # flat parameter buffer + native descriptors
    # ------------------------------------------------------------------------------------------
    def _flatten(self):
        """(Re)alias every parameter into one contiguous buffer on its current device.

        Weight rows are laid out with a pitch rounded up to 8 elements and every block starts on a multiple of 8
        elements, so that TMA can address each matrix directly both in the fp32 buffer (3xTF32 path) and in its
        split bf16 image `_wsplit` (same element offsets; bf16x3 update path); the parameters themselves are
        (out, in) views of that storage, the pad floats stay zero for ever (zero gradient)."""
        params = list(self.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise nat.NativeError("ActorCritic must live on a CUDA device: no CPU fallback for the hot path")
        self._layout = {}
        off = 0
        for name, p in self.named_parameters():
            if p.dim() == 2:
                rows, cols = p.shape
                ld = (cols + 7) // 8 * 8
                self._layout[name] = (off, (rows, cols), ld)
                off += rows * ld
            else:
                self._layout[name] = (off, tuple(p.shape), None)
                off += (p.numel() + 7) // 8 * 8
        n = off
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        for name, p in self.named_parameters():
            v = self.view_of(flat, name)
            v.copy_(p.data)
            p.data = v
        self._offsets = {k: v[0] for k, v in self._layout.items()}
        self._flat = flat
        self._wsplit = torch.zeros(2, 1, n, dtype=torch.int16, device=dev)     # split bf16 image of the flat buffer (hi, lo planes)
        self._wsplit_dirty = True
        self._wlo = torch.zeros(n, dtype=torch.float32, device=dev)            # tf32 residuals of the flat buffer (3xTF32 "lo" operand)
        self._wlo_dirty = True
        self._w16 = torch.zeros(2, 1, n, dtype=torch.int16, device=dev)        # fp16 hi / lo planes of flat * 2^10 (fp16x3 rollout chain)
        self._w16_dirty = True
        self.num_params = n                       # allocated floats (incl. pads) == length of every flat buffer
        self.num_real_params = sum(p.numel() for p in params)
        self._desc = {}
        for prefix, dims in (("actor", self._actor_dims), ("critic", self._critic_dims)):
            d = nat.MlpDesc()
            d.n_layers = len(dims) - 1
            for i, w in enumerate(dims):
                d.dims[i] = w
            for l in range(d.n_layers):
                d.w_off[l], _, d.ldw[l] = self._layout[f"{prefix}.{2 * l}.weight"]
                d.b_off[l] = self._layout[f"{prefix}.{2 * l}.bias"][0]
            self._desc[prefix] = d
        self._scratch = {}

    def view_of(self, flat, name):
        """The (possibly row-padded) view of parameter `name` inside a flat buffer with this module's layout."""
        off, shape, ld = self._layout[name]
        if ld is None:
            k = 1
            for s_ in shape:
                k *= s_
            return flat[off:off + k].view(shape)
        rows, cols = shape
        return flat[off:off + rows * ld].view(rows, ld)[:, :cols]

    def flat_params(self):
        first = next(self.parameters())
        if self._flat is None or first.data_ptr() != self._flat.data_ptr() or self._flat.device != first.device:
            self._flatten()
        return self._flat

    # ---- split-precision (bf16x3) update path -----------------------------------------------------
    def split_eligible(self):
        """hg_mlp_*_split needs >= 2 layers, hidden widths % 8 == 0 and an output layer <= 16 wide."""
        self.flat_params()
        ok = True
        for dims in (self._actor_dims, self._critic_dims):
            ok = ok and len(dims) >= 3 and dims[-1] <= 16 and all(w % 8 == 0 for w in dims[1:-1])
        return ok

    def refresh_split(self):
        """Re-split the whole flat parameter buffer into its bf16 hi / lo image (one launch, 3.7 MB): after every
        optimizer step, a checkpoint load or any direct write to the parameters."""
        flat = self.flat_params()
        sp = nat.Split.of(self._wsplit)
        nat.check(nat.lib.hg_split_bf16(flat.data_ptr(), flat.numel(), sp, 1, flat.numel(), nat.stream_ptr(flat.device.index)),
                  "hg_split_bf16(params)")
        self._wsplit_dirty = False

    def native_forward_split(self, which, x_split, out, hidden):
        """out (M, dims[-1]) fp32 <- MLP_which(x) on split tensors; x_split: nat.Split of the (M, K) input;
        hidden: int16 scratch of 2 * M * hidden_width(which) elements (every hidden activation, split)."""
        flat = self.flat_params()
        if self._wsplit_dirty:
            self.refresh_split()
        M = out.shape[0]
        nat.check(nat.lib.hg_mlp_forward_split(self._desc[which], flat.data_ptr(), self._wsplit.data_ptr(), self._wsplit.stride(0),
                                               x_split, hidden.data_ptr(), out.data_ptr(), M, nat.stream_ptr(flat.device.index)),
                  "hg_mlp_forward_split")

    def native_backward_split(self, which, x_split, hidden, d_out, dhidden, grads):
        flat = self.flat_params()
        M = d_out.shape[0]
        nat.check(nat.lib.hg_mlp_backward_split(self._desc[which], flat.data_ptr(), self._wsplit.data_ptr(), self._wsplit.stride(0),
                                                x_split, hidden.data_ptr(), d_out.data_ptr(), dhidden.data_ptr(), grads.data_ptr(),
                                                M, nat.stream_ptr(flat.device.index)), "hg_mlp_backward_split")

    def hidden_width(self, which):
        dims = self._actor_dims if which.startswith("actor") else self._critic_dims
        return sum(dims[1:-1])

    def _hidden_scratch(self, which, M):
        key = (which, M)
        if key not in self._scratch:
            self._scratch[key] = torch.empty(M * self.hidden_width(which), dtype=torch.float32, device=self._flat.device)
        return self._scratch[key]

    def invalidate_derived(self):
        """Call after writing parameters behind this module's back (p.data.copy_, optimizer steps of your own): the
        split images of the weights (bf16 hi / lo planes, tf32 residuals) are recomputed on next use."""
        self._wsplit_dirty = True
        self._wlo_dirty = True
        self._w16_dirty = True

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate_derived()
        return r

    def refresh_lo(self):
        """tf32 residuals of the whole flat buffer (one launch, 3.7 MB): the rollout's GEMMs then load weight-lo tiles by
        TMA and their splitter warps handle the activations only."""
        flat = self.flat_params()
        nat.check(nat.lib.hg_tf32_residual(flat.data_ptr(), self._wlo.data_ptr(), flat.numel(), nat.stream_ptr(flat.device.index)),
                  "hg_tf32_residual")
        self._wlo_dirty = False

    def refresh_w16(self):
        """fp16 hi / lo planes of the flat buffer scaled by hg_f16_weight_scale() (one launch): operand B of the fp16x3 rollout
        chain (hg_actor_critic_forward_f16)."""
        flat = self.flat_params()
        nat.check(nat.lib.hg_split_f16(flat.data_ptr(), flat.numel(), nat.Split.of(self._w16), 1, flat.numel(),
                                       nat.lib.hg_f16_weight_scale(), nat.stream_ptr(flat.device.index)), "hg_split_f16(params)")
        self._w16_dirty = False

    def refresh_rollout_weights(self):
        """Derived weight images the rollout reads (after an optimizer step / checkpoint load)."""
        if self.chain_f16:
            self.refresh_w16()
        else:
            self.refresh_lo()

    @property
    def chain_f16(self):
        """fp16x3 operands for the one-launch PPO.act chain (default); HG_CHAIN_F16=0 selects the 3xTF32 chain on fp32 tiles."""
        return os.environ.get("HG_CHAIN_F16", "1") != "0"

    def native_forward(self, which, x, out, hidden=None, sample=None):
        """out (M, dims[-1]) <- MLP_which(x); returns the hidden-activation scratch.
        sample: optional dict(std, eps, actions, log_prob, sigma, seed, step, step_dev) -> PPO.act fused into the output
        layer's epilogue (out receives the mean)."""
        flat = self.flat_params()
        M = x.shape[0]
        assert x.dtype == torch.float32 and x.stride(1) == 1
        if hidden is None:
            hidden = self._hidden_scratch(which, M)
        if self._wlo_dirty and not torch.cuda.is_current_stream_capturing():
            self.refresh_lo()
        o = nat.MlpFwdOpts()
        o.params_lo = None if self._wlo_dirty else self._wlo.data_ptr()
        if sample is not None:
            o.std, o.eps = sample["std"].data_ptr(), nat.ptr(sample.get("eps"))
            o.actions, o.log_prob, o.sigma = sample["actions"].data_ptr(), sample["log_prob"].data_ptr(), sample["sigma"].data_ptr()
            o.seed, o.step, o.step_dev = sample["seed"], sample["step"], sample.get("step_dev")
        nat.check(nat.lib.hg_mlp_forward_ex(self._desc[which], flat.data_ptr(), x.data_ptr(), x.stride(0),
                                            hidden.data_ptr(), out.data_ptr(), M, o, nat.stream_ptr(flat.device.index)),
                  "hg_mlp_forward_ex")
        return hidden

    def chain_eligible(self, *inputs):
        """hg_actor_critic_forward applies: a tensor-core engine is selected, <= 8 layers, TMA-addressable inputs."""
        if nat.lib.hg_set_gemm_mode(-1) not in (1, 4) or os.environ.get("HG_FUSED_ACT", "1") == "0":
            return False
        for x in inputs:
            if x.dtype != torch.float32 or x.stride(1) != 1 or x.stride(0) % 4 or x.data_ptr() % 16:
                return False
        if len(self._actor_dims) + len(self._critic_dims) - 2 > 8 or self.num_actions > 32:
            return False
        if self.chain_f16:
            if self._w16_dirty:
                if torch.cuda.is_current_stream_capturing():
                    return False
                self.refresh_w16()
        elif self._wlo_dirty:
            if torch.cuda.is_current_stream_capturing():
                return False
            self.refresh_lo()
        return True

    def native_chain(self, which, x, out, sample=None):
        """One persistent multi-layer launch (hg_actor_critic_forward) for the actor ("actor": out = mean, optional sampling
        epilogue, see native_forward), the critic ("critic": out = value) or both ("both": x = (obs, critic_obs),
        out = (mean, value)).  The caller checks chain_eligible() first."""
        flat = self.flat_params()
        obs, cobs = (x if which == "both" else ((x, None) if which == "actor" else (None, x)))
        mu, value = (out if which == "both" else ((out, None) if which == "actor" else (None, out)))
        M = (obs if obs is not None else cobs).shape[0]
        key = ("chain_counters", which, M)
        if key not in self._scratch:
            self._scratch[key] = torch.zeros(int(nat.lib.hg_actor_critic_counters_size(M)), dtype=torch.int32, device=flat.device)
        o = nat.MlpFwdOpts()
        if sample is not None:
            o.std, o.eps = sample["std"].data_ptr(), nat.ptr(sample.get("eps"))
            o.actions, o.log_prob, o.sigma = sample["actions"].data_ptr(), sample["log_prob"].data_ptr(), sample["sigma"].data_ptr()
            o.seed, o.step, o.step_dev = sample["seed"], sample["step"], sample.get("step_dev")
        a_on, c_on = obs is not None, cobs is not None
        P = nat.ptr
        if self.chain_f16:
            da, dc = (self._desc["actor"] if a_on else None), (self._desc["critic"] if c_on else None)
            skey = ("chain16_scratch", which, M)
            if skey not in self._scratch:
                self._scratch[skey] = torch.zeros(int(nat.lib.hg_actor_critic_f16_scratch_elems(da, dc, M)), dtype=torch.int16, device=flat.device)
            nat.check(nat.lib.hg_actor_critic_forward_f16(da, dc, flat.data_ptr(), self._w16.data_ptr(), self._w16.stride(0),
                                                          obs.data_ptr() if a_on else None, obs.stride(0) if a_on else 0,
                                                          cobs.data_ptr() if c_on else None, cobs.stride(0) if c_on else 0,
                                                          self._scratch[skey].data_ptr(), mu.data_ptr() if a_on else None,
                                                          value.data_ptr() if c_on else None, o, self._scratch[key].data_ptr(), M,
                                                          nat.stream_ptr(flat.device.index)), "hg_actor_critic_forward_f16")
            return
        ha = self._hidden_scratch("actor", M) if a_on else None
        la = self._hidden_scratch("actor_lo", M) if a_on else None
        hc = self._hidden_scratch("critic", M) if c_on else None
        lc = self._hidden_scratch("critic_lo", M) if c_on else None
        nat.check(nat.lib.hg_actor_critic_forward(self._desc["actor"] if a_on else None, self._desc["critic"] if c_on else None,
                                                  flat.data_ptr(), self._wlo.data_ptr(),
                                                  obs.data_ptr() if a_on else None, obs.stride(0) if a_on else 0,
                                                  cobs.data_ptr() if c_on else None, cobs.stride(0) if c_on else 0,
                                                  P(ha), P(hc), P(la), P(lc), mu.data_ptr() if a_on else None,
                                                  value.data_ptr() if c_on else None, o, self._scratch[key].data_ptr(), M,
                                                  nat.stream_ptr(flat.device.index)), "hg_actor_critic_forward")

    def native_act(self, obs, critic_obs, mu, value, sample=None):
        """Both nets in one launch (kept for tests / callers without a side stream)."""
        if not self.chain_eligible(obs, critic_obs):
            return False
        self.native_chain("both", (obs, critic_obs), (mu, value), sample)
        return True

    # ------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------
    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def _actor_mean(self, observations):
        obs = observations.to(torch.float32).contiguous()
        mean = torch.empty(obs.shape[0], self.num_actions, dtype=torch.float32, device=obs.device)
        self.native_forward("actor", obs, mean)
        return mean

    def update_distribution(self, observations):
        mean = self._actor_mean(observations)
        self.distribution = Normal(mean, mean * 0. + self.std.detach())

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        return self._actor_mean(observations)

    def evaluate(self, critic_observations, **kwargs):
        cobs = critic_observations.to(torch.float32).contiguous()
        value = torch.empty(cobs.shape[0], 1, dtype=torch.float32, device=cobs.device)
        self.native_forward("critic", cobs, value)
        return value
