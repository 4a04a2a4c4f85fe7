"""Network-level accuracy of the PPO.act engines against an fp64 evaluation of the same weights: the fp16x3 chain (default),
the 3xTF32 chain and the per-layer 3xTF32 launches, at the flagship shapes (705-512-256-128-12 / 219-768-256-128-1, M = 4096)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "humanoid-gym_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from humanoid.algo import ActorCritic  # noqa: E402

torch.manual_seed(0)
ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128]).cuda()
ac.flat_params()
M = 4096


def pad4(t):                                          # row pitch % 4 == 0: TMA-addressable (what the env / storage buffers have)
    buf = torch.zeros(t.shape[0], (t.shape[1] + 3) // 4 * 4, device=t.device)
    buf[:, :t.shape[1]] = t
    return buf[:, :t.shape[1]]


obs = torch.randn(M, 705, device="cuda").clamp(-18, 18)
obs[:, ::5] *= 1e-3                                   # small-magnitude channels (fp16 lo planes go subnormal there)
obs, cobs = pad4(obs), pad4(torch.randn(M, 219, device="cuda") * 3.0)


def ref(x, prefix):
    h = x.double()
    layers = [m for m in getattr(ac, prefix) if isinstance(m, torch.nn.Linear)]
    for i, m in enumerate(layers):
        h = h @ m.weight.double().t() + m.bias.double()
        if i + 1 < len(layers):
            h = torch.nn.functional.elu(h)
    return h


r_mu, r_v = ref(obs, "actor"), ref(cobs, "critic")


def err(a, b):
    return float((a.double() - b).norm() / b.norm()), float((a.double() - b).abs().max() / b.abs().max())


from humanoid import _native as nat  # noqa: E402
for label, env in (("exact-fp32 CUDA cores", {"HG_FUSED_ACT": "0", "simt": "1"}), ("fp16x3 chain", {"HG_CHAIN_F16": "1"}), ("3xTF32 chain", {"HG_CHAIN_F16": "0"}), ("3xTF32 per-layer launches", {"HG_FUSED_ACT": "0"})):
    os.environ.update({"HG_CHAIN_F16": "1", "HG_FUSED_ACT": "1"})
    os.environ.update({k: v for k, v in env.items() if k != "simt"})
    prev = nat.lib.hg_set_gemm_mode(0 if "simt" in env else 4)
    mu, v = torch.zeros(M, 12, device="cuda"), torch.zeros(M, 1, device="cuda")
    if not ac.native_act(obs, cobs, mu, v):
        ac.native_forward("actor", obs, mu)
        ac.native_forward("critic", cobs, v)
    torch.cuda.synchronize()
    nat.lib.hg_set_gemm_mode(prev)
    print(f"{label:28s} mean: rel-norm {err(mu, r_mu)[0]:.2e} max/max {err(mu, r_mu)[1]:.2e} | value: rel-norm {err(v, r_v)[0]:.2e} max/max {err(v, r_v)[1]:.2e}")
