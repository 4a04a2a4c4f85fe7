"""Does bf16x3 (or TF32 + bf16 corrections) hold the parity bars at NETWORK level?  (CPU emulation)

Runs the XBot-L actor and critic (705-512-256-128-12 / 219-768-256-128-1, ELU) forward and backward on one batch with
every GEMM replaced by an emulated split product (fp32 accumulation), and compares
  * the outputs with an fp64 forward (bar: 1e-5 relative),
  * every parameter gradient with an fp64 backward of the same upstream gradient (bar: 1e-4 relative L2 per tensor).
Usage: python tools/experiments/network_precision_study.py [B]"""
import sys

import torch

from split_precision_study import bf16, rna_tf32, trunc_tf32


def mm_fp32(a, b):
    return a @ b


def mm_3xtf32(a, b):
    ah, bh = trunc_tf32(a), trunc_tf32(b)
    al, bl = rna_tf32(a - ah), rna_tf32(b - bh)
    return al @ bh + ah @ bl + ah @ bh


def mm_bf16x3(a, b):
    a1, b1 = bf16(a), bf16(b)
    a2, b2 = bf16(a - a1), bf16(b - b1)
    return a2 @ b1 + a1 @ b2 + a1 @ b1


def mm_tf32_bf16corr(a, b):
    ah, bh = trunc_tf32(a), trunc_tf32(b)
    al, bl = rna_tf32(a - ah), rna_tf32(b - bh)
    return bf16(al) @ bf16(bh) + bf16(ah) @ bf16(bl) + ah @ bh


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def run(dims, X, dY, Ws, bs, mm, dtype=torch.float32):
    X, dY = X.to(dtype), dY.to(dtype)
    Ws, bs = [w.to(dtype) for w in Ws], [b.to(dtype) for b in bs]
    hs = [X]
    for l, (W, b) in enumerate(zip(Ws, bs)):
        z = mm(hs[-1], W.t().contiguous()) + b
        hs.append(z if l == len(Ws) - 1 else elu(z))
    out = hs[-1]
    grads, dZ = [], dY
    for l in reversed(range(len(Ws))):
        grads.append((mm(dZ.t().contiguous(), hs[l]), dZ.sum(0)))
        if l > 0:
            dH = mm(dZ, Ws[l])
            h = hs[l]
            dZ = dH * torch.where(h > 0, torch.ones_like(h), h + 1)
    return out, grads[::-1]


def rel(x, ref):
    return float((x.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-300))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    g = torch.Generator().manual_seed(1)
    nets = {"actor": [705, 512, 256, 128, 12], "critic": [219, 768, 256, 128, 1]}
    schemes = {"fp32": mm_fp32, "3xTF32": mm_3xtf32, "TF32+bf16corr": mm_tf32_bf16corr, "bf16x3": mm_bf16x3}
    for name, dims in nets.items():
        Ws = [(torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5 for i, o in zip(dims[:-1], dims[1:])]
        bs = [(torch.rand(o, generator=g) * 2 - 1) / i ** 0.5 for i, o in zip(dims[:-1], dims[1:])]
        X = torch.randn(B, dims[0], generator=g).clamp(-18, 18)
        dY = torch.randn(B, dims[-1], generator=g) / B
        ref_out, ref_g = run(dims, X, dY, Ws, bs, mm_fp32, torch.float64)
        print(f"\n{name}  (B={B})")
        print(f"{'scheme':16s}{'output':>11s}" + "".join(f"{'dW' + str(l):>11s}" for l in range(len(Ws))) + f"{'worst db':>11s}")
        for sn, mm in schemes.items():
            out, gr = run(dims, X, dY, Ws, bs, mm)
            row = f"{sn:16s}{rel(out, ref_out):11.2e}" + "".join(f"{rel(gw, rw):11.2e}" for (gw, _), (rw, _) in zip(gr, ref_g))
            row += f"{max(rel(gb, rb) for (_, gb), (_, rb) in zip(gr, ref_g)):11.2e}"
            print(row)


if __name__ == "__main__":
    main()
