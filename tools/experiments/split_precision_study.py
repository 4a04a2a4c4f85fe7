"""Numerics of the operand-split schemes considered for the ActorCritic GEMMs (CPU emulation, fp32 accumulation).

  3xTF32      a_hi = trunc_tf32(a), a_lo = rna_tf32(a - a_hi);  D = a_lo b_hi + a_hi b_lo + a_hi b_hi      (shipped)
  TF32+bf16   same hi terms, the two correction products with bf16-rounded operands                          (2 passes)
  bf16x3      a_hi = bf16(a), a_lo = bf16(a - a_hi);            D = a_hi b_hi + a_hi b_lo + a_lo b_hi        (1.5 passes)
  1xTF32, 1xbf16 for scale.

Relative Frobenius error against an fp64 product, for the three product shapes of one layer
(forward X W^T, dgrad dZ W, wgrad dZ^T X) at a reduced batch.  Usage: python tools/experiments/split_precision_study.py [B]"""
import sys

import torch


def trunc_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def rna_tf32(x):
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def mm(a, b):
    return a @ b           # fp32 accumulate (CPU sgemm)


def schemes(a, b):
    out = {}
    ah, bh = trunc_tf32(a), trunc_tf32(b)
    al, bl = rna_tf32(a - ah), rna_tf32(b - bh)
    out["1xTF32"] = mm(ah, bh)
    out["3xTF32"] = mm(al, bh) + mm(ah, bl) + mm(ah, bh)
    out["TF32+bf16 corr"] = mm(bf16(al), bf16(bh)) + mm(bf16(ah), bf16(bl)) + mm(ah, bh)
    a1, b1 = bf16(a), bf16(b)
    a2, b2 = bf16(a - a1), bf16(b - b1)
    out["1xbf16"] = mm(a1, b1)
    out["bf16x3"] = mm(a1, b1) + mm(a1, b2) + mm(a2, b1)
    out["fp32"] = mm(a, b)
    return out


def rel(x, ref):
    return float((x.double() - ref).norm() / ref.norm())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    g = torch.Generator().manual_seed(0)
    K, N = 705, 512
    X = torch.randn(B, K, generator=g).clamp(-18, 18)
    W = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    dZ = torch.randn(B, N, generator=g) * 1e-3 * (torch.rand(B, N, generator=g) > 0.3)
    cases = {"forward  X W^T   (K=705)": (X, W.t().contiguous()),
             "dgrad    dZ W    (K=512)": (dZ, W),
             f"wgrad    dZ^T X  (K={B})": (dZ.t().contiguous(), X)}
    names = ["1xbf16", "1xTF32", "bf16x3", "TF32+bf16 corr", "3xTF32", "fp32"]
    print(f"{'product':28s}" + "".join(f"{n:>16s}" for n in names))
    for title, (a, b) in cases.items():
        ref = a.double() @ b.double()
        r = schemes(a, b)
        print(f"{title:28s}" + "".join(f"{rel(r[n], ref):16.2e}" for n in names))


if __name__ == "__main__":
    main()
