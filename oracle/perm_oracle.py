"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of hg_randperm (csrc/hg_ppo.cu), the native
minibatch permutation that replaces `torch.randperm` in RolloutStorage.mini_batch_generator (reference
algo/ppo/rollout_storage.py:155).

The reference's permutation comes from torch's generator; any permutation is a valid stand-in (the reference's own
result changes with its seed), so what is pinned here is the KERNEL: integer work, bit-exact.  Only tests/ may import
this module.

Algorithm: 8 round keys from two Philox4x32-10 blocks keyed by (seed, counter); an alternating unbalanced Feistel
network over the enclosing power-of-two range [0, 2^b) with the murmur3 finalizer as round function; cycle walking
(re-apply until the value is < n)."""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed, c0, c1, c2, c3):
    """Salmon et al. 2011, as hg_philox in csrc/hg_common.cuh: key = (seed lo, seed hi), 10 rounds."""
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    c0, c1, c2, c3 = (int(x) & 0xFFFFFFFF for x in (c0, c1, c2, c3))
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & 0xFFFFFFFF, p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _fmix32(h):
    h = h & _M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h


def randperm(n, seed, counter):
    bits = 1
    while (1 << bits) < n:
        bits += 1
    bits = max(bits, 2)
    bl = bits // 2
    bh = bits - bl
    ml, mh = np.uint64((1 << bl) - 1), np.uint64((1 << bh) - 1)
    lo32, hi32 = int(counter) & 0xFFFFFFFF, (int(counter) >> 32) & 0xFFFFFFFF
    key = [np.uint64(k) for blk in (0, 1) for k in philox4x32_10(seed, lo32, hi32, 0x50455246, blk)]
    x = np.arange(n, dtype=np.uint64)
    todo = np.ones(n, dtype=bool)
    while todo.any():
        v = x[todo]
        lo, hi = v & ml, (v >> np.uint64(bl)) & mh
        for r in range(0, 8, 2):
            hi = hi ^ (_fmix32(lo ^ key[r]) & mh)
            lo = lo ^ (_fmix32(hi ^ key[r + 1]) & ml)
        v = (hi << np.uint64(bl)) | lo
        x[todo] = v
        todo[todo] = v >= np.uint64(n)
    return x.astype(np.int64)
