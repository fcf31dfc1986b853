"""Independent NumPy model of the CUDA Philox draw stream (test helper).

Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw, SC'11; Random123), plus the stream
layout of pldepth_b200/csrc/pld_common.cuh: counter = (list, image, word_block | offset_hi<<16,
offset_lo), key = (seed_lo, seed_hi); draw k = word k&3 of block k>>2, mapped to [0, M) with
Lemire's unbiased multiply-shift; the a-th redraw of draw k is word 0 of block
0x8000 | ((a-1)&63) << 9 | k.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def draw_selection(seed, offset, image, n, K, M):
    """Selections [n, K] (int64) for one image with M valid pixels."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    off_lo, off_hi = offset & 0xFFFFFFFF, ((offset >> 32) & 0xFFFF) << 16
    lists = np.arange(n, dtype=np.uint32)
    nblk = (K + 3) // 4
    words = np.empty((n, nblk * 4), dtype=np.uint32)
    for q in range(nblk):
        r = philox4x32_10(lists, np.uint32(image), np.uint32(q | off_hi), np.uint32(off_lo), k0, k1)
        for j in range(4):
            words[:, q * 4 + j] = r[j]
    words = words[:, :K]
    thresh = ((1 << 32) - M) % M
    m = words.astype(np.uint64) * np.uint64(M)
    sel = (m >> np.uint64(32)).astype(np.int64)
    low = (m & MASK32).astype(np.uint64)
    bad = np.argwhere(low < thresh)
    for l, k in bad:                      # rare redraw path
        a = 0
        lo = int(low[l, k])
        while lo < thresh:
            a += 1
            blk = 0x8000 | (((a - 1) & 63) << 9) | int(k)
            r = philox4x32_10(np.uint32(l), np.uint32(image), np.uint32(blk | off_hi), np.uint32(off_lo), k0, k1)
            mm = int(r[0]) * M
            lo = mm & 0xFFFFFFFF
            sel[l, k] = mm >> 32
    return sel
