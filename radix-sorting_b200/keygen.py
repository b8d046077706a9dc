"""Seeded synthetic key streams.

Counter-based: element i of a stream is a pure function of (seed, i), so the host (numpy,
here), the device generator (csrc/rsx_keygen.cu, `rsx_fill_keys`) and every shard of a
multi-GPU run produce identical bytes without sharing state.  The format is the reference's
key-file format: a raw little-endian array with no header (`40M_32bit_keys.dat`, reference
Makefile:79-82) -- that file itself is /dev/urandom output and cannot be reproduced, hence
the seed.

    word(seed, i) = splitmix64_finalise(seed * 0x9E3779B97F4A7C15 + i)

Distributions (`dist`):
    "uniform"   word
    "and2/3/4"  bitwise AND of 2/3/4 independent words (bit density 1/4, 1/8, 1/16): skewed
                digits without a constant column (config C2e)
    "zipf"      floor(2 ** (u * 32)) with u uniform in [0,1): P(v) ~ 1/v (config C5)
    "sorted"    i itself (presorted), "reverse"  (2^64-1 - i), "constant"  word(seed, 0)
`mask`/`orv` are applied afterwards (`(x & mask) | orv`, configs C2c/C2d); keys narrower
than 64 bits take the low bits.
"""
from __future__ import annotations

import numpy as np

GOLDEN = 0x9E3779B97F4A7C15
DISTS = ("uniform", "and2", "and3", "and4", "zipf", "sorted", "reverse", "constant")
DIST_CODE = {d: i for i, d in enumerate(DISTS)}
_M64 = (1 << 64) - 1


def _mix64(z: np.ndarray) -> np.ndarray:
    z = z.copy()
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def words(seed: int, start: int, count: int, stream: int = 0) -> np.ndarray:
    """`count` 64-bit words of stream (seed, stream) starting at element index `start`."""
    with np.errstate(over="ignore"):
        base = np.uint64(((seed + 0x632BE59BD9B4E019 * stream) * GOLDEN) & _M64)
        i = np.arange(start, start + count, dtype=np.uint64)
        return _mix64(base + i)


def fill(seed: int, start: int, count: int, key_bytes: int, dist: str = "uniform",
         mask: int = _M64, orv: int = 0) -> np.ndarray:
    """Unsigned little-endian keys of `key_bytes` bytes for indices [start, start+count)."""
    if dist == "uniform":
        x = words(seed, start, count)
    elif dist in ("and2", "and3", "and4"):
        x = words(seed, start, count)
        for s in range(1, int(dist[3])):
            x &= words(seed, start, count, stream=s)
    elif dist == "zipf":
        # u = top 32 bits / 2^32; value = floor(2^(32u)) computed in integers:
        # 2^(32u) = 2^e * 2^f with e = floor(32u); the fraction is approximated by the
        # linear term (1 + f), which keeps P(v) ~ 1/v per octave and is exactly reproducible.
        w = words(seed, start, count)
        hi = w >> np.uint64(32)                      # 32-bit uniform
        e = hi >> np.uint64(27)                      # 0..31
        f = hi & np.uint64((1 << 27) - 1)            # 27-bit fraction
        x = ((np.uint64(1) << np.uint64(27)) + f) << e >> np.uint64(27)
    elif dist == "sorted":
        x = np.arange(start, start + count, dtype=np.uint64)
    elif dist == "reverse":
        x = np.uint64(_M64) - np.arange(start, start + count, dtype=np.uint64)
    elif dist == "constant":
        x = np.full(count, words(seed, 0, 1)[0], dtype=np.uint64)
    else:
        raise ValueError(f"unknown dist {dist!r}")
    x = (x & np.uint64(mask & _M64)) | np.uint64(orv & _M64)
    return x.astype(np.dtype(f"<u{key_bytes}"))
