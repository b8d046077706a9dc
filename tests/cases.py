"""Shared input builders for the parity tests (seeded; identical here and on the GPU box)."""
import importlib
import hashlib

import numpy as np

import pyoracle
from pyoracle import TYPES

keygen = importlib.import_module("radix-sorting_b200.keygen")


def make_input(tname: str, n: int, seed: int, dist: str = "uniform", mask: int = (1 << 64) - 1,
               orv: int = 0) -> np.ndarray:
    """n elements of TYPES[tname]; scalar types take the key stream's bits verbatim (the way
    `./radix ... float` reinterprets the key file, radix_experiment.cpp:264-279); records get
    key = stream, payload = original position."""
    t = TYPES[tname]
    keys = keygen.fill(seed, 0, n, t.key_bytes, dist, mask, orv)
    if t.dtype.names is None:
        return keys.view(t.dtype).copy()
    out = np.zeros(n, dtype=t.dtype)
    out["key"] = keys.view(t.dtype["key"])
    second = [f for f in t.dtype.names if f not in ("key", "pad")][0]
    out[second] = np.arange(n, dtype=np.uint64).astype(t.dtype[second])
    return out


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# (type, n, dist, mask, orv) -- the matrix the golden file and the GPU parity tests share.
_M64 = (1 << 64) - 1
SIZES = [2, 3, 31, 255, 256, 257, 1000, 4097, 65535, 65536, 70001]
GOLDEN_CASES = []
for _t in ["u8", "u16", "u32", "u64", "i8", "i16", "i32", "i64", "f32", "f64"]:
    for _n in SIZES:
        GOLDEN_CASES.append((_t, _n, "uniform", _M64, 0))
for _t in ["rec16_u8", "rec8_u32", "rec16_u64"]:
    for _n in [2, 257, 4097, 70001]:
        GOLDEN_CASES.append((_t, _n, "uniform", _M64, 0))
# column skipping: the README's own masks (README.md:889-891) and constant-high-byte inputs (C2c/C2d)
for _t, _m, _o in [("u32", 0x00FFFFFF, 0), ("u32", 0x0000FFFF, 0), ("u32", 0x00FF00FF, 0), ("u32", 0xFF, 0),
                   ("u64", 0x000000FFFFFFFFFF, 0xAA00000000000000), ("u64", 0xFFFFFFFF, 0),
                   ("u64", 0x0000FFFFFFFFFFFF, 0), ("i32", 0x0000FFFF, 0), ("f32", 0x7FFFFF00, 0),
                   ("rec8_u32", 0x000FFFFF, 0)]:
    for _n in [1000, 70001]:
        GOLDEN_CASES.append((_t, _n, "uniform", _m, _o))
for _t in ["u32", "u64", "i64", "f32", "rec8_u32"]:
    for _d in ["sorted", "reverse", "constant", "and3", "zipf"]:
        GOLDEN_CASES.append((_t, 5000, _d, _M64, 0))


def case_id(c) -> str:
    t, n, dist, mask, orv = c
    s = f"{t}-n{n}-{dist}"
    if mask != _M64:
        s += f"-m{mask:x}"
    if orv:
        s += f"-o{orv:x}"
    return s
