"""ctypes bindings for the CPU checker (oracle/liboracle.so, oracle/_ref/libradix_ref.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.

* ``Oracle``  -- the C restatement of radix_sort.hpp / radix_sort_rank.hpp (rsx_oracle.c).
* ``Ref``     -- the unmodified reference headers behind ref_shim.cpp (only where
                 oracle/_ref/libradix_ref.so exists; it is built in the CPU container from
                 /root/reference and travels to the GPU box as a prebuilt file).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libradix_ref.so")

KDF_UNSIGNED, KDF_SIGNED, KDF_FLOAT = 0, 1, 2
FLAG_INVERT = 1


class OrcLayout(C.Structure):
    _fields_ = [("record_bytes", C.c_uint32), ("key_offset", C.c_uint32),
                ("key_bytes", C.c_uint32), ("kdf_kind", C.c_uint32), ("flags", C.c_uint32)]


class OrcReport(C.Structure):
    _fields_ = [("n_unsorted", C.c_uint64), ("early_exit", C.c_uint32), ("ncols", C.c_uint32),
                ("cols", C.c_uint32 * 8), ("result_in_aux", C.c_uint32)]


REC16_U8 = np.dtype([("key", "u1"), ("pad", "V7"), ("name", "<u8")])       # radix_tests.cpp:15-18
REC8_U32 = np.dtype([("key", "<u4"), ("payload", "<u4")])                   # config C4b
REC16_U64 = np.dtype([("key", "<u8"), ("payload", "<u8")])                  # u64 key + 8 B payload
# layouts the reference can express with a by-member KeyFunc but does not ship a test for
# (no ref_code: checked against the oracle's independent stable sort, not against ref_shim)
REC16_F64 = np.dtype([("payload", "<u8"), ("key", "<f8")])                  # double key in the 2nd word
REC16_F32 = np.dtype([("payload", "<u4"), ("key", "<f4"), ("pad", "V8")])   # float key at offset 4
# record sizes the tile kernels do not move directly (sorted as keys + indices, then one gather)
REC12_U32 = np.dtype([("key", "<u4"), ("payload", "<u8")])                  # radix_sort_u32.c:7-10 sortrec, packed: 12 bytes
REC24_F64 = np.dtype([("payload", "<u4"), ("pad", "V8"), ("key", "<f8"), ("tail", "<u4")])  # double key straddling 8-byte words
REC7_I16 = np.dtype([("payload", "<u4"), ("pad", "V1"), ("key", "<i2")])    # odd size, unaligned signed key


@dataclass(frozen=True)
class ElemType:
    name: str
    dtype: np.dtype
    record_bytes: int
    key_offset: int
    key_bytes: int
    kdf_kind: int
    ref_code: int

    def layout(self, descending: bool = False) -> OrcLayout:
        return OrcLayout(self.record_bytes, self.key_offset, self.key_bytes, self.kdf_kind,
                         FLAG_INVERT if descending else 0)


TYPES = {t.name: t for t in [
    ElemType("u8", np.dtype("u1"), 1, 0, 1, KDF_UNSIGNED, 0),
    ElemType("u16", np.dtype("<u2"), 2, 0, 2, KDF_UNSIGNED, 1),
    ElemType("u32", np.dtype("<u4"), 4, 0, 4, KDF_UNSIGNED, 2),
    ElemType("u64", np.dtype("<u8"), 8, 0, 8, KDF_UNSIGNED, 3),
    ElemType("i8", np.dtype("i1"), 1, 0, 1, KDF_SIGNED, 4),
    ElemType("i16", np.dtype("<i2"), 2, 0, 2, KDF_SIGNED, 5),
    ElemType("i32", np.dtype("<i4"), 4, 0, 4, KDF_SIGNED, 6),
    ElemType("i64", np.dtype("<i8"), 8, 0, 8, KDF_SIGNED, 7),
    ElemType("f32", np.dtype("<f4"), 4, 0, 4, KDF_FLOAT, 8),
    ElemType("f64", np.dtype("<f8"), 8, 0, 8, KDF_FLOAT, 9),
    ElemType("rec16_u8", REC16_U8, 16, 0, 1, KDF_UNSIGNED, 10),
    ElemType("rec8_u32", REC8_U32, 8, 0, 4, KDF_UNSIGNED, 11),
    ElemType("rec16_u64", REC16_U64, 16, 0, 8, KDF_UNSIGNED, 12),
    ElemType("rec16_f64", REC16_F64, 16, 8, 8, KDF_FLOAT, -1),
    ElemType("rec16_f32", REC16_F32, 16, 4, 4, KDF_FLOAT, -1),
    ElemType("rec12_u32", REC12_U32, 12, 0, 4, KDF_UNSIGNED, -1),
    ElemType("rec24_f64", REC24_F64, 24, 12, 8, KDF_FLOAT, -1),
    ElemType("rec7_i16", REC7_I16, 7, 5, 2, KDF_SIGNED, -1),
]}


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and _ref/libradix_ref.so (when /root/reference exists)."""
    if force or not os.path.exists(LIB_ORACLE) or \
            os.path.getmtime(LIB_ORACLE) < os.path.getmtime(os.path.join(HERE, "rsx_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/radix_sort.hpp") and (
            force or not os.path.exists(LIB_REF)
            or os.path.getmtime(LIB_REF) < os.path.getmtime(os.path.join(HERE, "ref_shim.cpp"))):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The C restatement (rsx_oracle.c)."""

    def __init__(self):
        build()
        L = C.CDLL(LIB_ORACLE)
        L.orc_kdf.restype = C.c_uint64
        L.orc_kdf.argtypes = [C.c_void_p, C.POINTER(OrcLayout)]
        L.orc_radix_sort.restype = C.c_void_p
        L.orc_radix_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(OrcLayout),
                                     C.c_void_p, C.POINTER(OrcReport)]
        L.orc_radix_sort_rank.restype = C.c_void_p
        L.orc_radix_sort_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(OrcLayout),
                                          C.c_int, C.c_int, C.POINTER(OrcReport)]
        L.orc_stable_sort.restype = None
        L.orc_stable_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(OrcLayout)]
        L.orc_stable_argsort.restype = None
        L.orc_stable_argsort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.POINTER(OrcLayout)]
        self.L = L

    def kdf(self, rec: np.ndarray, layout: OrcLayout) -> int:
        rec = np.ascontiguousarray(rec)
        return int(self.L.orc_kdf(_ptr(rec), C.byref(layout)))

    def radix_sort(self, data: np.ndarray, layout: OrcLayout, want_hist: bool = False):
        """Returns (sorted copy, report, hist or None).  `data` is not modified."""
        src = np.ascontiguousarray(data).copy()
        aux = np.empty_like(src)
        n = src.shape[0]
        rep = OrcReport()
        hist = np.zeros((layout.key_bytes, 256), dtype=np.uint64) if want_hist else None
        res = self.L.orc_radix_sort(_ptr(src), _ptr(aux), n, C.byref(layout),
                                    _ptr(hist) if want_hist else None, C.byref(rep))
        if n and res == aux.ctypes.data:
            assert rep.result_in_aux == 1
            out = aux
        else:
            assert rep.result_in_aux == 0
            out = src
        return out, rep, hist

    def radix_sort_rank(self, data: np.ndarray, layout: OrcLayout, idx_dtype=np.uint32,
                        as_shipped: bool = False):
        """Returns (ranks, report, full 2n index buffer)."""
        src = np.ascontiguousarray(data)
        n = src.shape[0]
        idx_dtype = np.dtype(idx_dtype)
        ib = np.full(2 * n, 0xEE, dtype=idx_dtype)
        rep = OrcReport()
        self.L.orc_radix_sort_rank(_ptr(src), _ptr(ib), n, C.byref(layout), idx_dtype.itemsize,
                                   1 if as_shipped else 0, C.byref(rep))
        ranks = ib[n:2 * n] if rep.result_in_aux else ib[:n]
        return ranks, rep, ib

    def stable_sort(self, data: np.ndarray, layout: OrcLayout) -> np.ndarray:
        a = np.ascontiguousarray(data).copy()
        tmp = np.empty_like(a)
        self.L.orc_stable_sort(_ptr(a), _ptr(tmp), a.shape[0], C.byref(layout))
        return a

    def stable_argsort(self, data: np.ndarray, layout: OrcLayout) -> np.ndarray:
        src = np.ascontiguousarray(data)
        n = src.shape[0]
        idx = np.empty(n, dtype=np.uint64)
        tmp = np.empty(n, dtype=np.uint64)
        self.L.orc_stable_argsort(_ptr(src), _ptr(idx), _ptr(tmp), n, C.byref(layout))
        return idx


class Ref:
    """The unmodified reference headers (ref_shim.cpp).  Raises FileNotFoundError if absent."""

    def __init__(self):
        build()
        if not os.path.exists(LIB_REF):
            raise FileNotFoundError(LIB_REF)
        L = C.CDLL(LIB_REF)
        L.ref_radix_sort.restype = C.c_int
        L.ref_radix_sort.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.ref_radix_sort_rank.restype = C.c_int
        L.ref_radix_sort_rank.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                          C.c_int]
        L.ref_record_bytes.restype = C.c_size_t
        L.ref_record_bytes.argtypes = [C.c_int]
        self.L = L
        for t in TYPES.values():
            if t.ref_code < 0:
                continue  # not instantiated in ref_shim.cpp
            assert L.ref_record_bytes(t.ref_code) == t.record_bytes == t.dtype.itemsize, t.name

    @staticmethod
    def available() -> bool:
        return os.path.exists(LIB_REF) or os.path.exists("/root/reference/radix_sort.hpp")

    def radix_sort_inplace(self, t: ElemType, src: np.ndarray, aux: np.ndarray,
                           descending: bool = False) -> int:
        """Runs the reference on caller buffers; returns 1 if the result is in aux."""
        r = self.L.ref_radix_sort(t.ref_code, _ptr(src), _ptr(aux), src.shape[0],
                                  1 if descending else 0)
        assert r >= 0
        return r

    def radix_sort(self, t: ElemType, data: np.ndarray, descending: bool = False):
        src = np.ascontiguousarray(data).copy()
        # zeros, not empty: the reference assigns structs member-wise, so padding bytes of a
        # record type may be left as they were in aux; inputs built by tests have zero padding.
        aux = np.zeros_like(src)
        r = self.radix_sort_inplace(t, src, aux, descending)
        return (aux if r else src), r

    def radix_sort_rank(self, t: ElemType, data: np.ndarray, idx_dtype=np.uint32,
                        descending: bool = False):
        src = np.ascontiguousarray(data)
        n = src.shape[0]
        idx_dtype = np.dtype(idx_dtype)
        ib = np.full(2 * n, 0xEE, dtype=idx_dtype)
        r = self.L.ref_radix_sort_rank(t.ref_code, _ptr(src), _ptr(ib), n, idx_dtype.itemsize,
                                       1 if descending else 0)
        assert r >= 0
        return (ib[n:2 * n] if r else ib[:n]), r, ib
