"""CPU-side checks of the drop-in boundary: librsx.so loads, exports every symbol that
include/rsx.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rsx.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsx_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(rsx):
    syms = declared_symbols()
    assert set(syms) == set(rsx.EXPORTS), "include/rsx.h and the Python binding list disagree"
    L = rsx.lib()
    for s in syms:
        assert hasattr(L, s), f"librsx.so does not export {s}"
    out = subprocess.run(["nm", "-D", "--defined-only", rsx.LIB_PATH], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", out), f"{s} is not a defined text symbol"


def test_every_entry_point_is_documented():
    """INTEGRATION.md names every entry point of include/rsx.h (with the reference interface it replaces)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in declared_symbols() if s not in doc]
    assert not missing, f"not in INTEGRATION.md: {missing}"


def test_version_and_strerror(rsx):
    L = rsx.lib()
    assert L.rsx_version() == 100
    assert L.rsx_strerror(0) == b"ok"
    assert b"no CPU path" in L.rsx_strerror(rsx.RSX_ERR_NO_DEVICE)


def test_layout_validation_needs_no_device(rsx):
    L = rsx.lib()
    res = C.c_void_p()
    buf = (C.c_uint8 * 64)()
    for bad in [(3, 0, 1, 0, 0), (4, 0, 3, 0, 0), (4, 2, 4, 0, 0), (16, 6, 4, 0, 0), (4, 0, 4, 3, 0),
                (4, 0, 4, 0, 2), (4, 0, 2, 2, 0)]:
        lay = rsx.RsxLayout(*bad)
        assert L.rsx_sort(buf, buf, 8, C.byref(lay), C.byref(res), None, None) == rsx.RSX_ERR_INVALID, bad
    lay = rsx.RsxLayout(4, 0, 4, 0, 0)
    # n < 2 never touches the device (radix_sort.hpp:100-101): returns src
    assert L.rsx_sort(buf, buf, 1, C.byref(lay), C.byref(res), None, None) == 0
    assert res.value == C.addressof(buf)
    assert L.rsx_workspace_bytes(1 << 20, C.byref(lay), 0) > 0


def test_no_cpu_fallback(rsx):
    """Without a CUDA device a real sort must FAIL, never silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present; the failure path is exercised in the CPU container")
    L = rsx.lib()
    res = C.c_void_p()
    src = (C.c_uint32 * 8)(5, 4, 3, 2, 1, 0, 9, 8)
    aux = (C.c_uint32 * 8)()
    lay = rsx.RsxLayout(4, 0, 4, 0, 0)
    st = L.rsx_sort(src, aux, 8, C.byref(lay), C.byref(res), None, None)
    assert st in (rsx.RSX_ERR_NO_DEVICE, rsx.RSX_ERR_CUDA)
    assert list(src) == [5, 4, 3, 2, 1, 0, 9, 8] and list(aux) == [0] * 8


def test_product_does_not_touch_the_oracle():
    """Nothing under radix-sorting_b200/ or include/ may reference oracle/."""
    bad = []
    for base in ("radix-sorting_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"pyoracle|liboracle|rsx_oracle|libradix_ref|orc_radix", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_workspace_sizing_and_options_need_no_device(rsx):
    """rsx_workspace_bytes is pure host arithmetic: a fixed head plus, per key column,
    one 256-word look-back row per tile of the SMALLEST tile any kernel of that record size runs with."""
    L = rsx.lib()
    L.rsx_workspace_bytes.restype = C.c_size_t
    L.rsx_workspace_bytes.argtypes = [C.c_size_t, C.c_void_p, C.c_int]
    u32, u64 = rsx.RsxLayout(4, 0, 4, 0, 0), rsx.RsxLayout(8, 0, 8, 0, 0)
    n = 1_000_000_000
    head = L.rsx_workspace_bytes(2, C.byref(u32), 0)
    w32 = L.rsx_workspace_bytes(n, C.byref(u32), 0)
    assert 0 < w32 - 4 * -(-n // 10240) * 256 * 4 <= head + 4096   # 4 columns x 10 240-record tiles, 4-byte status words
    w64 = L.rsx_workspace_bytes(n, C.byref(u64), 0)
    assert 0 < w64 - 8 * -(-n // 9216) * 256 * 4 <= head + 4096    # 8 columns x 9 216-record tiles
    wide = L.rsx_workspace_bytes(1 << 30, C.byref(u32), 0)         # n >= 2^30: 8-byte status words
    assert wide > 2 * w32 - head
    # rank sort: two record buffers beside the indices, nothing extra for 4/8-byte index types
    wr = L.rsx_workspace_bytes(n, C.byref(u32), 4)
    assert wr >= w32 + 2 * n * 4 and wr - 2 * n * 4 < 2 * w32  # (key + index tiles are smaller: more look-back rows)
    assert L.rsx_workspace_bytes(n, C.byref(u32), 2) >= wr + 2 * n * 4  # narrow index types sort through u32 lanes
    bad = rsx.RsxLayout(12, 0, 3, 0, 0)  # 12-byte records are fine (keys + gather path), a 3-byte key is not
    assert L.rsx_workspace_bytes(n, C.byref(bad), 0) == 0
    # options are validated on the host
    assert L.rsx_set_option(b"scatter_variant", 99) == rsx.RSX_ERR_INVALID
    assert L.rsx_set_option(b"rank_mode", 7) == rsx.RSX_ERR_INVALID
    assert L.rsx_set_option(b"no_such_option", 1) == rsx.RSX_ERR_INVALID
    assert L.rsx_set_option(b"scatter_variant", 0) == 0 and L.rsx_set_option(b"rank_mode", -1) == 0


def test_key_compaction_plan_is_order_preserving(rsx):
    """rsx_plan_compaction (host arithmetic of N4): the runs cover every varying bit, gather them in
    order (so comparing compacted keys == comparing original keys), scatter back exactly, and the
    plan is refused unless it saves enough passes.  Bit-level model in numpy, no device."""
    import numpy as np
    rng = np.random.default_rng(11)
    L = rsx.lib()
    for kb in (4, 8):
        full = (1 << (8 * kb)) - 1
        for trial in range(300):
            nbits = int(rng.integers(1, 8 * kb // 2))
            varying = 0
            for b in rng.choice(8 * kb, size=nbits, replace=False):
                varying |= 1 << int(b)
            const_ones = int(rng.integers(0, 1 << 62)) & full & ~varying
            key_or, key_nand = varying | const_ones, (varying | (~const_ones & full)) & full
            live = sum(1 for c in range(kb) if (varying >> (8 * c)) & 0xFF)
            runs = (C.c_uint32 * 24)()
            cb = C.c_uint64(0)
            passes = L.rsx_plan_compaction(key_or, key_nand, kb, live, runs, C.byref(cb))
            assert passes >= 0
            if passes == 0:
                continue
            need = 3 if kb == 4 else 2
            assert passes + need <= live, (hex(varying), passes, live)
            rr = [(runs[3 * i], runs[3 * i + 1], runs[3 * i + 2]) for i in range(8) if runs[3 * i + 1]]
            covered = 0
            at = 0
            for s_, w_, d_ in rr:
                assert d_ == at and w_ >= 1 and not (covered >> s_), "runs ascending, packed without gaps"
                covered |= ((1 << w_) - 1) << s_
                at += w_
            assert covered & varying == varying and passes == (at + 7) // 8
            assert cb.value == (const_ones & ~covered)

            def compact(k):
                return sum(((k >> s_) & ((1 << w_) - 1)) << d_ for s_, w_, d_ in rr)

            def expand(v):
                return cb.value | sum(((v >> d_) & ((1 << w_) - 1)) << s_ for s_, w_, d_ in rr)

            keys = [const_ones | (int(rng.integers(0, 1 << 62)) & varying) for _ in range(40)]
            for a in keys:
                assert expand(compact(a)) == a
            srt = sorted(keys)
            assert sorted(keys, key=compact) == srt or [compact(k) for k in sorted(keys, key=compact)] == [compact(k) for k in srt]
    assert L.rsx_plan_compaction(0xFF, 0xFF, 3, 1, None, None) == rsx.RSX_ERR_INVALID
