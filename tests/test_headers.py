"""The C++ drop-in headers (include/radix_sort.hpp, radix_sort_rank.hpp, radix_sort_basic_kdf.hpp):
host-side mirror of the reference interface over the C ABI."""
import importlib
import os
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")


@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "radix-sorting_b200", "csrc"), "-j", "8"],
                          stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", TOOLS], stdout=subprocess.DEVNULL)
    return TOOLS


def _compile(src: str, tmp_path, extra=()):
    f = tmp_path / "t.cpp"
    f.write_text(textwrap.dedent(src))
    return subprocess.run(["g++", "-std=gnu++17", "-fsyntax-only", f"-I{ROOT}/include", *extra, str(f)],
                          capture_output=True, text=True)


def test_reference_call_syntax_compiles(tmp_path):
    """Every call form the reference's own programs use with the default KDF compiles unchanged
    (radix_tests.cpp:163,193; radix_experiment.cpp:205; radix_bench.cpp:92)."""
    r = _compile("""
        #include "radix_sort.hpp"
        #include "radix_sort_rank.hpp"
        #include <cstdint>
        template <typename T> void f(T* s, T* a, size_t n) { T* r = radix_sort(s, a, n); (void)r; }
        void g() {
            f<uint8_t>(0,0,0); f<uint16_t>(0,0,0); f<uint32_t>(0,0,0); f<uint64_t>(0,0,0);
            f<int8_t>(0,0,0); f<int16_t>(0,0,0); f<int32_t>(0,0,0); f<int64_t>(0,0,0);
            f<float>(0,0,0); f<double>(0,0,0);
            const uint32_t* src = nullptr; uint32_t* ib = nullptr; uint8_t* ib8 = nullptr;
            uint32_t* r = radix_sort_rank(src, ib, 0); (void)r;
            uint8_t* r8 = radix_sort_rank(src, ib8, 0, basic_kdfs::descending{}); (void)r8;
            static_assert(basic_kdfs::highbit<int32_t>() == 0x80000000u, "");
            static_assert(basic_kdfs::highbit<int64_t>() == 0x8000000000000000ull, "");
        }
    """, tmp_path)
    assert r.returncode == 0, r.stderr


def test_arbitrary_lambda_is_a_compile_error(tmp_path):
    """A host-only callable cannot run on the device: rejected at compile time, no CPU fallback."""
    r = _compile("""
        #include "radix_sort.hpp"
        void g(int* s, int* a) { radix_sort(s, a, 4, [](const int& v) -> unsigned { return ~v; }); }
    """, tmp_path)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_host_kdf_matches_reference_identities(tmp_path, built):
    src = """
        #include "radix_sort_basic_kdf.hpp"
        #include <cstdio>
        #include <cmath>
        struct rec { uint32_t pad; int16_t key; };
        int main() {
            using namespace basic_kdfs;
            bool ok = kdf(int32_t(-1)) == 0x7FFFFFFFu && kdf(int32_t(0)) == 0x80000000u && kdf(uint16_t(7)) == 7
                   && kdf(-0.0f) == 0x7FFFFFFFu && kdf(0.0f) == 0x80000000u && kdf(-INFINITY) == 0x007FFFFFu
                   && kdf(1.0) == 0xBFF0000000000000ull && kdf(int8_t(-128)) == 0
                   && descending{}(uint32_t(1)) == 0xFFFFFFFEu && ascending{}(int64_t(-1)) == 0x7FFFFFFFFFFFFFFFull;
            rsx_layout L = by_member<&rec::key, desc>::layout<rec>();
            ok = ok && L.record_bytes == 8 && L.key_offset == 4 && L.key_bytes == 2 && L.kdf_kind == RSX_KDF_SIGNED && L.flags == 1;
            ok = ok && by_member<&rec::key>{}(rec{0, -2}) == 0x7FFE;
            printf(ok ? "ok\\n" : "bad\\n");
            return ok ? 0 : 1;
        }
    """
    f = tmp_path / "k.cpp"
    f.write_text(textwrap.dedent(src))
    exe = tmp_path / "k"
    subprocess.check_call(["g++", "-std=gnu++17", f"-I{ROOT}/include", "-o", str(exe), str(f)])
    assert subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip() == "ok"


def test_cpp_programs_fail_loudly_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    r = subprocess.run([os.path.join(built, "radix_tests_b200")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_reference_scenarios_on_gpu(built):
    r = subprocess.run([os.path.join(built, "radix_tests_b200")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "All tests OK." in r.stdout, r.stdout + r.stderr


def _fnv1a(buf):
    """64-bit FNV-1a of a uint8 array, through a 10-line C helper compiled once (a Python loop over
    up to 160 MB would take minutes)."""
    import ctypes as C
    import tempfile
    global _FNV
    if "_FNV" not in globals():
        d = tempfile.mkdtemp()
        src = os.path.join(d, "fnv.c")
        open(src, "w").write("#include <stdint.h>\n#include <stddef.h>\nuint64_t fnv(const unsigned char*b,size_t n){uint64_t h="
                             "0xCBF29CE484222325ULL;for(size_t i=0;i<n;++i)h=(h^b[i])*0x100000001B3ULL;return h;}\n")
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", os.path.join(d, "fnv.so"), src])
        _FNV = C.CDLL(os.path.join(d, "fnv.so"))
        _FNV.fnv.restype = C.c_uint64
        _FNV.fnv.argtypes = [C.c_void_p, C.c_size_t]
    return int(_FNV.fnv(buf.ctypes.data, buf.shape[0]))


@pytest.mark.gpu
@pytest.mark.parametrize("args", [["1000000"], ["0", "0", "0", "uint32_t", "00FFFFFF"], ["3000000", "0", "0", "uint64_t"],
                                  ["2000000", "0", "0", "float"], ["2000000", "0", "0", "int64_t"], ["65536", "0", "0", "uint8_t"],
                                  ["70000", "0", "0", "uint16_t"], ["1000000", "0", "0", "double"], ["1000000", "0", "0", "int32_t"]])
def test_radix_cli_on_gpu(built, tmp_path, args):
    """N1: the `radix` CLI on the device path.  Its output digest must equal the ORACLE's radix_sort
    of the same seeded key bytes (radix_experiment.cpp:203-223 only checks "is it ordered")."""
    import re
    import numpy as np
    import pyoracle
    keygen = importlib.import_module("radix-sorting_b200.keygen")
    r = subprocess.run([os.path.join(built, "radix_b200"), *args], capture_output=True, text=True, timeout=300,
                       cwd=str(tmp_path))
    assert r.returncode == 0 and "device-resident" in r.stdout and "CPU yardstick" in r.stdout, r.stdout + r.stderr
    m = re.search(r"digest fnv1a64=([0-9a-f]{16})", r.stdout)
    assert m, r.stdout
    tname = {"uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "int32_t": "i32", "int64_t": "i64",
             "float": "f32", "double": "f64"}[args[3] if len(args) > 3 else "uint32_t"]
    t = pyoracle.TYPES[tname]
    raw = keygen.fill(1, 0, 160_000_000 // 8, 8).view(np.uint8)  # the CLI's seeded stand-in for 40M_32bit_keys.dat
    n = raw.shape[0] // t.record_bytes
    if int(args[0]):
        n = min(n, int(args[0]))
    data = raw[: n * t.record_bytes].view(t.dtype).copy()
    if len(args) > 4:
        u = data.view(f"<u{t.record_bytes}")
        u &= np.array(int(args[4], 16) & ((1 << (8 * t.record_bytes)) - 1), dtype=u.dtype)
    want, _, _ = pyoracle.Oracle().radix_sort(data, t.layout())
    code = _fnv1a(np.ascontiguousarray(want).view(np.uint8))
    assert f"{code:016x}" == m.group(1), "CLI output differs from the oracle's radix_sort"
