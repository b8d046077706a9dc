"""GPU parity at sizes where every persistent CTA of the scatter pass works through SEVERAL tiles
(mbarrier phase flips / register prefetch, hot-digit carry-over, counter re-zeroing, look-back over
hundreds of predecessor tiles) -- the code the small goldens cannot reach -- for the
stability-sensitive kinds: records with a payload and rank sorts, with heavy ties.

Part 1 compares bytes with the CPU oracle (rsx_oracle.c, pinned to the unmodified reference by
tests/test_oracle.py) at n = 12 M.  Part 2 runs BASELINE config 4 at its full size (1 B records),
where no CPU oracle finishes in test time: the result is pinned by properties that admit exactly
one answer -- ordered by (key, original position), a permutation, every payload still attached
to its key -- which is the reference's stable order (radix_sort.hpp:83-90, radix_sort_rank.hpp:77-91).
"""
import numpy as np
import pytest

from cases import make_input
from pyoracle import TYPES
from test_gpu_parity import gpu_rank, gpu_sort, kf_for, to_dev

pytestmark = pytest.mark.gpu

N_LARGE = 12_000_007  # 4-byte keys: 11 264-record tiles x 296 CTAs -> 3.6 tiles per CTA; 8/16-byte: 6.6 / 13


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the product has no CPU path (run with -m gpu on the GPU box)")
    return torch


_cache = {}


def _case(oracle, tname, dist, mask, rank_idx=None):
    """Input + oracle answer, computed once per module (the oracle is the slow side)."""
    key = (tname, dist, mask, rank_idx)
    if key not in _cache:
        t = TYPES[tname]
        data = make_input(tname, N_LARGE, 4242, dist, mask)
        if rank_idx is None:
            want, orep, _ = oracle.radix_sort(data, t.layout())
        else:
            want, orep, _ = oracle.radix_sort_rank(data, t.layout(), rank_idx)
        while len(_cache) >= 3:  # a few cases resident: 12 M x 16 B x (input + answer) each
            _cache.pop(next(iter(_cache)))
        _cache[key] = (data, want, orep)
    return _cache[key]


@pytest.fixture(params=[(0, 0), (1, 0), (0, 1), (1, 1)], ids=["ticket", "ballot", "ticket-wide", "ballot-wide"])
def modes(request, rsx):
    rank_mode, wide = request.param
    assert rsx.lib().rsx_set_option(b"rank_mode", rank_mode) == 0
    assert rsx.lib().rsx_set_option(b"force_wide", wide) == 0
    yield request.param
    rsx.lib().rsx_set_option(b"rank_mode", -1)
    rsx.lib().rsx_set_option(b"force_wide", 0)


TIES = [("uniform", 0x000FFFFF), ("and3", (1 << 64) - 1)]


@pytest.mark.parametrize("dist,mask", TIES, ids=["mask20", "and3"])
@pytest.mark.parametrize("tname", ["rec8_u32", "rec16_u64", "rec16_u8", "u32", "u64"])
def test_value_sort_many_tiles_per_cta(rsx, torch, oracle, modes, tname, dist, mask):
    data, want, orep = _case(oracle, tname, dist, mask)
    out, rep, _ = gpu_sort(rsx, torch, tname, data)
    assert rep.ncols == orep.ncols and rep.result_in_aux == orep.result_in_aux
    assert out.tobytes() == want.tobytes(), "differs from the oracle (stability / multi-tile loop)"


@pytest.mark.parametrize("dist,mask", TIES, ids=["mask20", "and3"])
@pytest.mark.parametrize("tname,idt", [("u32", np.uint32), ("u32", np.uint64), ("u64", np.uint64), ("f32", np.uint32),
                                       ("rec8_u32", np.uint32)])
def test_rank_sort_many_tiles_per_cta(rsx, torch, oracle, modes, tname, idt, dist, mask):
    data, want, orep = _case(oracle, tname, dist, mask, idt)
    ranks, rep, _ = gpu_rank(rsx, torch, tname, data, idt)
    assert rep.ncols == orep.ncols and rep.result_in_aux == orep.result_in_aux
    assert np.array_equal(ranks, want), "ranks differ from the oracle's stable argsort"


@pytest.mark.parametrize("tname,col", [("rec8_u32", 0), ("rec8_u32", 2), ("u64", 1)])
def test_single_pass_many_tiles_is_stable(rsx, torch, modes, tname, col):
    """One K3 pass with a payload lane over ~1000 tiles == numpy's stable argsort on that digit."""
    t = TYPES[tname]
    n = N_LARGE
    data = make_input(tname, n, 99, "and2")
    src = to_dev(torch, data)
    dst = torch.zeros_like(src)
    pl_src = torch.arange(n, dtype=torch.int32, device="cuda")
    pl_dst = torch.zeros_like(pl_src)
    rsx.scatter_pass(src, dst, col, kf_for(rsx, tname), pl_src, pl_dst)
    torch.cuda.synchronize()
    keys = (data["key"] if data.dtype.names else data).view(f"<u{t.key_bytes}")
    digits = ((keys >> keys.dtype.type(8 * col)) & keys.dtype.type(0xFF)).astype(np.uint8)
    perm = np.argsort(digits, kind="stable")
    assert np.array_equal(pl_dst.cpu().numpy(), perm.astype(np.int32))
    assert dst.cpu().numpy().tobytes() == data[perm].tobytes()


# ---- BASELINE config 4 at full size -------------------------------------------------------------------------

def _chunks(n, step=1 << 27):
    for s in range(0, n, step):
        yield s, min(n, s + step)


def _check_sorted_by_key_then_position(torch, key_at, pos_at, n):
    """key non-decreasing, ties in ascending original position -- checked in chunks with one
    element of overlap; `key_at(lo, hi)` returns int64 keys (unsigned order), `pos_at` int64."""
    for lo, hi in _chunks(n):
        h = min(n, hi + 1)
        k, p = key_at(lo, h), pos_at(lo, h)
        ok = (k[1:] > k[:-1]) | ((k[1:] == k[:-1]) & (p[1:] > p[:-1]))
        assert bool(ok.all()), f"order / stability violated in [{lo}, {h})"


@pytest.mark.parametrize("mask", [(1 << 64) - 1, 0x000FFFFF], ids=["uniform", "heavy-ties"])
def test_config4a_rank_sort_1B(rsx, torch, modes, mask):
    """C4a: radix_sort_rank over 1 B u32 keys, u32 indices (radix_sort_rank.hpp:97-112)."""
    if modes[1]:
        pytest.skip("n >= 2^30 would be needed for wide offsets anyway; covered by the non-forced run")
    n = 1_000_000_000
    keys = torch.empty(n, dtype=torch.int32, device="cuda")
    rsx.fill_keys(keys, seed=6, mask=mask)
    ib = torch.empty(2 * n, dtype=torch.int32, device="cuda")
    rep = rsx.RsxReport()
    ranks = rsx.radix_sort_rank(keys, ib, n, rsx.KeyFunc(rsx.KDF_UNSIGNED), report=rep)
    assert rep.ncols == (4 if mask > 0xFFFFFFFF else 3)
    assert ranks.data_ptr() == ib.data_ptr() + (4 * n if rep.ncols & 1 else 0)  # radix_sort_rank.hpp:91

    def key_at(lo, hi):
        return torch.gather(keys, 0, ranks[lo:hi].to(torch.int64)).to(torch.int64) & 0xFFFFFFFF

    _check_sorted_by_key_then_position(torch, key_at, lambda lo, hi: ranks[lo:hi].to(torch.int64), n)
    seen = torch.zeros(n, dtype=torch.bool, device="cuda")
    for lo, hi in _chunks(n):
        seen[ranks[lo:hi].to(torch.int64)] = True
    assert bool(seen.all()), "ranks are not a permutation of 0..n-1"


@pytest.mark.parametrize("mask", [(1 << 64) - 1, 0x000FFFFF], ids=["uniform", "heavy-ties"])
def test_config4b_record_sort_1B(rsx, torch, modes, mask):
    """C4b: 1 B {u32 key, u32 payload = original position} records, stable by key."""
    if modes[1]:
        pytest.skip("covered by the non-forced run")
    n = 1_000_000_000
    recs = torch.empty(n, 2, dtype=torch.int32, device="cuda")
    k = torch.empty(n, dtype=torch.int32, device="cuda")
    rsx.fill_keys(k, seed=6, mask=mask)
    recs[:, 0] = k
    recs[:, 1] = torch.arange(n, dtype=torch.int32, device="cuda")
    src = recs.reshape(-1)
    aux = torch.empty_like(src)
    rep = rsx.RsxReport()
    res = rsx.radix_sort(src, aux, None, rsx.KeyFunc(rsx.KDF_UNSIGNED, False, 8, 0, 4), report=rep)
    assert rep.ncols == (4 if mask > 0xFFFFFFFF else 3)
    assert (res.data_ptr() == aux.data_ptr()) == bool(rep.ncols & 1)  # radix_sort.hpp:89-92
    r2 = res.view(n, 2)
    _check_sorted_by_key_then_position(torch, lambda lo, hi: r2[lo:hi, 0].to(torch.int64) & 0xFFFFFFFF,
                                       lambda lo, hi: r2[lo:hi, 1].to(torch.int64), n)
    for lo, hi in _chunks(n):  # every payload still sits next to the key it started with
        assert bool((torch.gather(k, 0, r2[lo:hi, 1].to(torch.int64)) == r2[lo:hi, 0]).all())
    seen = torch.zeros(n, dtype=torch.bool, device="cuda")
    for lo, hi in _chunks(n):
        seen[r2[lo:hi, 1].to(torch.int64)] = True
    assert bool(seen.all())
