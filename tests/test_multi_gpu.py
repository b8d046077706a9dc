"""Multi-GPU C ABI (include/rsx.h: rsx_multi_route, rsx_sort_shard, rsx_sort_multi).

CPU part: the routing arithmetic (pure host code in librsx.so) needs no device.
GPU part (needs >= 2 GPUs on the box, skipped otherwise): `rsx_sort_multi` through ctypes and the
torchrun path (`tests/multi_gpu_selftest.py`: fused peer stores and NCCL all-to-all) against the oracle.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import make_input
from pyoracle import TYPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _route(rsx, hist, rank, thr=1.15):
    world, cols, _ = hist.shape
    r = rsx.RsxRoute()
    st = rsx.lib().rsx_multi_route(np.ascontiguousarray(hist).ctypes.data_as(C.POINTER(C.c_uint64)), world, cols, rank, thr,
                                   C.byref(r))
    assert st == 0
    return r


def test_route_is_consistent_across_ranks(rsx):
    """Every rank derives the same owner table; send/recv/offset tables agree pairwise."""
    rng = np.random.default_rng(3)
    world, cols = 4, 4
    hist = np.zeros((world, cols, 256), dtype=np.uint64)
    for g in range(world):
        keys = rng.integers(0, 1 << 32, 5000 + 100 * g, dtype=np.uint64)
        for c in range(cols):
            hist[g, c] = np.bincount(((keys >> (8 * c)) & 0xFF).astype(np.int64), minlength=256)
    routes = [_route(rsx, hist, r) for r in range(world)]
    owner = np.array(list(routes[0].owner))
    assert np.all(np.diff(owner.astype(np.int64)) >= 0) and owner.max() == world - 1
    for r in routes:
        assert r.routing_column == 3 and r.live_mask == 0xF and r.key_range == 0
        assert list(r.owner) == list(routes[0].owner)
        assert r.n_total == int(hist[:, 0].sum())
    for g in range(world):
        for d in range(world):
            want = int(hist[g, 3][owner == d].sum())
            assert routes[g].send_counts[d] == want and routes[d].recv_counts[g] == want
            assert routes[g].dest_offset[d] == sum(int(hist[h, 3][owner == d].sum()) for h in range(g))
        assert routes[g].n_out == sum(routes[g].recv_counts[h] for h in range(world))
    assert max(r.n_out for r in routes) == routes[0].max_n_out
    assert routes[0].imbalance < 1.1


def test_route_skips_globally_constant_columns_and_flags_skew(rsx):
    world = 2
    hist = np.zeros((world, 8, 256), dtype=np.uint64)
    for g in range(world):
        hist[g, :, 0] = 1000  # every column constant ...
        hist[g, 0, 0] = 0
        hist[g, 0, :4] = 250  # ... except column 0 (4 buckets) and column 2
        hist[g, 2, 0] = 0
        hist[g, 2, 7] = 990
        hist[g, 2, 9] = 10
    r = _route(rsx, hist, 0)
    assert r.routing_column == 2 and r.live_mask == 0b101
    assert r.key_range == 1  # 99 % of the records sit in one bucket of the routing digit
    const = np.zeros((world, 4, 256), dtype=np.uint64)
    const[:, :, 5] = 77
    r = _route(rsx, const, 1)
    assert r.routing_column == -1 and r.n_out == 77


def test_splitters_are_quantiles(rsx):
    dsort = importlib.import_module("radix-sorting_b200.dist")
    s = np.arange(1000, dtype=np.uint64)[::-1].copy()
    assert dsort.choose_splitters(s, 4) == [250, 500, 750]


# ---- on hardware ---------------------------------------------------------------------------------------

def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _sort_multi_vs_oracle(rsx, oracle, tname, dist, mask, devices):
    """Ragged shards (one per entry of `devices`) through rsx_sort_multi; concatenated output == oracle."""
    import torch
    ng = len(devices)
    t = TYPES[tname]
    n_per = [1_000_003, 700_001, 1_200_007, 5][:ng]  # ragged shards
    data = make_input(tname, sum(n_per), 606, dist, mask & ((1 << 64) - 1))
    cap = int(max(n_per) * 1.6) + 4096
    src, aux = [], []
    off = 0
    for g in range(ng):
        raw = np.ascontiguousarray(data[off:off + n_per[g]]).view(np.uint8).reshape(-1)
        off += n_per[g]
        b = torch.zeros(cap * t.record_bytes, dtype=torch.uint8, device=f"cuda:{devices[g]}")
        b[: raw.shape[0]] = torch.from_numpy(raw.copy()).to(f"cuda:{devices[g]}")
        src.append(b)
        aux.append(torch.zeros_like(b))
    L = rsx.RsxLayout(t.record_bytes, t.key_offset, t.key_bytes, t.kdf_kind, 0)
    devs = (C.c_int * ng)(*devices)
    srcp = (C.c_void_p * ng)(*[b.data_ptr() for b in src])
    auxp = (C.c_void_p * ng)(*[b.data_ptr() for b in aux])
    ns = (C.c_size_t * ng)(*n_per)
    res = (C.c_void_p * ng)()
    nout = (C.c_size_t * ng)()
    reps = (rsx.RsxMultiReport * ng)()
    for g in set(devices):
        torch.cuda.synchronize(g)
    st = rsx.lib().rsx_sort_multi(ng, devs, srcp, auxp, ns, cap, C.byref(L), 0, res, nout, reps)
    assert st == 0, (st, rsx.lib().rsx_last_cuda_error())
    got = b""
    for g in range(ng):
        buf = src[g] if res[g] == src[g].data_ptr() else aux[g]
        assert res[g] in (src[g].data_ptr(), aux[g].data_ptr())
        got += buf[: nout[g] * t.record_bytes].cpu().numpy().tobytes()
    want, _, _ = oracle.radix_sort(data, t.layout())
    assert sum(nout) == sum(n_per)
    assert got == want.tobytes(), "concatenated shards differ from radix_sort of the concatenated input"
    if dist == "zipf":
        assert reps[0].key_range == 1
    if dist == "uniform":
        assert reps[0].fused == 1
        if mask == -1:
            assert reps[0].routing_column == t.key_bytes - 1
    return reps


_MULTI_CASES = [("u32", "uniform", -1), ("u64", "uniform", -1), ("rec8_u32", "uniform", 0xFFFFF),
                ("u32", "zipf", -1), ("f32", "uniform", -1), ("u64", "constant", -1)]


@pytest.mark.gpu
@pytest.mark.parametrize("tname,dist,mask", _MULTI_CASES)
def test_sort_multi_matches_oracle(rsx, oracle, tname, dist, mask):
    """rsx_sort_multi (one process, a thread per GPU, fused peer stores) on 2+ GPUs == oracle."""
    ng = min(_gpus(), 4)
    if ng < 2:
        pytest.skip("needs >= 2 GPUs on one box (run with gpurun --gpus 2)")
    _sort_multi_vs_oracle(rsx, oracle, tname, dist, mask, list(range(ng)))


@pytest.mark.gpu
@pytest.mark.parametrize("tname,dist,mask", _MULTI_CASES)
def test_sort_multi_with_every_shard_on_one_gpu(rsx, oracle, tname, dist, mask):
    """The same orchestration with three shards that live on ONE device (devices = {0, 0, 0}): routing,
    the fused partition + exchange (exact offsets for records, append cursors for keys) and the local
    sorts run exactly as on three GPUs, only the "peer" stores stay on the device -- so the multi-GPU
    path is covered on a single-GPU box too."""
    reps = _sort_multi_vs_oracle(rsx, oracle, tname, dist, mask, [0, 0, 0])
    if tname in ("u32", "u64", "f32") and dist != "constant":
        assert reps[0].append == 1  # keys-only, world <= record_bytes (or key ranges): append mode
    if tname == "rec8_u32":
        assert reps[0].append == 0  # payloads: order-preserving exact offsets


@pytest.mark.gpu
def test_torchrun_selftest_fused_and_nccl():
    """multi_gpu_selftest under torchrun: every case through fused peer stores AND all_to_all, bytes
    compared with the oracle (this is the leg bench.py runs before its timed N > 1 steps)."""
    ng = min(_gpus(), 8)
    if ng < 2:
        pytest.skip("needs >= 2 GPUs on one box (run with gpurun --gpus 2)")
    code = ("import os, sys, json, importlib, torch, torch.distributed as dist\n"
            f"sys.path[:0] = [{ROOT!r}, {os.path.join(ROOT, 'oracle')!r}, {os.path.join(ROOT, 'tests')!r}]\n"
            "rank = int(os.environ['RANK']); torch.cuda.set_device(int(os.environ['LOCAL_RANK']))\n"
            "dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))\n"
            "dist.init_process_group('nccl', device_id=dev)\n"
            "rsx = importlib.import_module('radix-sorting_b200'); d = importlib.import_module('multi_gpu_selftest')\n"
            "r = d.selftest(rsx, rank, dist.get_world_size(), dev, n_per=1 << 20)\n"
            "print('SELFTEST', json.dumps(r)) if rank == 0 else None\n"
            "dist.destroy_process_group()\n")
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix="_rsx_selftest.py", delete=False) as f:
        f.write(code)
        script = f.name
    try:
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ng}",
                              "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                             capture_output=True, text=True, timeout=900)
    finally:
        os.unlink(script)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "SELFTEST" in out.stdout and '"passed": true' in out.stdout
