"""Multi-process run of the partitioned sort's host orchestration -- the C++ `rsx_sort_shard` of
csrc/rsx_multi.cu, driven through radix-sorting_b200/dist.py -- at world_size 2 and 3 over gloo on
CPU.  The local primitives are oracle-backed callbacks defined HERE (`rsx_shard_ops`; tests may use
the oracle, the product's primitives are the CUDA kernels) and the exchange is gloo's
all_to_all."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_oracle_ops(tname):
    """rsx_shard_ops backed by the CPU oracle / numpy: lets the C++ orchestration of the partitioned
    sort (csrc/rsx_multi.cu, rsx_sort_shard) run on host memory over gloo.  Test infrastructure:
    the product always passes ops = NULL (the CUDA kernels)."""
    import ctypes as C
    import pyoracle
    rsx = importlib.import_module("radix-sorting_b200")
    dsort = importlib.import_module("radix-sorting_b200.dist")
    orc = pyoracle.Oracle()
    t = pyoracle.TYPES[tname]
    L = t.layout()

    def view(ptr, n):
        if n == 0:
            return np.zeros(0, dtype=t.dtype)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * t.record_bytes,))
        return raw.view(t.dtype)

    def derived(a):
        return dsort.derive_np(np.ascontiguousarray(a).view(np.uint8).reshape(a.shape[0], t.record_bytes), L)

    def hist(ctx, src, n, lay, out, stream):
        h = np.zeros((t.key_bytes, 256), dtype=np.uint64)
        d = derived(view(src, n))
        for c in range(t.key_bytes):
            h[c] = np.bincount(((d >> np.uint64(8 * c)) & np.uint64(0xFF)).astype(np.int64), minlength=256)
        np.ctypeslib.as_array(out, shape=(t.key_bytes * 256,))[:] = h.reshape(-1)
        return 0

    def hist_column(ctx, src, n, lay, col, out, stream):
        d = derived(view(src, n))
        np.ctypeslib.as_array(out, shape=(256,))[:] = np.bincount(
            ((d >> np.uint64(8 * col)) & np.uint64(0xFF)).astype(np.int64), minlength=256)
        return 0

    def sample(ctx, src, n, lay, count, out, stream):
        d = derived(view(src, n))
        stride = n // count
        np.ctypeslib.as_array(out, shape=(count,))[:] = d[np.arange(count) * stride]
        return 0

    def dest_of(a, col, owner, split, nsplit):
        d = derived(a)
        if col >= 0:
            own = np.ctypeslib.as_array(owner, shape=(256,)).astype(np.int64)
            return own[((d >> np.uint64(8 * col)) & np.uint64(0xFF)).astype(np.int64)]
        sp = np.ctypeslib.as_array(split, shape=(nsplit,)).copy()
        return np.searchsorted(sp, d, side="right")

    def split_counts(ctx, src, n, lay, split, nsplit, counts, stream):
        dest = dest_of(view(src, n), -1, None, split, nsplit)
        np.ctypeslib.as_array(counts, shape=(nsplit + 1,))[:] = np.bincount(dest, minlength=nsplit + 1)
        return 0

    def partition_to(ctx, src, n, lay, col, owner, split, nsplit, dest_base, ndest, stream):
        a = view(src, n).copy()
        dest = dest_of(a, col, owner, split, nsplit)
        for d in range(ndest):
            part = np.ascontiguousarray(a[dest == d])  # stable: input order inside a destination
            if part.shape[0]:
                C.memmove(int(dest_base[d]), part.ctypes.data, part.nbytes)
        return 0

    def sort(ctx, src, aux, n, lay, result, stream):
        a = view(src, n)
        out, _, _ = orc.radix_sort(a, L)
        a[:] = out
        result[0] = src
        return 0

    ops = rsx.RsxShardOps(rsx.OPS_HIST_FN(hist), rsx.OPS_SAMPLE_FN(sample), rsx.OPS_SPLIT_COUNTS_FN(split_counts),
                          rsx.OPS_PARTITION_FN(partition_to), rsx.OPS_SORT_FN(sort), None,
                          rsx.OPS_HIST_COLUMN_FN(hist_column))
    return ops


def _worker(rank, world, port, tname, n_per, dist_name, mask, q):
    import faulthandler
    faulthandler.dump_traceback_later(45, exit=True)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pyoracle
        from cases import make_input
        rsx = importlib.import_module("radix-sorting_b200")
        dsort = importlib.import_module("radix-sorting_b200.dist")
        t = pyoracle.TYPES[tname]
        # rank r holds global positions [r*n_per, (r+1)*n_per) of one seeded stream
        allkeys = make_input(tname, n_per[-1], 4321, dist_name, mask)
        lo = n_per[rank]
        hi = n_per[rank + 1]
        tdt = {4: torch.int32, 8: torch.int64}[t.key_bytes]
        keys = torch.from_numpy(allkeys[lo:hi].view(np.int32 if t.key_bytes == 4 else np.int64).copy())
        kf = rsx.KeyFunc(t.kdf_kind, False, t.record_bytes, t.key_offset, t.key_bytes)
        ops = make_oracle_ops(tname)
        res, info = dsort.partitioned_sort(keys, kf, ops=ops, fused=False)
        q.put((rank, res.numpy().tobytes(), info.n_out, info.routing_column, info.imbalance))
    finally:
        dist.destroy_process_group()


def _run(world, tname, bounds, dist_name="uniform", mask=(1 << 64) - 1):
    import pyoracle
    from cases import make_input
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + world * 13 + len(dist_name)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, tname, bounds, dist_name, mask, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=60) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = pyoracle.TYPES[tname]
    allkeys = make_input(tname, bounds[-1], 4321, dist_name, mask)
    want, _, _ = pyoracle.Oracle().radix_sort(allkeys, t.layout())
    got = b"".join(o[1] for o in outs)
    assert got == want.tobytes(), "concatenated shards differ from radix_sort of the concatenated input"
    assert sum(o[2] for o in outs) == bounds[-1]
    return outs


@pytest.mark.parametrize("tname", ["u32", "u64", "i32", "f32"])
def test_partitioned_sort_world2(tname):
    outs = _run(2, tname, [0, 3000, 6000])
    assert max(o[4] for o in outs) < 1.2  # uniform keys balance at bucket granularity


def test_partitioned_sort_world3_ragged_and_skipped_columns():
    # ragged shards; constant high bytes: the routing digit must fall back to the highest LIVE column
    outs = _run(3, "u64", [0, 1000, 1001, 5000], mask=0x0000000000FFFFFF)
    assert outs[0][3] == 2
    _run(3, "u32", [0, 10, 2000, 2500], dist_name="zipf")


def test_partitioned_sort_skewed_keys_use_key_range_routing():
    """zipf keys: most of the mass sits in one bucket of the routing digit, so the sort switches to
    sample-based key-range splitters; the result is still bit-identical and roughly balanced."""
    outs = _run(3, "u32", [0, 4000, 8000, 12000], dist_name="zipf")
    assert all(o[3] == -1 for o in outs)
    assert max(o[4] for o in outs) < 1.35
    outs = _run(2, "u64", [0, 3000, 9000], dist_name="zipf")
    assert all(o[3] == -1 for o in outs)


def test_partitioned_sort_constant_and_presorted():
    outs = _run(2, "u32", [0, 500, 1000], dist_name="constant")
    assert outs[0][3] is None  # no live column at all: nothing to route
    _run(2, "u32", [0, 500, 1000], dist_name="sorted")


def test_assign_buckets_is_contiguous_and_balanced():
    dsort = importlib.import_module("radix-sorting_b200.dist")
    rng = np.random.default_rng(0)
    for world in (2, 4, 8):
        counts = rng.integers(0, 1000, 256).astype(np.uint64)
        owner = dsort.assign_buckets(counts, world)
        assert owner[0] == 0 and np.all(np.diff(owner) >= 0) and owner.max() <= world - 1
        loads = np.array([counts[owner == r].sum() for r in range(world)], dtype=np.float64)
        assert loads.max() <= counts.sum() / world + counts.max()
    skew = np.zeros(256, dtype=np.uint64)
    skew[7] = 10**6
    owner = dsort.assign_buckets(skew, 8)
    assert np.all(np.diff(owner) >= 0)
