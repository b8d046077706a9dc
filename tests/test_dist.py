"""Multi-process host logic of the partitioned sort (radix-sorting_b200/dist.py), world_size 2
and 3 over gloo on CPU.  The local primitives come from an oracle-backed engine defined HERE
(tests may use the oracle; the product's default engine is CUDA-only)."""
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


class OracleEngine:
    """CPU stand-in for CudaEngine: same four primitives, computed by the oracle / numpy."""

    def __init__(self, tname):
        import pyoracle
        self.orc = pyoracle.Oracle()
        self.t = pyoracle.TYPES[tname]

    def _np(self, x):
        return x.numpy().view(self.t.dtype)

    def histogram(self, keys, kf):
        _, rep, hist = self.orc.radix_sort(self._np(keys), self.t.layout(), want_hist=True)
        if keys.numel() == 1:
            k = self.orc.kdf(self._np(keys)[:1], self.t.layout())
            for c in range(self.t.key_bytes):
                hist[c, (k >> (8 * c)) & 0xFF] = 1
        return hist

    def scatter_pass(self, src, dst, col, kf):
        a = self._np(src)
        L = self.t.layout()
        digits = np.array([(self.orc.kdf(a[i:i + 1], L) >> (8 * col)) & 0xFF for i in range(a.shape[0])], dtype=np.int64)
        dst.copy_(src[torch.from_numpy(np.argsort(digits, kind="stable"))])
        return dst

    def sort(self, src, aux, kf):
        out, _, _ = self.orc.radix_sort(self._np(src), self.t.layout())
        src.copy_(torch.from_numpy(out.view(src.numpy().dtype)))
        return src

    def empty(self, n, like):
        return torch.empty(n, dtype=like.dtype)

    # key-range routing (skewed inputs)
    def _derived(self, keys):
        dsort = importlib.import_module("radix-sorting_b200.dist")
        a = self._np(keys)
        return dsort.derive_np(a.view(np.uint8).reshape(a.shape[0], self.t.record_bytes), self.t.layout())

    def sample_keys(self, keys, kf, count):
        d = self._derived(keys)
        idx = np.linspace(0, len(d) - 1, num=min(count, len(d))).astype(np.int64)
        return d[idx]

    def split_counts(self, keys, splitters, kf):
        dest = np.searchsorted(np.array(splitters, dtype=np.uint64), self._derived(keys), side="right")
        return [int((dest == j).sum()) for j in range(len(splitters) + 1)]

    def split_partition(self, keys, splitters, counts, kf):
        dest = np.searchsorted(np.array(splitters, dtype=np.uint64), self._derived(keys), side="right")
        return keys[torch.from_numpy(np.argsort(dest, kind="stable"))].clone()


def _worker(rank, world, port, tname, n_per, dist_name, mask, q):
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
        res, info = dsort.partitioned_sort(keys, kf, engine=OracleEngine(tname))
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
    outs = sorted(q.get(timeout=180) for _ in range(world))
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
