"""On-hardware parity of the multi-GPU path (SURVEY.md section 8e "Verification"), run under torchrun:
world x n_per records through the SAME code path as bench.py's timed N > 1 steps -- fused peer
stores and the NCCL all-to-all -- gathered on rank 0 and compared byte for byte with the CPU
oracle's radix_sort of the concatenated input.

Lives under tests/ because it uses the oracle (the checker); callers: tests/test_multi_gpu.py and
bench.py's N > 1 arm (which reports the outcome as "selftest" next to its timings)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SELFTEST_CASES = [  # (element type, dist, mask): bytes compared with the CPU oracle on rank 0
    ("u32", "uniform", (1 << 64) - 1), ("u64", "uniform", (1 << 64) - 1), ("f32", "uniform", (1 << 64) - 1),
    ("rec8_u32", "uniform", 0x000FFFFF),  # heavy ties + payload = global position: stability ACROSS source ranks
    ("u32", "zipf", (1 << 64) - 1), ("u64", "zipf", (1 << 64) - 1),  # skew: key-range routing
    ("u64", "uniform", 0x0000FFFFFFFFFFFF),  # top two columns constant: routing falls back to column 5
]


def selftest(rsx, rank, world, dev, n_per=1 << 21):
    """Returns a JSON-able dict; raises on any difference.  Collective over the default group."""
    import torch
    import torch.distributed as dist
    import pyoracle
    partitioned_sort = importlib.import_module("radix-sorting_b200.dist").partitioned_sort
    keygen = importlib.import_module("radix-sorting_b200.keygen")
    results = []
    for tname, dname, mask in SELFTEST_CASES:
        t = pyoracle.TYPES[tname]
        kf = rsx.KeyFunc(t.kdf_kind, False, t.record_bytes, t.key_offset, t.key_bytes)
        kdt = torch.int32 if t.key_bytes == 4 else torch.int64
        k = torch.empty(n_per, dtype=kdt, device=dev)
        rsx.fill_keys(k, seed=77, start=rank * n_per, dist=dname, mask=mask)
        if t.dtype.names:  # {key, payload = global position}
            rec = torch.empty(n_per, 2, dtype=kdt, device=dev)
            rec[:, 0] = k
            rec[:, 1] = torch.arange(rank * n_per, (rank + 1) * n_per, dtype=kdt, device=dev)
            shard = rec.reshape(-1)
        else:
            shard = k
        want = None
        if rank == 0:
            keys = keygen.fill(77, 0, world * n_per, t.key_bytes, dname, mask)
            if t.dtype.names:
                data = np.zeros(world * n_per, dtype=t.dtype)
                data["key"] = keys
                data["payload"] = np.arange(world * n_per, dtype=data.dtype["payload"])
            else:
                data = keys.view(t.dtype)
            want, _, _ = pyoracle.Oracle().radix_sort(data, t.layout())
        # keys-only records take the append-mode exchange by default; every case also runs the exact
        # fused exchange and the NCCL all-to-all
        for fused, exact in ((True, False), (True, True), (False, False)):
            work = torch.empty(int(shard.numel() * 1.3) + 4096, dtype=shard.dtype, device=dev)
            work[: shard.numel()].copy_(shard)
            res, info = partitioned_sort(work, kf, n=n_per, fused=fused, exact=exact)
            mine = res.cpu().numpy().tobytes()
            parts = [None] * world if rank == 0 else None
            dist.gather_object(mine, parts, dst=0)
            ok = True
            if rank == 0:
                ok = b"".join(parts) == want.tobytes()
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.broadcast(flag, src=0)
            results.append({"type": tname, "dist": dname, "fused": bool(info.fused), "append": bool(info.append),
                            "requested": "fused" + ("-exact" if exact else "") if fused else "nccl",
                            "routing_column": info.routing_column, "imbalance": round(info.imbalance, 3),
                            "bit_exact_vs_oracle": bool(flag.item())})
            if not int(flag.item()):
                raise AssertionError(f"multi-GPU selftest: {tname}/{dname} fused={fused} exact={exact} differs from the oracle")
            del work, res
        del shard, k
        torch.cuda.empty_cache()
    return {"records_per_gpu": n_per, "world": world, "cases": results, "passed": True}


