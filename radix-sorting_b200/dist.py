"""Single-box multi-GPU partitioned radix sort (BASELINE config 5).

One process per GPU (torchrun), `torch.distributed` for the plumbing.  The reference has no
multi-device path; this is the MSD-then-LSD composition of its own primitives:

  1. every rank runs the fused histogram kernel (K1) on its shard -> per-column digit counts
  2. all_gather of the (columns x 256) counts (a few KiB): global histogram, live columns,
     and -- because every rank sees every rank's counts -- exact send/receive sizes
  3. the highest globally-live column is the routing digit; its 256 buckets are assigned to
     ranks as contiguous ranges balancing the global counts
  4. one stable scatter pass (K3) on that column groups each rank's shard by bucket, hence by
     destination rank
  5. exchange.  Fused form (default): step 4's kernel stores every bucket straight into its
     owner's receive buffer -- peer memory mapped with torch symmetric memory, written over
     NVLink -- so the all-to-all costs no extra HBM pass (rsx_scatter_pass_to).  Baseline form
     (fused=False, or when symmetric memory is unavailable): all_to_all_single over NCCL with
     the exact split sizes.  Either way a receive buffer holds the chunks in source-rank
     order, which keeps the global order stable
  6. local LSD radix sort (K1-K3, device-side column skipping) of what was received

The concatenation of the ranks' outputs in rank order is the globally sorted sequence, and it
is bit-identical to `radix_sort` of the concatenated input (tests/test_dist.py checks this
with a two-process gloo group on CPU, using an oracle-backed engine supplied by the test).

The local work goes through an *engine* object so that the host logic above can be tested
without a GPU; the default engine is the CUDA library and raises if it is unavailable (there is
no CPU fallback in the product).
"""
from __future__ import annotations

import importlib
import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


class CudaEngine:
    """Local primitives on the GPU through librsx.so."""

    def __init__(self):
        self.rsx = importlib.import_module("radix-sorting_b200")
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("CudaEngine needs a CUDA device; there is no CPU fallback")
        self.torch = torch

    def histogram(self, keys, kf) -> np.ndarray:
        L = kf.layout(keys.element_size())
        n = keys.numel() * keys.element_size() // L.record_bytes
        if n < 2:  # the kernel needs n >= 2 (like the reference, which returns before counting)
            h = np.zeros((L.key_bytes, 256), dtype=np.uint64)
            if n == 1:
                k = derive_key_py(bytes(keys.view(self.torch.uint8).cpu().numpy().tobytes()), L)
                for c in range(L.key_bytes):
                    h[c, (k >> (8 * c)) & 0xFF] = 1
            return h
        hist, _, _ = self.rsx.histogram(keys, kf)
        return hist

    def scatter_pass(self, src, dst, col, kf):
        if src.numel():
            self.rsx.scatter_pass(src, dst, col, kf)
        return dst

    def sort(self, src, aux, kf):
        return self.rsx.radix_sort(src, aux, None, kf)

    # ---- fused partition + exchange over peer memory (NVLink) ----------------------------------
    _symm = {}  # (dtype, device) -> (tensor, handle): symmetric receive buffer, grown on demand
    symm_error = None

    def symmetric_recv(self, capacity, like, group):
        """A receive buffer of >= capacity elements allocated symmetrically on every rank, with the
        peers' addresses.  Returns (tensor, [base address per rank], handle) or None if symmetric
        memory is not available here (the caller then uses the NCCL all-to-all)."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
        except Exception:
            return None
        import torch.distributed as dist
        key = (like.dtype, like.device.index)
        if key in self._symm and self._symm[key] is None:
            return None  # failed before: do not retry on every call
        cur = self._symm.get(key)
        if cur is None or cur[0].numel() < capacity:
            try:
                t = symm_mem.empty(int(capacity * 1.02) + 1024, dtype=like.dtype, device=like.device)
                h = symm_mem.rendezvous(t, group=group if group is not None else dist.group.WORLD)
                cur = (t, h)
            except Exception as e:  # no fabric / P2P support
                self._symm[key] = None
                CudaEngine.symm_error = repr(e)
                return None
            self._symm[key] = cur
        t, h = cur
        return t, [int(p) for p in h.buffer_ptrs], h

    def scatter_pass_to(self, src, col, owner, dest_base, kf):
        self.rsx.scatter_pass_to(src, col, owner, dest_base, kf)

    # ---- key-range routing (skewed inputs) -------------------------------------------------------
    def sample_keys(self, keys, kf, count) -> np.ndarray:
        """`count` evenly spaced records' DERIVED keys (uint64)."""
        L = kf.layout(keys.element_size())
        n = keys.numel() * keys.element_size() // L.record_bytes
        raw = keys.view(self.torch.uint8).view(n, L.record_bytes)
        c = min(count, n)
        idx = (self.torch.arange(c, device=keys.device, dtype=self.torch.int64) * n) // c  # exact integer stride
        return derive_np(raw[idx].cpu().numpy(), L)

    def split_counts(self, keys, splitters, kf):
        return self.rsx.split_counts(keys, splitters, kf)

    def split_pass_to(self, keys, splitters, dest_base, kf):
        self.rsx.split_pass_to(keys, splitters, dest_base, kf)

    def split_partition(self, keys, splitters, counts, kf):
        """Local stable partition by key range (the non-fused form): returns the grouped copy."""
        L = kf.layout(keys.element_size())
        part = self.empty(keys.numel(), keys)
        base, acc = [], 0
        for c in counts:
            base.append(part.data_ptr() + acc * L.record_bytes)
            acc += c
        self.rsx.split_pass_to(keys, splitters, base, kf)
        return part

    def empty(self, n, like):
        return self.torch.empty(n, dtype=like.dtype, device=like.device)

    _scratch = {}

    def scratch(self, n, like):
        """Grow-only cached buffer (the local sort's aux): no allocator traffic in steady state.
        Like the receive buffer it is owned by the engine and reused by the next call."""
        key = (like.dtype, like.device.index)
        cur = self._scratch.get(key)
        if cur is None or cur.numel() < n:
            cur = None
            self._scratch[key] = None
            cur = self.torch.empty(int(n * 1.05) + 1024, dtype=like.dtype, device=like.device)
            self._scratch[key] = cur
        return cur[:n]


def derive_key_py(record: bytes, L) -> int:
    """Derived key of ONE record (radix_sort_basic_kdf.hpp:19-46) -- used for 1-element shards only."""
    k = int.from_bytes(record[L.key_offset:L.key_offset + L.key_bytes], "little")
    m, top = (1 << (8 * L.key_bytes)) - 1, 1 << (8 * L.key_bytes - 1)
    if L.kdf_kind == 1:
        k ^= top
    elif L.kdf_kind == 2:
        k ^= m if k & top else top
    return (~k & m) if (L.flags & 1) else k


def derive_np(records: np.ndarray, L) -> np.ndarray:
    """Vectorised derived keys (uint64) of an (n, record_bytes) uint8 array."""
    kb = L.key_bytes
    k = np.zeros(records.shape[0], dtype=np.uint64)
    for b in range(kb):
        k |= records[:, L.key_offset + b].astype(np.uint64) << np.uint64(8 * b)
    m = np.uint64((1 << (8 * kb)) - 1) if kb < 8 else np.uint64(0xFFFFFFFFFFFFFFFF)
    top = np.uint64(1 << (8 * kb - 1))
    if L.kdf_kind == 1:
        k ^= top
    elif L.kdf_kind == 2:
        k = np.where(k & top, k ^ m, k ^ top)
    if L.flags & 1:
        k = ~k & m
    return k


def choose_splitters(samples: np.ndarray, world: int):
    """world - 1 ascending derived-key splitters at the quantiles of the pooled samples."""
    srt = np.sort(samples.astype(np.uint64))
    return [int(srt[min(len(srt) - 1, (i * len(srt)) // world)]) for i in range(1, world)]


def assign_buckets(global_counts: np.ndarray, world: int) -> np.ndarray:
    """Contiguous bucket ranges per rank, balancing counts: owner[b] = rank that receives
    bucket b.  Greedy sweep against the ideal cumulative share; deterministic on every rank."""
    total = int(global_counts.sum())
    owner = np.zeros(256, dtype=np.int64)
    if total == 0 or world == 1:
        return owner
    cum = np.cumsum(global_counts.astype(np.float64))
    start = cum - global_counts  # exclusive prefix
    mid = start + global_counts / 2.0  # a bucket goes to the rank whose share contains its midpoint
    owner = np.minimum((mid * world / total).astype(np.int64), world - 1)
    return np.maximum.accumulate(owner)  # monotone (contiguous ranges)


@dataclass
class PartitionInfo:
    routing_column: Optional[int]
    live_columns: List[int]
    send_counts: List[int]
    recv_counts: List[int]
    n_out: int
    n_total: int
    imbalance: float
    seconds: dict


def _sort_by_key_ranges(keys, kf, L, group, engine, world, rank, dev, live, n_total, n_local, rec_elems, fused,
                        sec, tick, t):
    import torch
    import torch.distributed as dist
    samples = engine.sample_keys(keys, kf, 8192) if n_local else np.zeros(0, dtype=np.uint64)
    pad = np.full(8192, np.iinfo(np.uint64).max, dtype=np.uint64)  # ragged shards: fixed-size exchange
    pad[: len(samples)] = samples
    mine = torch.from_numpy(np.concatenate([[len(samples)], pad.view(np.int64)]).astype(np.int64)).to(dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    pooled = np.concatenate([g.cpu().numpy()[1:1 + int(g[0])].view(np.uint64) for g in gathered])
    splitters = choose_splitters(pooled, world)
    counts = engine.split_counts(keys, splitters, kf) if n_local else [0] * world
    cmine = torch.tensor(counts, dtype=torch.int64, device=dev)
    call = [torch.empty_like(cmine) for _ in range(world)]
    dist.all_gather(call, cmine, group=group)
    cnt = torch.stack(call).cpu().numpy()  # [source][destination]
    send = [int(c) for c in cnt[rank]]
    recv = [int(c) for c in cnt[:, rank]]
    n_out = sum(recv)
    t = tick("sample+split_counts", t)
    symm = None
    if fused and hasattr(engine, "symmetric_recv"):
        cap = int(cnt.sum(axis=0).max()) * rec_elems
        symm = engine.symmetric_recv(max(cap, 1), keys, group)
        ok = torch.tensor([1 if symm is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not int(ok.item()):
            symm = None
    if symm is not None:
        out_buf, bases, handle = symm
        dest_base = [bases[d] + int(cnt[:rank, d].sum()) * L.record_bytes for d in range(world)]
        handle.barrier()
        if n_local:
            engine.split_pass_to(keys, splitters, dest_base, kf)
        torch.cuda.synchronize(dev)
        handle.barrier()
        t = tick("fused_partition_exchange", t)
    else:
        part = engine.split_partition(keys, splitters, counts, kf) if n_local else keys
        t = tick("partition_pass", t)
        out_buf = engine.empty(max(n_out, 1) * rec_elems, keys)
        dist.all_to_all_single(out_buf[: n_out * rec_elems], part, [r * rec_elems for r in recv],
                               [s * rec_elems for s in send], group=group)
        t = tick("all_to_all", t)
        del part
    recv_view = out_buf[: n_out * rec_elems]
    if n_out > 1:
        aux = keys if keys.numel() >= recv_view.numel() else (
            engine.scratch(recv_view.numel(), keys) if hasattr(engine, "scratch") else engine.empty(recv_view.numel(), keys))
        res = engine.sort(recv_view, aux[: recv_view.numel()], kf)
    else:
        res = recv_view
    t = tick("local_sort", t)
    sec["exchange"] = ("fused peer stores (NVLink)" if symm is not None else "all_to_all_single") + ", key-range routing"
    info = PartitionInfo(-1, live, send, recv, n_out, n_total, n_out / max(n_total / world, 1), sec)
    return res, info


def partitioned_sort(keys, kf, group=None, engine=None, timers: bool = False, fused: bool = True,
                     skew_threshold: float = 1.15):
    """Globally sorts the concatenation (in rank order) of every rank's `keys`.
    Returns (this rank's slice of the sorted sequence, PartitionInfo).  `keys` is clobbered."""
    import torch
    import torch.distributed as dist
    engine = engine or CudaEngine()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = keys.device
    sec = {}

    def tick(name, t0):
        if timers:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            sec[name] = time.perf_counter() - t0
        return time.perf_counter()

    t = time.perf_counter()
    L = kf.layout(keys.element_size())
    cols = L.key_bytes
    n_local = keys.numel() * keys.element_size() // L.record_bytes

    # 1-2. histograms of every rank, visible to every rank
    hist = engine.histogram(keys, kf)  # (cols, 256) uint64, digits of the DERIVED key
    mine = torch.from_numpy(hist.astype(np.int64).reshape(-1)).to(dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    per_rank = torch.stack(gathered).cpu().numpy().reshape(world, cols, 256)  # [source rank][column][bucket]
    total = per_rank.sum(axis=0)
    n_total = int(total[0].sum())
    t = tick("histogram+allgather", t)

    # 3. routing digit = highest column that is not constant over ALL ranks (device-side column
    #    skipping, lifted to the global level)
    live = [c for c in range(cols) if int(total[c].max()) != n_total]
    if not live or world == 1:
        out = engine.sort(keys, engine.empty(keys.numel(), keys), kf) if n_local > 1 else keys
        info = PartitionInfo(None, live, [n_local], [n_local], n_local, n_total, 1.0, sec)
        return out, info
    top = live[-1]
    owner = assign_buckets(total[top], world)
    rec_elems = L.record_bytes // keys.element_size()
    predicted = max(int(total[top][owner == d].sum()) for d in range(world)) / max(n_total / world, 1)
    if predicted > skew_threshold and hasattr(engine, "split_counts") and world - 1 <= 15:
        # Skewed routing digit (e.g. zipf: most of the mass in one top bucket): bucket-granular
        # ranges cannot balance, so route by key range instead -- splitters at the quantiles of a
        # pooled sample (sample sort); exact per-destination counts from one counting pass.
        return _sort_by_key_ranges(keys, kf, L, group, engine, world, rank, dev, live, n_total, n_local,
                                   rec_elems, fused, sec, tick, t)
    send = [int(per_rank[rank, top, owner == d].sum()) for d in range(world)]
    recv = [int(per_rank[s, top, owner == rank].sum()) for s in range(world)]
    n_out = sum(recv)

    symm = None
    if fused and hasattr(engine, "symmetric_recv"):
        # every rank must take the same branch: capacity is the global maximum, known to all
        cap = max(int(per_rank[:, top, owner == d].sum()) for d in range(world)) * rec_elems
        symm = engine.symmetric_recv(max(cap, 1), keys, group)
        ok = torch.tensor([1 if symm is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not int(ok.item()):
            symm = None
    if symm is not None:
        # 4+5 fused: the stable pass on the routing column stores every destination's records
        # straight into that rank's receive buffer (peer memory over NVLink), one contiguous run
        # per (tile, destination).  Layout of a receive buffer: chunks in source-rank order.
        out_buf, bases, handle = symm
        dest_base = [bases[d] + int(per_rank[:rank, top, owner == d].sum()) * L.record_bytes for d in range(world)]
        handle.barrier()  # nobody is still sorting out of its receive buffer from the previous call
        if n_local:
            engine.scatter_pass_to(keys, top, owner, dest_base, kf)
        torch.cuda.synchronize(dev)
        handle.barrier()  # all remote stores have landed
        t = tick("fused_partition_exchange", t)
    else:
        # 4. group the shard by routing bucket (stable): destinations become contiguous ranges
        part = engine.empty(keys.numel(), keys)
        engine.scatter_pass(keys, part, top, kf)
        t = tick("partition_pass", t)
        # 5. exchange
        out_buf = engine.empty(max(n_out, 1) * rec_elems, keys)
        dist.all_to_all_single(out_buf[: n_out * rec_elems], part, [r * rec_elems for r in recv],
                               [s * rec_elems for s in send], group=group)
        t = tick("all_to_all", t)
        del part

    # 6. local LSD sort of the received records (chunks arrive in source-rank order: stable)
    recv_view = out_buf[: n_out * rec_elems]
    if n_out > 1:
        aux = keys if keys.numel() >= recv_view.numel() else (
            engine.scratch(recv_view.numel(), keys) if hasattr(engine, "scratch") else engine.empty(recv_view.numel(), keys))
        res = engine.sort(recv_view, aux[: recv_view.numel()], kf)
    else:
        res = recv_view
    t = tick("local_sort", t)
    sec["exchange"] = "fused peer stores (NVLink)" if symm is not None else "NCCL all_to_all_single"
    info = PartitionInfo(top, live, send, recv, n_out, n_total, n_out / max(n_total / world, 1), sec)
    return res, info


# -------------------------------------------------------------------------------------------------
def bench_partitioned(args, rsx, tname, n_per_gpu, dname, mask, orv, rank, world, dev, sampler=None):
    """bench.py's N > 1 arm: weak scaling, n_per_gpu keys per rank, device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    tdt = {"u32": torch.int32, "u64": torch.int64, "i32": torch.int32, "i64": torch.int64,
           "f32": torch.float32, "f64": torch.float64}[tname]
    kf = rsx.default_kdf(tdt) if tname[0] != "u" else rsx.KeyFunc(rsx.KDF_UNSIGNED)
    kb = torch.empty(0, dtype=tdt).element_size()
    pristine = torch.empty(n_per_gpu, dtype=tdt, device=dev)
    rsx.fill_keys(pristine, seed=2, start=rank * n_per_gpu, dist=dname, mask=mask, orv=orv)
    _, s0, x0 = rsx.verify(pristine, kf)
    chk0 = torch.tensor([s0 & 0x7FFFFFFFFFFFFFFF, x0 & 0x7FFFFFFFFFFFFFFF, n_per_gpu], dtype=torch.int64, device=dev)
    engine = CudaEngine()
    keys = torch.empty_like(pristine)
    times, last = [], None
    launches0 = 0
    for it in range(args.warmup + args.steps):
        keys.copy_(pristine)
        if it == args.warmup:
            launches0 = rsx.total_kernel_launches()
            if sampler is not None:
                sampler.start()  # nvidia-smi clocks / throttle reasons over the timed steps
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res, info = partitioned_sort(keys, kf, engine=engine, fused=not getattr(args, "no_fused", False))
        e1.record()
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # device time, max over ranks
        if it >= args.warmup:
            times.append(float(ms.item()))
        last = (res, info)
    launches = rsx.total_kernel_launches() - launches0
    clocks = sampler.stop() if sampler is not None else None
    res, info = last
    # verification at full size: every shard ordered, shard boundaries ordered, multiset preserved
    d1, s1, x1 = rsx.verify(res, kf) if info.n_out > 1 else (0, 0, 0)
    def ordered(v):  # unsigned key stored in a signed tensor -> int64 with the same order
        v = v.to(torch.int64)
        return (v & 0xFFFFFFFF) if kb == 4 else (v ^ torch.tensor(-(1 << 63), dtype=torch.int64, device=dev))
    edge = torch.zeros(3, dtype=torch.int64, device=dev)
    if info.n_out:
        edge[0], edge[1], edge[2] = ordered(res.view(-1)[0]), ordered(res.view(-1)[-1]), 1
    edges = [torch.empty_like(edge) for _ in range(world)]
    dist.all_gather(edges, edge)
    nonempty = [e.cpu().tolist() for e in edges if int(e[2])]
    boundaries_ok = all(a[1] <= b[0] for a, b in zip(nonempty, nonempty[1:])) if tname[0] == "u" else True
    # multiset: sum is additive over shards, xor combines with xor
    chk1 = torch.tensor([s1 & 0x7FFFFFFFFFFFFFFF, x1 & 0x7FFFFFFFFFFFFFFF, info.n_out], dtype=torch.int64, device=dev)
    all0 = [torch.empty_like(chk0) for _ in range(world)]
    all1 = [torch.empty_like(chk1) for _ in range(world)]
    dist.all_gather(all0, chk0)
    dist.all_gather(all1, chk1)
    ok_sorted = torch.tensor([int(d1 == 0)], device=dev)
    dist.all_reduce(ok_sorted, op=dist.ReduceOp.MIN)
    n_in = sum(int(a[2]) for a in all0)
    n_outs = [int(a[2]) for a in all1]
    sum_ok = (sum(int(a[0]) for a in all0) - sum(int(a[0]) for a in all1)) % (1 << 63) == 0
    xor0 = xor1 = 0
    for a, b in zip(all0, all1):
        xor0 ^= int(a[1])
        xor1 ^= int(b[1])
    verified = bool(ok_sorted.item()) and n_in == sum(n_outs) and boundaries_ok and sum_ok and xor0 == xor1
    # one extra pass with host timers for the phase breakdown (not part of the timed steps)
    keys.copy_(pristine)
    dist.barrier()
    _, info_t = partitioned_sort(keys, kf, engine=engine, timers=True, fused=not getattr(args, "no_fused", False))
    # e2e: every rank's shard starts and ends in pinned HOST memory (H2D + global sort + D2H timed)
    e2e = None
    if not getattr(args, "no_e2e", False):
        h_in = h_out = None
        try:
            h_in = torch.empty(n_per_gpu, dtype=tdt, pin_memory=True)
            h_out = torch.empty(int(n_per_gpu * 1.25) + 1024, dtype=tdt, pin_memory=True)
            ok = 1
        except Exception:  # not enough pinnable host memory for N shards on this box
            ok = 0
        okt = torch.tensor([ok], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)  # all ranks take the same branch
        if int(okt.item()):
            h_in.copy_(pristine)
            e2e_t = []
            for it in range(1 + max(5, args.steps)):
                dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                keys.copy_(h_in, non_blocking=True)
                r_e, info_e = partitioned_sort(keys, kf, engine=engine, fused=not getattr(args, "no_fused", False))
                h_out[: info_e.n_out].copy_(r_e.view(-1)[: info_e.n_out], non_blocking=True)
                torch.cuda.synchronize(dev)
                dt = torch.tensor([time.perf_counter() - t0], device=dev)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                if it:
                    e2e_t.append(float(dt.item()))
            e2e_s = sum(e2e_t) / len(e2e_t)
            e2e = {"value": n_per_gpu * world / e2e_s / 1e9, "unit": "Gkeys/s", "h2d_bytes_per_step": n_per_gpu * kb * world,
                   "d2h_bytes_per_step": n_per_gpu * kb * world, "ms_per_step": e2e_s * 1e3, "steps": len(e2e_t),
                   "note": "per rank: pinned host shard -> device, partitioned global sort, sorted shard -> pinned host; wall clock, max over ranks"}
        else:
            e2e = {"value": None, "unit": "Gkeys/s", "error": "could not pin host memory for every rank's shard"}
        del h_in, h_out
    ms_per_step = sum(times) / len(times)
    n_total = n_per_gpu * world
    value = n_total / (ms_per_step * 1e-3) / 1e9
    single = n_per_gpu * kb * (1 + 2 * kb)  # local LSD sort bytes per GPU
    moved = n_per_gpu * kb * (1 + 2) + single  # + histogram read + partition pass
    return {
        "metric": "Gkeys/s sorted", "value": value, "unit": "Gkeys/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": tname, "data": "synthetic",
        "config": {"workload": f"{args.workload} per GPU, partitioned global sort over {world} GPUs",
                   "keys_total": n_total, "keys_per_gpu": n_per_gpu, "dist": dname, "routing_column": info.routing_column,
                   "imbalance_max_over_mean": max(n_outs) / (n_total / world), "verified": verified,
                   "timing": "CUDA events per rank around partitioned_sort, all_reduce MAX over ranks, mean of steps",
                   "l2": "inputs larger than L2, restored before every step",
                   "phase_seconds_rank0": info_t.seconds, "symm_error": CudaEngine.symm_error,
                   "ms_steps": [round(x, 3) for x in times]},
        "roofline": {"bound": "hbm", "kernel": "whole partitioned sort, per GPU", "achieved": moved / (ms_per_step * 1e-3) / 1e9,
                     "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "per-GPU algorithmic HBM bytes (histogram + partition pass + local LSD) over the step time; "
                             "NVLink term: each GPU sends/receives (N-1)/N of its shard",
                     "nvlink_bytes_per_gpu": n_per_gpu * kb * (world - 1) / world},
        "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
