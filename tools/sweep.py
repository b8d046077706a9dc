#!/usr/bin/env python
"""N2: the reference's Google-Benchmark sweep n = 1 .. 40 M (x10) for radix_sort and
radix_sort_rank on u32 keys (radix_bench.cpp:86-138), on the device path, input restored before
every iteration (the reference's loop does not, radix_bench.cpp:91-93).  Prints KeyRate per n for
device-resident buffers (CUDA events) and for the whole synchronous call (wall clock)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rsx = importlib.import_module("radix-sorting_b200")
dev = torch.device("cuda", 0)
U = rsx.KeyFunc(rsx.KDF_UNSIGNED)
rows = []
n = 1
while n <= 40_000_000:
    pristine = torch.empty(n, dtype=torch.int32, device=dev); rsx.fill_keys(pristine, seed=5)
    src = torch.empty_like(pristine); aux = torch.empty_like(pristine); ib = torch.empty(2 * n, dtype=torch.int32, device=dev)
    res = {}
    for name, fn in (("radix_sort", lambda: rsx.radix_sort(src, aux, None, U)), ("radix_sort_rank", lambda: rsx.radix_sort_rank(src, ib, n, U))):
        ev, wall = [], []
        for it in range(12):
            src.copy_(pristine); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record(); fn(); e1.record(); e1.synchronize(); t1 = time.perf_counter()
            if it >= 2:
                ev.append(e0.elapsed_time(e1) * 1e-3); wall.append(t1 - t0)
        res[name] = {"device_us": min(ev) * 1e6, "wall_us": min(wall) * 1e6, "KeyRate_Mkeys_s": n / min(wall) / 1e6}
    rows.append({"n": n, **res})
    print(n, {k: {a: round(b, 2) for a, b in v.items()} for k, v in res.items()}, flush=True)
    n *= 10 if n < 10_000_000 else 4
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
