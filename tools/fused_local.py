"""Bench-only: the fused partition pass with LOCAL destinations (no NVLink) vs the plain pass."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
rsx = importlib.import_module("radix-sorting_b200")
dev = torch.device("cuda", 0)
U = rsx.KeyFunc(rsx.KDF_UNSIGNED)
for dt, col in ((torch.int32, 3), (torch.int64, 7)):
    n = 1_000_000_000
    src = torch.empty(n, dtype=dt, device=dev); rsx.fill_keys(src, seed=2)
    dst = torch.empty_like(src)
    for ndest in (2, 8):
        owner = (np.arange(256) * ndest // 256)
        hist, _, _ = rsx.histogram(src, U)
        counts = [int(hist[col][owner == d].sum()) for d in range(ndest)]
        offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
        bases = [dst.data_ptr() + int(o) * src.element_size() for o in offs]
        best = 1e9
        for r in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rsx.scatter_pass_to(src, col, owner, bases, U); e1.record(); e1.synchronize()
            if r: best = min(best, e0.elapsed_time(e1))
        print(dt, "fused pass, local destinations, ndest", ndest, round(best, 3), "ms (includes K1+K2 front)", flush=True)
    best = 1e9
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rsx.scatter_pass(src, dst, col, U); e1.record(); e1.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    print(dt, "plain pass (includes K1+K2 front)", round(best, 3), "ms", flush=True)
    del src, dst; torch.cuda.empty_cache()
