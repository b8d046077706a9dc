#!/usr/bin/env python
"""Scatter-kernel geometry sweep: device time per pass (CUDA events inside librsx, "profile" option)
for every tuning variant, on 1 B u32 and 1 B u64 uniform keys, each result verified (sorted +
multiset checksum).  Writes gpurun_out/variants.json.   usage: sweep_variants.py [n] [variants...]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rsx = importlib.import_module("radix-sorting_b200")
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
variants = [int(v) for v in sys.argv[2:]] or [0] + list(range(10, 30))
U = rsx.KeyFunc(rsx.KDF_UNSIGNED)
rows = []
for tname, tdt, dist in (("u32", torch.int32, "uniform"), ("u64", torch.int64, "uniform"), ("u32", torch.int32, "and3")):
    kb = 4 if tname == "u32" else 8
    pristine = torch.empty(n, dtype=tdt, device=dev)
    rsx.fill_keys(pristine, seed=2, dist=dist)
    src, aux = torch.empty_like(pristine), torch.empty_like(pristine)
    _, s0, x0 = rsx.verify(pristine, U)
    rsx.set_profile(True)
    for v in variants:
        if rsx.lib().rsx_set_option(b"scatter_variant", v) != 0:
            continue
        try:
            rsx.reserve(rsx.workspace_bytes(n, U.layout(kb)))
            tot, passes = [], []
            for it in range(4):
                src.copy_(pristine)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                rep = rsx.RsxReport()
                e0.record(); res = rsx.radix_sort(src, aux, None, U, report=rep); e1.record(); e1.synchronize()
                if it:
                    tot.append(e0.elapsed_time(e1)); p = rsx.get_profile(); passes += [x for x in p[2:2 + kb] if x > 0]
            d1, s1, x1 = rsx.verify(res, U)
            ok = d1 == 0 and (s1, x1) == (s0, x0)
            row = {"type": tname, "dist": dist, "variant": v, "sort_ms": sum(tot) / len(tot), "pass_ms": sum(passes) / len(passes),
                   "pass_GBps": 2 * n * kb / (sum(passes) / len(passes)) / 1e6, "verified": bool(ok)}
        except Exception as e:  # noqa: BLE001
            row = {"type": tname, "dist": dist, "variant": v, "error": repr(e)[:200]}
        rows.append(row)
        print(row, flush=True)
    rsx.lib().rsx_set_option(b"scatter_variant", 0)
    del pristine, src, aux
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)
