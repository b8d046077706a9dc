#!/usr/bin/env python
"""Device-time table for the BASELINE configs that are not the headline bench line: key types via
KDFs, column skipping, presorted input, rank sort and key+payload records (configs 1-4), each
verified at full size by size-independent properties.  Writes gpurun_out/configs.json."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
rsx = importlib.import_module("radix-sorting_b200")
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
M64 = (1 << 64) - 1
out = []

def timed(fn, restore, reps=4):
    best = 1e30
    for r in range(reps):
        restore()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = fn(); e1.record(); e1.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best, res

def value_sort(name, dtype, n, kf, dist="uniform", mask=M64, orv=0, passes=None, alg=None):
    pristine = torch.empty(n, dtype=dtype, device=dev); rsx.fill_keys(pristine, seed=11, dist=dist, mask=mask, orv=orv)
    src = torch.empty_like(pristine); aux = torch.empty_like(pristine)
    _, s0, x0 = rsx.verify(pristine, kf)
    rep = rsx.RsxReport()
    ms, res = timed(lambda: rsx.radix_sort(src, aux, None, kf, report=rep), lambda: src.copy_(pristine))
    d1, s1, x1 = rsx.verify(res, kf)
    ok = d1 == 0 and (s1, x1) == (s0, x0)
    kb = pristine.element_size()
    alg_bytes = n * kb * (1 + 2 * rep.ncols)
    out.append({"config": name, "n": n, "live_passes": rep.ncols, "early_exit": rep.early_exit, "ms": ms, "Gkeys_s": n / ms / 1e6,
                "alg_GBps": alg_bytes / ms / 1e6, "frac_of_measured_peak": alg_bytes / ms / 1e6 / PEAK, "verified": bool(ok)})
    print(out[-1], flush=True)
    del pristine, src, aux; torch.cuda.empty_cache()

U = rsx.KeyFunc(rsx.KDF_UNSIGNED)
N1 = 1_000_000_000
value_sort("C1 40M u32 uniform", torch.int32, 40_000_000, U)
value_sort("C2a 1B u32 uniform", torch.int32, N1, U)
value_sort("C2b 1B u64 uniform", torch.int64, N1, U)
value_sort("C2c 1B u32 & 0x00FFFFFF (3 live)", torch.int32, N1, U, mask=0x00FFFFFF)
value_sort("C2c 1B u32 & 0x0000FFFF (2 live)", torch.int32, N1, U, mask=0x0000FFFF)
value_sort("C2d 1B u64 40 live bits | const top (5 live)", torch.int64, N1, U, mask=0x000000FFFFFFFFFF, orv=0xAA00000000000000)
value_sort("C2d 1B u64 & 0xFFFFFFFF (4 live)", torch.int64, N1, U, mask=0xFFFFFFFF)
value_sort("C2e 1B u32 and3 (low entropy)", torch.int32, N1, U, dist="and3")
value_sort("C2e 1B u64 and4 (low entropy)", torch.int64, N1, U, dist="and4")
value_sort("C2f 1B u32 presorted (early exit)", torch.int32, N1, U, dist="sorted")
value_sort("C2f 1B u32 constant (early exit)", torch.int32, N1, U, dist="constant")
value_sort("C3a 500M f32 random bit patterns", torch.float32, 500_000_000, rsx.default_kdf(torch.float32))
value_sort("C3b 500M i64 uniform", torch.int64, 500_000_000, rsx.default_kdf(torch.int64))
value_sort("C3 500M f64 random bit patterns", torch.float64, 500_000_000, rsx.default_kdf(torch.float64))
value_sort("1B u32 descending", torch.int32, N1, rsx.KeyFunc(rsx.KDF_UNSIGNED, True))

# ---- C4a: rank sort, 1B u32 keys, u32 indices ---------------------------------------------------
for name, mask in [("C4a 1B u32 rank sort (u32 idx)", M64), ("C4a 1B u32 & 0xFFFFF rank sort (heavy ties)", 0x000FFFFF)]:
    n = N1
    keys = torch.empty(n, dtype=torch.int32, device=dev); rsx.fill_keys(keys, seed=6, mask=mask)
    ib = torch.empty(2 * n, dtype=torch.int32, device=dev)
    rep = rsx.RsxReport()
    ms, ranks = timed(lambda: rsx.radix_sort_rank(keys, ib, n, U, report=rep), lambda: None)
    g = keys[ranks.long()] if False else torch.gather(keys, 0, ranks.to(torch.int64))
    gu = g.to(torch.int64) & 0xFFFFFFFF
    sorted_ok = bool((gu[1:] >= gu[:-1]).all())
    stable_ok = bool(((gu[1:] != gu[:-1]) | (ranks[1:] > ranks[:-1])).all())
    seen = torch.zeros(n, dtype=torch.bool, device=dev); seen[ranks.to(torch.int64)] = True
    perm_ok = bool(seen.all())
    P = rep.ncols
    alg = n * (4 + (4 + 8) + max(P - 2, 0) * 16 + (8 + 4)) if P >= 2 else n * (4 + 4 + 4)
    out.append({"config": name, "n": n, "live_passes": P, "ms": ms, "Gkeys_s": n / ms / 1e6, "alg_GBps": alg / ms / 1e6,
                "frac_of_measured_peak": alg / ms / 1e6 / PEAK, "verified": sorted_ok and stable_ok and perm_ok})
    print(out[-1], flush=True)
    del keys, ib, g, gu, seen; torch.cuda.empty_cache()

# ---- C4b: {u32 key, u32 payload} records -------------------------------------------------------------
for name, mask in [("C4b 1B {u32 key,u32 payload} records", M64), ("C4b 1B records, key & 0xFFFFF (heavy ties)", 0x000FFFFF)]:
    n = N1
    k = torch.empty(n, dtype=torch.int32, device=dev); rsx.fill_keys(k, seed=6, mask=mask)
    recs = torch.empty(n, 2, dtype=torch.int32, device=dev)
    recs[:, 0] = k; recs[:, 1] = torch.arange(n, dtype=torch.int32, device=dev)
    del k
    pristine = recs.reshape(-1).clone(); src = recs.reshape(-1); aux = torch.empty_like(src)
    kf = rsx.KeyFunc(rsx.KDF_UNSIGNED, False, 8, 0, 4)
    rep = rsx.RsxReport()
    ms, res = timed(lambda: rsx.radix_sort(src, aux, None, kf, report=rep), lambda: src.copy_(pristine))
    r2 = res.view(n, 2); ku = r2[:, 0].to(torch.int64) & 0xFFFFFFFF; pay = r2[:, 1]
    sorted_ok = bool((ku[1:] >= ku[:-1]).all()); stable_ok = bool(((ku[1:] != ku[:-1]) | (pay[1:] > pay[:-1])).all())
    orig_key = pristine.view(n, 2)[:, 0]
    carried_ok = bool((torch.gather(orig_key, 0, pay.to(torch.int64)) == r2[:, 0]).all())
    alg = n * (8 + rep.ncols * 16)
    out.append({"config": name, "n": n, "live_passes": rep.ncols, "ms": ms, "Gkeys_s": n / ms / 1e6, "alg_GBps": alg / ms / 1e6,
                "frac_of_measured_peak": alg / ms / 1e6 / PEAK, "verified": sorted_ok and stable_ok and carried_ok})
    print(out[-1], flush=True)
    del recs, pristine, src, aux, r2, ku, pay, orig_key; torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
