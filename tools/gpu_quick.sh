#!/bin/bash
# quick A/B: parity tests + a few device-time lines
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import importlib, sys, os, json
sys.path.insert(0, os.getcwd())
import torch
rsx = importlib.import_module("radix-sorting_b200")
dev = torch.device("cuda", 0)
U = rsx.KeyFunc(rsx.KDF_UNSIGNED)
def run(name, dtype, n, dist="uniform", mask=(1<<64)-1):
    pristine = torch.empty(n, dtype=dtype, device=dev); rsx.fill_keys(pristine, seed=11, dist=dist, mask=mask)
    src = torch.empty_like(pristine); aux = torch.empty_like(pristine)
    _, s0, x0 = rsx.verify(pristine, U)
    rsx.set_profile(True)
    best = 1e9; prof=None
    for r in range(5):
        src.copy_(pristine)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = rsx.radix_sort(src, aux, None, U); e1.record(); e1.synchronize()
        if r and e0.elapsed_time(e1) < best: best = e0.elapsed_time(e1); prof = rsx.get_profile()
    d1, s1, x1 = rsx.verify(res, U)
    print(f"{name:34s} {best:8.3f} ms  {n/best/1e6:7.2f} Gkeys/s  ok={d1==0 and (s1,x1)==(s0,x0)}  K1={prof[0]:.3f} passes={[round(x,3) for x in prof[2:] if x>0.01]}", flush=True)
    del pristine, src, aux; torch.cuda.empty_cache()
N=1_000_000_000
run("1B u32 uniform", torch.int32, N)
run("1B u64 uniform", torch.int64, N)
run("1B u32 and3", torch.int32, N, "and3")
run("1B u32 and2", torch.int32, N, "and2")
run("1B u64 and4", torch.int64, N, "and4")
run("1B u32 zipf", torch.int32, N, "zipf")
run("1B u32 & 0x0F0F0F0F", torch.int32, N, "uniform", 0x0F0F0F0F)
PY
