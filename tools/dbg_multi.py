import os, sys, importlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
rsx = importlib.import_module("radix-sorting_b200"); dsort = importlib.import_module("radix-sorting_b200.dist")
n=int(sys.argv[1]); dname=sys.argv[2]
kf = rsx.KeyFunc(rsx.KDF_UNSIGNED)
pristine = torch.empty(n, dtype=torch.int64, device=dev); rsx.fill_keys(pristine, seed=2, start=rank*n, dist=dname)
keys = torch.empty_like(pristine); eng = dsort.CudaEngine()
for it in range(5):
    keys.copy_(pristine); dist.barrier(); torch.cuda.synchronize()
    t0=time.perf_counter()
    res, info = dsort.partitioned_sort(keys, kf, engine=eng, timers=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(rank, it, round(dt*1e3,1), {k:(round(v*1e3,1) if isinstance(v,float) else v) for k,v in info.seconds.items()}, "n_out", info.n_out, "mem", round(torch.cuda.memory_allocated()/1e9,1), flush=True)
dist.destroy_process_group()
