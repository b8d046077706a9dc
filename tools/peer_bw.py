"""Bench-only probe: bandwidth of SM-issued stores into a peer GPU's memory (symmetric memory),
vectorised (16 B/lane) vs scalar (4 B/lane, misaligned view), next to a copy-engine peer copy."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 29  # 2 GiB of int32
t = symm_mem.empty(n + 64, dtype=torch.int32, device=dev)
h = symm_mem.rendezvous(t, group=dist.group.WORLD)
peer = h.get_buffer((rank + 1) % world, (n + 64,), torch.int32)
src = torch.arange(n, dtype=torch.int32, device=dev)
def timeit(fn, name, bytes_):
    h.barrier(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    h.barrier()
    if rank == 0: print(f"{name:50s} {best:8.3f} ms  {bytes_/best/1e6:8.1f} GB/s", flush=True)
timeit(lambda: peer[:n].copy_(src), "copy_ to peer (aligned)", 4*n)
timeit(lambda: torch.add(src, 1, out=peer[:n]), "elementwise add -> peer, aligned (16 B/lane)", 4*n)
timeit(lambda: torch.add(src[: n - 1], 1, out=peer[1:n]), "elementwise add -> peer, misaligned (4 B/lane)", 4*n)
timeit(lambda: torch.add(src, 1, out=t[:n]), "elementwise add -> local (for reference)", 4*n)
dist.destroy_process_group()
