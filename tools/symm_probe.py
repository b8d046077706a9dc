import os, sys, traceback
import torch, torch.distributed as dist
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    print(rank, "symm_mem attrs:", [a for a in dir(symm_mem) if not a.startswith("_")][:40], flush=True)
    try:
        symm_mem.enable_symm_mem_for_group(dist.group.WORLD.group_name)
    except Exception as e:
        print(rank, "enable_symm_mem_for_group:", repr(e), flush=True)
    t = symm_mem.empty(1<<20, dtype=torch.int32, device=dev)
    h = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, "rendezvous ok", type(h), [hex(int(p)) for p in h.buffer_ptrs], flush=True)
    t.fill_(rank+1)
    h.barrier()
    torch.cuda.synchronize()
    peer = h.get_buffer((rank+1)%world, (16,), torch.int32)
    print(rank, "peer view", peer[:4].tolist(), flush=True)
except Exception as e:
    print(rank, "SYMM FAILED", repr(e), flush=True); traceback.print_exc()
# raw CUDA IPC via cuda-python / ctypes
try:
    import ctypes as C
    cudart = C.CDLL("libcudart.so.12")
    buf = torch.full((1<<20,), rank+10, dtype=torch.int32, device=dev)
    class H(C.Structure): _fields_=[("r", C.c_char*64)]
    h = H()
    # caching allocator sub-allocates; IPC handle of the base allocation + offset
    st = cudart.cudaIpcGetMemHandle(C.byref(h), C.c_void_p(buf.data_ptr()))
    print(rank, "cudaIpcGetMemHandle status", st, flush=True)
    handles=[None]*world
    dist.all_gather_object(handles, bytes(h.r))
    peerh = H(); C.memmove(C.byref(peerh), handles[(rank+1)%world], 64)
    p = C.c_void_p()
    st = cudart.cudaIpcOpenMemHandle(C.byref(p), peerh, C.c_uint(1))
    print(rank, "cudaIpcOpenMemHandle status", st, hex(p.value or 0), flush=True)
    can = C.c_int(0); cudart.cudaDeviceCanAccessPeer(C.byref(can), local, (local+1)%world)
    print(rank, "canAccessPeer", can.value, flush=True)
except Exception as e:
    print(rank, "IPC FAILED", repr(e), flush=True)
dist.barrier(); dist.destroy_process_group()
