"""Concurrent pinned H2D rate of all ranks at once: ordinary pinned memory against write-combined pinned memory
(cudaHostAllocWriteCombined: no CPU-cache snoop on the DMA reads).  torchrun --nproc-per-node N tools/h2d_wc_probe.py"""
import ctypes
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl")
rt = ctypes.CDLL("libcudart.so.12")
NB = 120 << 20
dev = torch.empty(NB, dtype=torch.uint8, device="cuda")
res = {}
for name, flags in (("pinned", 0), ("write_combined", 4)):          # cudaHostAllocWriteCombined = 0x04
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB), ctypes.c_uint(flags)) == 0
    ctypes.memset(p, 7, NB)
    s = torch.cuda.current_stream().cuda_stream
    def copy():
        assert rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr()), p, ctypes.c_size_t(NB), 1, ctypes.c_void_p(s)) == 0
    for _ in range(3):
        copy()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(40):
        copy()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[name] = 40 * NB / dt / 1e9
    rt.cudaFreeHost(p)
    if world > 1:
        dist.barrier()
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    for k in ("pinned", "write_combined"):
        v = [o[k] for o in out]
        print(f"{k}: min rank {min(v):.1f} GB/s, sum {sum(v):.1f} GB/s over {world} ranks")
