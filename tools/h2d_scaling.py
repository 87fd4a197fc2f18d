#!/usr/bin/env python
"""Host-to-device copy bandwidth per GPU when N ranks upload at the same time, for ordinary pinned host memory and for
write-combined pinned memory (cudaHostAllocWriteCombined).  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/h2d_scaling.py
"""
import ctypes as C
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    rt = C.CDLL("libcudart.so.12")
    nbytes = 2 << 30
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res = {}
    for name, flags in (("pinned", 0), ("pinned_wc", 4)):   # cudaHostAllocWriteCombined = 0x04
        p = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(nbytes), C.c_uint(flags)) == 0
        C.memset(p, 1, 1 << 20)
        for rep in range(6):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            assert rt.cudaMemcpy(C.c_void_p(dev.data_ptr()), p, C.c_size_t(nbytes), 1) == 0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep:
                res.setdefault(name, []).append(nbytes / dt / 1e9)
        rt.cudaFreeHost(p)
    line = f"rank {rank}/{world}: " + ", ".join(f"{k} {min(v):.1f}-{max(v):.1f} GB/s" for k, v in res.items())
    print(line, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
