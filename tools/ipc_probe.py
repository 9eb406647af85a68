"""Probe (torchrun, 2+ GPUs): can one process map another process' device memory (cudaIpc) and read/write it over NVLink?"""
import ctypes as C
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
rt = C.CDLL("libcudart.so.12")


class Handle(C.Structure):      # cudaIpcMemHandle_t is passed BY VALUE to cudaIpcOpenMemHandle
    _fields_ = [("reserved", C.c_char * 64)]


rt.cudaIpcGetMemHandle.argtypes = [C.POINTER(Handle), C.c_void_p]
rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), Handle, C.c_uint]
buf = torch.full((1 << 20,), float(rank + 1), device="cuda")
h = Handle()
rc = rt.cudaIpcGetMemHandle(C.byref(h), C.c_void_p(buf.data_ptr()))
print(rank, "cudaIpcGetMemHandle rc", rc, flush=True)
box = [None] * world
dist.all_gather_object(box, bytes(h))
peers = []
for r in range(world):
    if r == rank:
        peers.append(buf.data_ptr())
        continue
    p = C.c_void_p()
    hh = Handle.from_buffer_copy(box[r])
    rc = rt.cudaIpcOpenMemHandle(C.byref(p), hh, C.c_uint(1))
    print(rank, "open handle of", r, "rc", rc, hex(p.value or 0), flush=True)
    peers.append(p.value)
dist.barrier()
# read the peer's buffer with a device-to-device copy through the mapped pointer
nxt = (rank + 1) % world
dst = torch.empty(1 << 20, device="cuda")
rc = rt.cudaMemcpy(C.c_void_p(dst.data_ptr()), C.c_void_p(peers[nxt]), C.c_size_t(4 << 20), C.c_int(3))
torch.cuda.synchronize()
print(rank, "peer read rc", rc, "value", float(dst[0]), "expected", float(nxt + 1), flush=True)
can = C.c_int()
rt.cudaDeviceCanAccessPeer(C.byref(can), C.c_int(torch.cuda.current_device()), C.c_int(nxt))
print(rank, "canAccessPeer", can.value, flush=True)
dist.barrier()
dist.destroy_process_group()
