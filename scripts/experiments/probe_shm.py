"""Probe: can a POSIX shared-memory segment be page-locked (cudaHostRegister) and used as a D2H target?"""
import ctypes, os, time
from multiprocessing import shared_memory
import torch
os.system("df -h /dev/shm | tail -1")
size = 400 << 20
shm = shared_memory.SharedMemory(create=True, size=size)
addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
rt = torch.cuda.cudart()
t0 = time.time()
rc = rt.cudaHostRegister(addr, size, 1)   # cudaHostRegisterPortable
print("cudaHostRegister rc", rc, "%.3f s" % (time.time() - t0))
x = torch.randint(0, 255, (64 << 20,), dtype=torch.uint8, device="cuda")
lib = ctypes.CDLL("libcudart.so.12") if False else None
import numpy as np
host = torch.frombuffer(shm.buf, dtype=torch.uint8, count=size)
print("is_pinned:", host.is_pinned())
s = torch.cuda.Stream()
for n in (2 << 20, 20 << 20, 64 << 20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s):
        host[:n].copy_(x[:n], non_blocking=True)
    s.synchronize(); dt = time.perf_counter() - t0
    print("D2H %d MB: %.3f ms, %.1f GB/s" % (n >> 20, dt * 1e3, n / dt / 1e9), bool((host[:n] == x[:n].cpu()).all()))
rt.cudaHostUnregister(addr)
del host
shm.close(); shm.unlink()
