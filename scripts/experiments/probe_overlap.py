"""Probe: does a D2H copy into cudaHostRegister'ed POSIX shm overlap with kernels like one into cudaMallocHost memory?"""
import ctypes, time
from multiprocessing import shared_memory
import torch
size = 64 << 20
shm = shared_memory.SharedMemory(create=True, size=size)
addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
for flag in (1, 3):     # portable ; portable|mapped
    print("cudaHostRegister flags", flag, torch.cuda.cudart().cudaHostRegister(addr, size, flag))
    host_shm = torch.frombuffer(shm.buf, dtype=torch.uint8, count=size)
    host_pin = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    src = torch.randint(0, 255, (size,), dtype=torch.uint8, device="cuda")
    a = torch.randn(4096, 4096, device="cuda"); b = torch.randn(4096, 4096, device="cuda")
    comp, cp = torch.cuda.Stream(), torch.cuda.Stream()
    n = 10 << 20
    def work():
        with torch.cuda.stream(comp):
            for _ in range(6):
                torch.mm(a, b)
    for name, host in (("pinned", host_pin), ("shm", host_shm)):
        for _ in range(3):
            work(); torch.cuda.synchronize()
        t0 = time.perf_counter(); work(); comp.synchronize(); t_k = time.perf_counter() - t0
        t0 = time.perf_counter()
        with torch.cuda.stream(cp):
            host[:n].copy_(src[:n], non_blocking=True)
        cp.synchronize(); t_c = time.perf_counter() - t0
        t0 = time.perf_counter(); work()
        with torch.cuda.stream(cp):
            host[:n].copy_(src[:n], non_blocking=True)
        t_issue = time.perf_counter() - t0
        cp.synchronize(); t_cdone = time.perf_counter() - t0
        comp.synchronize(); t_both = time.perf_counter() - t0
        print("%-7s kernels %.3f ms, copy %.3f ms, both %.3f ms (issue %.3f, copy done at %.3f)" % (name, t_k*1e3, t_c*1e3, t_both*1e3, t_issue*1e3, t_cdone*1e3))
    del host_shm
    torch.cuda.cudart().cudaHostUnregister(addr)
shm.close(); shm.unlink()
