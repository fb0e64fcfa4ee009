"""Probe: latency of a small D2H into page-locked POSIX shm while the ray kernel runs (why is rank 0's d2h ~0.22 ms at any N?)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from multiprocessing import shared_memory
import torch
import rtc_b200
from rtc_b200 import scenes
ctx = rtc_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
name = "config3_4k_1024"
objs = scenes.config_scene(name); p = scenes.config_camera(name)
ctx.set_objects(objs)
W = p.x - 1
color = torch.empty(W * p.y * 3 + 64, dtype=torch.uint8, device="cuda")
size = 64 << 20
shm = shared_memory.SharedMemory(create=True, size=size)
addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
print(torch.cuda.cudart().cudaHostRegister(addr, size, 1))
host_shm = torch.frombuffer(shm.buf, dtype=torch.uint8, count=size)
host_pin = torch.empty(size, dtype=torch.uint8, pin_memory=True)
src = torch.randint(0, 255, (size,), dtype=torch.uint8, device="cuda")
cp = torch.cuda.Stream()
def busy(rows):
    for _ in range(4):
        ctx.trace_band(p, rtc_b200.RGB_PIXEL, 0, rows, color.data_ptr(), 0)
for label, host in (("pinned", host_pin), ("shm", host_shm)):
    for n in (4096, 1 << 20, 3 << 20):
        for rows, what in ((0, "idle GPU"), (2160, "under 4 frames of trace")):
            torch.cuda.synchronize()
            ts = []
            for rep in range(5):
                if rows: busy(rows)
                t0 = time.perf_counter()
                with torch.cuda.stream(cp):
                    host[12345:12345 + n].copy_(src[:n], non_blocking=True)
                cp.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
                torch.cuda.synchronize()
            print("%-7s %8d B  %-24s  %s ms" % (label, n, what, " ".join("%.3f" % t for t in ts)))
del host_shm
torch.cuda.cudart().cudaHostUnregister(addr); shm.close(); shm.unlink()
