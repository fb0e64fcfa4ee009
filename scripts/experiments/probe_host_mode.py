"""Probe: GPU-side timeline of HostAssembledRenderer.submit/collect at world=1 (frame durations and gaps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import rtc_b200
from rtc_b200 import multigpu, scenes
name = "config3_4k_1024"
ctx = rtc_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
objs = scenes.config_scene(name); p = scenes.config_camera(name)
ctx.set_objects(objs)
r = multigpu.HostAssembledRenderer(ctx, None, 0, 1, p.x, p.y, rtc_b200.RGB_PIXEL)
variant = sys.argv[1] if len(sys.argv) > 1 else "full"
N = 12
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N + 1)]
def submit(i):
    ctx.set_objects(objs)
    ev[i][0].record(st)
    s = r.submit(p)
    ev[i][1].record(st)
    return s
for _ in range(3):
    submit(0); r.collect()
torch.cuda.synchronize()
t0 = time.perf_counter()
submit(0)
for i in range(N):
    submit(i + 1)
    r.collect()
t1 = time.perf_counter()
r.collect()
torch.cuda.synchronize()
print(variant, "wall per frame %.3f ms" % ((t1 - t0) * 1e3 / N))
print("frame gpu ms:", ["%.3f" % ev[i][0].elapsed_time(ev[i][1]) for i in range(N)])
print("gap to next :", ["%.3f" % ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(N)])
r.close()
