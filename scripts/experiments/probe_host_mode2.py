"""Probe: GPU-side timeline of HostAssembledRenderer.submit/collect under torchrun (2+ ranks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import rtc_b200
from rtc_b200 import multigpu, scenes
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
name = "config3_4k_1024"
ctx = rtc_b200.Context(lr)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
objs = scenes.config_scene(name); p = scenes.config_camera(name)
ctx.set_objects(objs)
r = multigpu.HostAssembledRenderer(ctx, dist, rank, world, p.x, p.y, rtc_b200.RGB_PIXEL)
N = 12
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N + 1)]
def submit(i):
    ctx.set_objects(objs)
    ev[i][0].record(st)
    s = r.submit(p)
    ev[i][1].record(st)
for _ in range(3):
    submit(0); r.collect()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
stamps = []
submit(0)
for i in range(N):
    a = time.perf_counter(); submit(i + 1); b = time.perf_counter(); r.collect(); c = time.perf_counter()
    stamps.append((b - a, c - b))
t1 = time.perf_counter()
r.collect()
torch.cuda.synchronize()
for q in range(world):
    dist.barrier()
    if q == rank:
        print("rank", rank, "wall per frame %.3f ms" % ((t1 - t0) * 1e3 / N), "band", r.r0, r.r1)
        print(" frame gpu ms:", ["%.3f" % ev[i][0].elapsed_time(ev[i][1]) for i in range(N)])
        print(" gap to next :", ["%.3f" % ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(N)])
        print(" host submit/collect ms:", ["%.2f/%.2f" % (x * 1e3, y * 1e3) for x, y in stamps])
        sys.stdout.flush()
r.close()
dist.destroy_process_group()
