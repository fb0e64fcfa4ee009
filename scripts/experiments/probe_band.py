"""Probe: time rtc_trace_band (hoist + trace + shade) for an 8-GPU-sized band of config 3 and for whole frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import rtc_b200
from rtc_b200 import scenes
ctx = rtc_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
for name, bands in (("config3_4k_1024", [(0, 2160), (945, 1215), (945, 1216), (540, 1080)]), ("config2_1080p_64", [(0, 1080)]), ("config4_8k_4096", [(0, 4320), (1890, 2430)])):
    objs = scenes.config_scene(name); p = scenes.config_camera(name)
    ctx.set_objects(objs)
    W = p.x - 1
    color = torch.empty(W * p.y * 3 + 64, dtype=torch.uint8, device="cuda")
    for (r0, r1) in bands:
        for _ in range(3):
            ctx.trace_band(p, rtc_b200.RGB_PIXEL, r0, r1, color.data_ptr(), 0)
        torch.cuda.synchronize()
        n = 20 if name != "config4_8k_4096" or r1 - r0 < 4000 else 5
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(n):
            ctx.trace_band(p, rtc_b200.RGB_PIXEL, r0, r1, color.data_ptr(), 0)
        b.record(st); torch.cuda.synchronize()
        print(os.environ.get("RTC_TRACE_THREADS_FORCE", "auto"), name, (r0, r1), "%.4f ms" % (a.elapsed_time(b) / n))
