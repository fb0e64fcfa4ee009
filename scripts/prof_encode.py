"""Driver for ncu on the encoder: config-5 worst case (random RGB) and a rendered config-3 frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtc_b200
from rtc_b200 import scenes
ctx = rtc_b200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
x, y = 7681, 4320
W = x - 1
g = torch.Generator(device="cuda"); g.manual_seed(5)
rgb = torch.randint(0, 256, (W * y * 3,), dtype=torch.uint8, device="cuda", generator=g)
cap = rtc_b200.encode_capacity(x, y, rtc_b200.RGB_PIXEL)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
total = torch.zeros(1, dtype=torch.int64, device="cuda")
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    ctx.encode(rgb.data_ptr(), 0, x, y, rtc_b200.RGB_PIXEL, out.data_ptr(), cap, total.data_ptr())
    b.record(st); torch.cuda.synchronize()
    n = int(total.item())
    print("worst case: %.3f ms, %d B out, %.1f GB/s algorithmic" % (a.elapsed_time(b), n, (3 * W * y + n) / a.elapsed_time(b) / 1e6))
# realistic: rendered config-4 frame colours
name = "config4_8k_4096"
ctx.set_objects(scenes.config_scene(name)); p = scenes.config_camera(name)
color = torch.empty(W * y * 3, dtype=torch.uint8, device="cuda")
ctx.trace_band(p, rtc_b200.RGB_PIXEL, 0, y, color.data_ptr(), 0)
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    ctx.encode(color.data_ptr(), 0, x, y, rtc_b200.RGB_PIXEL, out.data_ptr(), cap, total.data_ptr())
    b.record(st); torch.cuda.synchronize()
    n = int(total.item())
    print("rendered config4: %.3f ms, %d B out, %.1f GB/s algorithmic" % (a.elapsed_time(b), n, (3 * W * y + n) / a.elapsed_time(b) / 1e6))
