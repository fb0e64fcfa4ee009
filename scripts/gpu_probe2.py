import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtc_b200
from rtc_b200 import scenes
ctx = rtc_b200.Context(0)
res = {}
for v in (0, 1, 2, 3):
    res["peak_variant%d_tflops" % v] = max(ctx.fp32_peak(v, 3000 if v < 3 else 6000)[0] for _ in range(3))
name = "config3_4k_1024"
p = scenes.config_camera(name)
objs = scenes.config_scene(name)
rays = (p.x - 1) * p.y
def run(tag, o):
    ctx.set_objects(o)
    for _ in range(4):
        ctx.render(p, rtc_b200.RGB_PIXEL); ctx.frame_ansi_device()
    t = ctx.timings(); n = int((o["type"] == 2).sum())
    res[tag] = dict(trace_ms=t["trace_ms"], shade_ms=t["shade_ms"], encode_ms=t["encode_ms"], tflops=7.0 * rays * n / (t["trace_ms"] * 1e-3) / 1e12)
run("config3", objs)
far = objs.copy(); far["center"][:, 2] -= 1000.0        # everything behind the camera: no candidates at all
run("config3_allmiss", far)
sp = objs[objs["type"] == 2]
run("config3_noplane", sp)
print(json.dumps(res, indent=1))
