"""Quick on-box probe: FP32 peak microbench + per-kernel timings of the bench configs."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtc_b200
from rtc_b200 import scenes

ctx = rtc_b200.Context(0)
info = ctx.device_info()
res = {"device": info}
for v in (0, 1):
    best = max(ctx.fp32_peak(v, 4000)[0] for _ in range(3))
    res["fp32_peak_tflops_variant%d" % v] = best
for name in ["config1_400x150", "config2_1080p_64", "config3_4k_1024", "config4_8k_4096"]:
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    ctx.set_objects(objs)
    ts = []
    for it in range(5):
        ctx.render(p, rtc_b200.RGB_PIXEL)
        _, n = ctx.frame_ansi_device()
        ts.append(ctx.timings())
    t = ts[-1]
    n_sph = int((objs["type"] == 2).sum())
    rays = (p.x - 1) * p.y
    res[name] = dict(t, stream_bytes=n, rays=rays, n_spheres=n_sph,
                     trace_tflops=7.0 * rays * n_sph / (t["trace_ms"] * 1e-3) / 1e12 if t["trace_ms"] > 0 else 0,
                     mrays_s=rays / (t["total_ms"] * 1e-3) / 1e6)
print(json.dumps(res, indent=1))
