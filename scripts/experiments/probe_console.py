"""The reference's own console sizes (240x64, 400x150, default scene): frame time with 8 and 4 rays per thread."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) < 2:
    for v in ("8", "4", "auto"):
        env = dict(os.environ)
        if v != "auto":
            env["RTC_TRACE_RAYS_FORCE"] = v
        print("rays per thread:", v, flush=True)
        subprocess.run([sys.executable, __file__, v], env=env)
    sys.exit(0)
import rtc_b200  # noqa: E402
from rtc_b200 import scenes  # noqa: E402

ctx = rtc_b200.Context(0)
for name in ("config1_240x64", "config1_400x150"):
    ctx.set_objects(scenes.default_scene())
    p = scenes.config_camera(name)
    for _ in range(20):
        ctx.render(p, rtc_b200.RGB_PIXEL)
        ctx.frame_ansi_device()
    acc = {"trace_ms": 0.0, "total_ms": 0.0}
    for _ in range(50):
        ctx.render(p, rtc_b200.RGB_PIXEL)
        ctx.frame_ansi_device()
        t = ctx.timings()
        for k in acc:
            acc[k] += t[k] / 50
    print("  %-16s trace %.4f ms, frame (4 launches) %.4f ms" % (name, acc["trace_ms"], acc["total_ms"]), flush=True)
