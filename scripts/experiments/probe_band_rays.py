"""Ray-kernel time of row bands of a 4K frame (what one of N GPUs traces) with 8 and with 4 rays per thread.
RTC_TRACE_RAYS_FORCE is read once per process, so each setting runs in its own process: python probe_band_rays.py [8|4|auto]"""
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
import torch  # noqa: E402
import rtc_b200  # noqa: E402
from rtc_b200 import scenes  # noqa: E402

ctx = rtc_b200.Context(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, rows_list in (("config3_4k_1024", (2160, 1081, 541, 271, 136)), ("config4_8k_4096", (4320, 541)), ("config2_1080p_64", (1080,))):
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    ctx.set_objects(objs)
    W = p.x - 1
    color = torch.empty(W * p.y * 3 + 64, dtype=torch.uint8, device="cuda")
    for rows in rows_list:
        r0 = (p.y - rows) // 2
        for flags in (0, rtc_b200.FLAG_CULL):
            for _ in range(3):
                ctx.trace_band(p, rtc_b200.RGB_PIXEL, r0, r0 + rows, color.data_ptr(), 0, flags)
            torch.cuda.synchronize()
            ms = 0.0
            for _ in range(10):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                ctx.trace_band(p, rtc_b200.RGB_PIXEL, r0, r0 + rows, color.data_ptr(), 0, flags)
                b.record(st)
                torch.cuda.synchronize()
                ms += a.elapsed_time(b)
            print("  %-18s rows %4d %s: %.4f ms  (%.1f us per 100 rows)" % (name, rows, "cull" if flags else "    ", ms / 10, ms / 10 * 1e3 / rows * 100), flush=True)
