"""Latency of a SYNCHRONOUS frame (what the reference's RayTracingManager::Update is: call, wait, bytes in host memory) when one
GPU traces the frame as k row bands through the multi-GPU driver (device_ids = [0]*k): band j's stream crosses PCIe while
band j+1 is still tracing.  Wall clock around rtc_update / rtc_mgpu_update, scene re-uploaded every frame.
python scripts/experiments/probe_sync_bands.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402,F401
import rtc_b200  # noqa: E402
from rtc_b200 import scenes  # noqa: E402


def timed(fn, n=40, warm=6):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[len(ts) // 2], 1e3 * ts[0]


for name in ("config3_4k_1024", "config2_1080p_64", "config4_8k_4096"):
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    for flags in (0, rtc_b200.FLAG_CULL):
        ctx = rtc_b200.Context(0)

        def one():
            ctx.set_objects(objs)
            return ctx.update(p, rtc_b200.RGB_PIXEL, 0.0, flags)
        want = bytes(one())
        med, best = timed(one, 40 if name != "config4_8k_4096" else 8, 6 if name != "config4_8k_4096" else 2)
        print("%-18s %s rtc_update (1 band)        : median %.3f ms, best %.3f" % (name, "cull" if flags else "    ", med, best), flush=True)
        del ctx
        for k in (2, 3, 4, 6, 8):
            with rtc_b200.MultiGpu([0] * k, rtc_b200.GATHER_HOST) as m:
                def onem():
                    m.set_objects(objs)
                    return m.update(p, rtc_b200.RGB_PIXEL, 0.0, flags)
                same = bytes(onem()) == want
                med, best = timed(onem, 40 if name != "config4_8k_4096" else 8, 6 if name != "config4_8k_4096" else 2)
                print("%-18s %s rtc_mgpu_update [0]*%d       : median %.3f ms, best %.3f  bytes equal: %s" % (name, "cull" if flags else "    ", k, med, best, same), flush=True)
