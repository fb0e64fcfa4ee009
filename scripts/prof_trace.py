"""Small driver for ncu: a few frames of one config (default config 3)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtc_b200
from rtc_b200 import scenes

name = sys.argv[1] if len(sys.argv) > 1 else "config3_4k_1024"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = sys.argv[3].split("+") if len(sys.argv) > 3 else []          # cull, packet, cull+packet
flags = (rtc_b200.FLAG_CULL if "cull" in opts else 0) | (rtc_b200.FLAG_PACKET if "packet" in opts else 0)
ctx = rtc_b200.Context(0)
ctx.set_objects(scenes.config_scene(name))
p = scenes.config_camera(name)
for _ in range(n):
    ctx.render(p, rtc_b200.RGB_PIXEL, flags)
    ctx.frame_ansi_device()
print(ctx.timings())
