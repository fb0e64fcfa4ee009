"""Timeline of the multi-GPU frame driver: per device, when each of a few consecutive frames was enqueued, seen finished,
copied.  python scripts/experiments/mgpu_timeline.py [n_gpus] [cull]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import rtc_b200
from rtc_b200 import scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flags = rtc_b200.FLAG_UPDATE_REF_LAUNCH_LIMIT | (rtc_b200.FLAG_CULL if "cull" in sys.argv else 0)
name = "config3_4k_1024"
objs = scenes.config_scene(name); p = scenes.config_camera(name)
m = rtc_b200.MultiGpu(list(range(n)))
m.set_objects(objs)
for _ in range(5):
    m.update(p, 3, 0.0, flags)
K = 40
m.set_objects(objs); m.submit(p, 3, 0.0, flags); m.set_objects(objs); m.submit(p, 3, 0.0, flags)
for _ in range(K):
    m.set_objects(objs); m.submit(p, 3, 0.0, flags); m.collect()
m.collect(); m.collect()
w, mt = m.debug_trace()
last = 5 + K + 2 - 1                      # id of the last frame
frames = [last - 12 + i for i in range(8)]
t0 = mt[frames[0] & 63][0]
print("frame | submit | collect-ret | per device: enq_start enq_end kernels_done(dev_us) copy_issue copy_done")
for f in frames:
    k = f & 63
    print("%3d | %7.0f | %7.0f" % (f, mt[k][0] - t0, mt[k][1] - t0))
    for g in range(n):
        r = w[g][k]
        print("      dev%d  %7.0f %7.0f  %7.0f (%4.0f)  %7.0f %7.0f" % (g, r[0] - t0, r[1] - t0, r[2] - t0, r[5], r[3] - t0, r[4] - t0))
