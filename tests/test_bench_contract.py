"""CPU: the bench.py contract of the reference arm (`--impl reference`): one JSON line with the agreed keys, nothing of the
product loaded (the camera block comes from the oracle, the timed code is the reference's own RayTrace_* compiled for CPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same config keys as the GPU arm prints
    for k in ("workload", "x", "y", "rays_per_frame", "spheres", "objects", "mode", "parallelism", "gather", "driver", "devices",
              "bands", "camera_orbit_frames", "shadow_rays", "sphere_culling", "l2"):
        assert k in d["config"], k
    assert d["config"]["workload"] == "config3_4k_1024" and d["config"]["rays_per_frame"] == 3840 * 2160


def test_reference_arm_does_not_load_the_product():
    code = ("import sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-seconds', '1'];"
            "import bench; bench.main(); print('LOADED' if 'librtc_b200' in open('/proc/self/maps').read() else 'CLEAN')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout.strip().endswith("CLEAN")
