#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native console ray tracer hot path.

Metric (BASELINE.json): Mrays/s (and frames/s) at 3840x2160, 1024 random spheres + plane,
RGB_PIXEL mode, on 1/2/4/8 B200 with a row-band split gathered to GPU 0; ray-kernel fraction of
the FP32 roofline; encoder fraction of HBM bandwidth.

A "step" is one whole frame of the hot path: ray kernel (scene hoist, nearest hit, shade+quantise) ->
(N>1: gather of the RGB8 bands to GPU 0) -> ANSI encode.  `value` is timed on the device with
CUDA events (inputs resident in HBM); `e2e` is the same frame through the C-ABI with HOST
buffers: scene + camera block uploaded, minimised stream copied back to pinned host memory.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference's own CPU code on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3_4k_1024")
    ap.add_argument("--mode", default="RGB_PIXEL")
    ap.add_argument("--gather", default="host", choices=["host", "p2p"],
                    help="N > 1: host = every device encodes its band and copies its piece of the stream into one pinned host "
                         "frame (no data-path collective); p2p = bands stored straight into GPU 0's planes over NVLink, encoded there. "
                         "The other mode is measured too and reported as a sub-record.")
    ap.add_argument("--headline-only", action="store_true", help="skip the sub-records (culling, other gather, config 4, config 2 + shadows, baselines)")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--orbit", type=int, default=0, help="camera orbit of this many frames (config 4: 120); 0 = fixed camera")
    ap.add_argument("--cull", action="store_true", help="per-tile sphere culling on (identical results, fewer tests executed)")
    ap.add_argument("--shadows", action="store_true", help="shadow-ray extension on (second, light-origin trace pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def ncu_traffic():
    """DRAM bytes per launch of our kernels from the committed ncu --set full captures (profiles/)."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))     # the latest round's captures
    try:
        return json.load(open(paths[-1]))
    except Exception:
        return {}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d.get("hbm_gbs", 6650.0)), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index if isinstance(index, (list, tuple)) else [index]      # one GPU or several (median / max over all)
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                txt = subprocess.run(["nvidia-smi", "-i", ",".join(str(i) for i in self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for ln in txt.splitlines():
                    out = ln.split(",")
                    self.samples.append(float(out[0]))
                    self.max_mhz = float(out[1])
                    for nm, v in zip(names, out[2:6]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sm_mhz_min": float(np.min(self.samples)) if self.samples else None}


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(name, mode, seconds, threads=None, camera_fn=None):
    """Time the reference's own CPU code (oracle/_ref, built from the unmodified sources) on a
    bounded sample of the workload: a window of 16-row block rows of the frame, all host threads.
    Falls back to the restated oracle (kind 'port') where oracle/_ref is absent."""
    from oracle.oracle import Oracle, Reference
    from rtc_b200 import scenes
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name, camera_fn=camera_fn)
    threads = threads or os.cpu_count() or 1
    n_obj = len(objs)
    gy = (p.y + 15) // 16
    mid = gy // 2
    if Reference.available():
        R = Reference()
        kind = "reference"

        def run(b0, b1):
            return R.trace_blockrows(objs, p, mode, b0, b1, threads)
    else:
        O = Oracle()
        kind = "port"

        def run(b0, b1):
            r0, r1 = b0 * 16, min(b1 * 16, p.y)
            return O.time_trace(objs, p, mode, r0, r1, threads), (r1 - r0) * (p.x - 1)
    # calibrate on a thin window through the middle of the frame, then size the sample
    nb = max(1, min(gy, (threads + 239) // 240))
    secs, rays = run(mid, mid + nb)
    rate = rays / max(secs, 1e-9)
    want_rays = rate * seconds
    rows_per_block = 16 * (p.x - 1)
    nblocks = int(max(nb, min(gy, round(want_rays / rows_per_block))))
    b0 = max(0, mid - nblocks // 2)
    b1 = min(gy, b0 + nblocks)
    secs, rays = run(b0, b1)
    return dict(kind=kind, cores=threads, secs=secs, rays=rays, mrays_s=rays / secs / 1e6,
                sample="%d of %d rows (block rows %d..%d through the frame centre) of %s, %d objects, %s kernel only"
                       % (min(b1 * 16, p.y) - b0 * 16, p.y, b0, b1, name, n_obj, "reference RayTrace_*" if kind == "reference" else "oracle"))


def encoder_stress(ctx, stream, flush, iters=20):
    """BASELINE config 5: 7680x4320 i.i.d. random RGB -> ANSI stream (almost every cell emits its 20-byte
    escape: the worst case, 99.5 MB in + 663.6 MB out).  Timed with CUDA events on the launching stream, L2
    flushed between iterations.  Returns (ms per encode, stream bytes)."""
    import torch
    import rtc_b200
    x, y = 7681, 4320
    W = x - 1
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    rgb = torch.randint(0, 256, (W * y * 3,), dtype=torch.uint8, device="cuda", generator=g)
    cap = rtc_b200.encode_capacity(x, y, rtc_b200.RGB_PIXEL)
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.encode(rgb.data_ptr(), 0, x, y, rtc_b200.RGB_PIXEL, out.data_ptr(), cap, total.data_ptr())
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        ctx.encode(rgb.data_ptr(), 0, x, y, rtc_b200.RGB_PIXEL, out.data_ptr(), cap, total.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    n = int(total.item())
    del rgb, out
    return ms / iters, n, 3 * W * y


def ref_cuda_sample(name, mode, frames=3):
    """Second reported baseline: the reference's OWN CUDA kernels rebuilt for sm_100 (oracle/_ref/
    ref_cuda_sm100, compiled from the unmodified sources with the vcxproj's flags)."""
    import struct
    import tempfile
    from rtc_b200 import scenes
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cuda_sm100")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_cuda_sm100 not built (needs /root/reference at build time)"}
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(struct.pack("<II", len(objs), mode))
        f.write(bytes(p))
        f.write(objs.tobytes())
        path = f.name
    try:
        r = subprocess.run([exe, path, str(frames)], capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            return {"unavailable": "ref_cuda_sm100 exited %d: %s" % (r.returncode, (r.stderr or r.stdout)[-200:])}
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"unavailable": repr(e)}
    finally:
        os.unlink(path)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores.  Nothing of the product is loaded
    here: the camera block comes from the oracle's restatement of Camera3D (orc_camera_params), not from librtc_b200."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rtc_b200 import _types, scenes                        # numpy / ctypes PODs only; does not dlopen the library
    from oracle.oracle import Oracle
    orc = Oracle()
    mode = _types.MODE_NAMES.index(args.mode)
    p = scenes.config_camera(args.workload, camera_fn=orc.camera_params)
    objs = scenes.config_scene(args.workload)
    per_step = max(1.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(args.workload, mode, per_step, camera_fn=orc.camera_params)
        if i >= args.warmup:
            vals.append(last)
    rays = sum(v["rays"] for v in vals)
    secs = sum(v["secs"] for v in vals)
    value = rays / secs / 1e6
    frame_rays = (p.x - 1) * p.y
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * frame_rays / (value * 1e6),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "x": p.x, "y": p.y, "rays_per_frame": frame_rays,
                   "spheres": int((objs["type"] == 2).sum()), "objects": int(len(objs)), "mode": args.mode,
                   "parallelism": "host threads", "gather": None, "driver": "reference RayTrace_* kernels compiled for CPU (oracle/_ref)",
                   "devices": [], "bands": [[0, p.y]], "camera_orbit_frames": args.orbit, "shadow_rays": False, "sphere_culling": False,
                   "l2": "n/a", "note": "ms_per_step = whole-frame time extrapolated from the sample rate"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                         "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": value * 1e6 / frame_rays,
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def sha16(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def measure_ctx(ctx, torch, stream, flush, objs, cams, mode, flags, steps, warmup, e2e_steps=None, sample_e2e_clocks=False):
    """One GPU through rtc_ctx.  Device-timed frames (CUDA events on the launching stream, L2 flushed between steps)
    and the end-to-end loop through rtc_scene_set_objects + rtc_submit / rtc_collect with host buffers."""
    import rtc_b200
    k = [0]

    def cam():
        c = cams[k[0] % len(cams)]
        k[0] += 1
        return c
    ctx.set_objects(objs)
    for _ in range(max(3, warmup)):
        ctx.render(cam(), mode, flags)
    torch.cuda.synchronize()
    stage = {"prep_ms": 0.0, "trace_ms": 0.0, "shade_ms": 0.0, "encode_ms": 0.0}
    total_ms, launches, t = 0.0, 0, None
    for _ in range(steps):
        flush.zero_()                              # evict L2 between timed iterations (untimed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        ctx.render(cam(), mode, flags)
        b.record(stream)
        torch.cuda.synchronize()
        total_ms += a.elapsed_time(b)
        t = ctx.timings()
        for kk in stage:
            stage[kk] += t[kk]
        launches = t["launches"]
    res = {"ms_per_step": total_ms / steps, "stages_ms": {kk: v / steps for kk, v in stage.items()}, "launches_per_step": launches,
           "sphere_tests": t["sphere_tests"]}
    # ---- end to end: every step uploads the scene + camera block from host memory and brings that step's stream
    # back to pinned host memory; the D2H of frame k overlaps the kernels of frame k+1 (rtc_submit / rtc_collect)
    n = e2e_steps or steps
    upd = rtc_b200.FLAG_UPDATE_REF_LAUNCH_LIMIT | flags
    for _ in range(4):                             # untimed: per-frame uploads touch every slot of the scene ring
        ctx.set_objects(objs)
        ctx.update(cam(), mode, dt=0.0, flags=upd)
    torch.cuda.synchronize()
    ctx.set_objects(objs)
    ctx.submit(cam(), mode, 0.0, upd)
    sampler = ClockSampler(ctx.device) if sample_e2e_clocks else None
    if sampler:                                    # (the pipelined loop keeps the GPU busy back to back: sustained, not boost, clocks)
        sampler.start()
        n = max(n, int(1500.0 / max(res["ms_per_step"], 1e-3)))      # >= 1.5 s so that nvidia-smi gets a few samples in
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(n):
        ctx.set_objects(objs)
        ctx.submit(cam(), mode, 0.0, upd)          # frame k+1
        nbytes = len(ctx.collect())                # frame k: stream in pinned host memory
    t1 = time.perf_counter()
    ctx.collect()
    if sampler:
        res["e2e_clocks"] = sampler.stop()
    res["e2e_ms"] = (t1 - t0) * 1e3 / n
    res["d2h_bytes"] = int(nbytes + 8)
    res["h2d_bytes"] = int(objs.nbytes + 96)
    t2 = time.perf_counter()
    for _ in range(n):
        ctx.set_objects(objs)
        ctx.update(cam(), mode, dt=0.0, flags=upd)
    res["sync_ms"] = (time.perf_counter() - t2) * 1e3 / n
    return res


def measure_mgpu(m, objs, cams, mode, flags, steps, warmup, e2e_steps=None, sample_e2e_clocks=False):
    """N GPUs through rtc_mgpu (one process, row bands).  `value`: per step every device's L2 is flushed (untimed), the
    frame is submitted and collected; the step's time is what the CUDA events on each device's stream say (first kernel
    to end of that device's last kernel; in the p2p gather device 0's last kernel is the encode of the assembled frame);
    summed per device, maximum over devices.  `e2e`: wall clock of the pipelined loop, scene + camera uploaded from host
    memory every frame, the assembled stream back in one pinned host buffer every frame."""
    import rtc_b200
    k = [0]

    def cam():
        c = cams[k[0] % len(cams)]
        k[0] += 1
        return c
    upd = rtc_b200.FLAG_UPDATE_REF_LAUNCH_LIMIT | flags
    m.set_objects(objs)
    for _ in range(max(3, warmup)):
        m.update(cam(), mode, 0.0, upd)
    per_dev = np.zeros(m.n)
    enc = 0.0
    info = None
    for _ in range(steps):
        m.flush_l2()
        m.update(cam(), mode, 0.0, upd)
        info = m.last_frame()
        per_dev += np.array(info["device_ms"])
        enc += info["encode_ms"]
    res = {"ms_per_step": float(per_dev.max()) / steps, "per_device_ms": [float(v) / steps for v in per_dev], "bands": info["bands"],
           "encode_ms_on_gpu0": enc / steps}
    n = e2e_steps or steps
    for _ in range(2):                             # untimed: the pipelined path with per-frame uploads, every ring slot touched
        m.set_objects(objs)
        m.submit(cam(), mode, 0.0, upd)
        m.set_objects(objs)
        m.submit(cam(), mode, 0.0, upd)
        m.collect(); m.collect()
    m.host_stats()                                 # reset the driver's host-side accounting
    sampler = ClockSampler(sorted(set(m.devices))) if sample_e2e_clocks else None
    if sampler:                                    # (the pipelined loop keeps every GPU busy back to back: sustained, not boost, clocks)
        sampler.start()
        n = max(n, int(1500.0 / max(res["ms_per_step"], 1e-3)))      # >= 1.5 s so that nvidia-smi gets a few samples in
    m.set_objects(objs)
    m.submit(cam(), mode, 0.0, upd)
    m.set_objects(objs)
    m.submit(cam(), mode, 0.0, upd)
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(n):
        m.set_objects(objs)                        # scene + camera block from host memory every frame
        m.submit(cam(), mode, 0.0, upd)            # frame k+2
        nbytes = len(m.collect())                  # frame k: the assembled stream in pinned host memory
    t1 = time.perf_counter()
    m.collect(); m.collect()
    if sampler:
        res["e2e_clocks"] = sampler.stop()
    res["host_us_per_frame"] = m.host_stats()
    w, _ = m.debug_trace()
    res["device_us_in_pipelined_loop"] = [round(float(np.median(w[g][:, 5])), 1) for g in range(m.n)]
    res["e2e_ms"] = (t1 - t0) * 1e3 / n
    res["d2h_bytes"] = int(nbytes + 8 * m.n)
    res["h2d_bytes"] = int(objs.nbytes + 96) * m.n
    t2 = time.perf_counter()
    for _ in range(max(3, n // 4)):
        m.set_objects(objs)
        m.update(cam(), mode, 0.0, upd)
    res["sync_ms"] = (time.perf_counter() - t2) * 1e3 / max(3, n // 4)
    return res


def parity_streams(ctx1, objs, cams, mode, flags):
    """Single-GPU streams (rtc_update on one context) of a few cameras: the yardstick of the N-GPU parity check."""
    import rtc_b200
    ctx1.set_objects(objs)
    return [sha16(ctx1.update(c, mode, dt=0.0, flags=flags | rtc_b200.FLAG_UPDATE_REF_LAUNCH_LIMIT)) for c in cams]


def parity_mgpu(m, objs, cams, mode, flags):
    import rtc_b200
    m.set_objects(objs)
    return [sha16(m.update(c, mode, 0.0, flags | rtc_b200.FLAG_UPDATE_REF_LAUNCH_LIMIT)) for c in cams]


def run_ours(args):
    import torch
    import rtc_b200
    from rtc_b200 import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    N = args.gpus
    torch.cuda.set_device(local_rank)
    dist, gloo = None, None
    if world > 1:
        # One rank per GPU is the launch contract; the FRAME PATH is one process: rank 0 drives all N devices through
        # the library's multi-GPU driver (rtc_mgpu: a worker thread + stream per device).  NCCL carries the rank check
        # (a barrier over all N ranks on their GPUs) and the timing reduction; the other ranks then wait on a gloo
        # barrier (host-side, so that nothing of theirs spins on the GPUs rank 0 is timing).
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")
        dist.barrier()
        torch.cuda.synchronize()
        if rank != 0:
            dist.barrier(group=gloo)                            # rank 0 has finished the headline measurement
            tt = torch.zeros(2, dtype=torch.float64, device=torch.device("cuda", local_rank))
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.barrier(group=gloo)                            # rank 0 has finished the sub-records too: only now tear the
            dist.destroy_process_group()                        # CUDA contexts on the other GPUs down (it disturbs short timings)
            return

    mode = rtc_b200.MODE_NAMES.index(args.mode)
    name = args.workload
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    cams = [scenes.config_camera(name, frame=k, n_frames=args.orbit) for k in range(args.orbit)] if args.orbit > 0 else [p]
    rflags = (rtc_b200.FLAG_SHADOWS if args.shadows else 0) | (rtc_b200.FLAG_CULL if args.cull else 0)
    x, y = p.x, p.y
    W = x - 1
    bpp = rtc_b200.mode_bpp(mode)
    n_spheres = int((objs["type"] == 2).sum())
    frame_rays = W * y
    n_vis = torch.cuda.device_count()
    devices = list(range(N)) if n_vis >= N else [g % n_vis for g in range(N)]     # fewer GPUs visible: bands share devices (dev boxes)

    ctx = rtc_b200.Context(devices[0])            # raises without a GPU: no CPU fallback
    stream = torch.cuda.Stream()                  # one explicit stream for torch and the rtc kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    sampler = ClockSampler(devices[0])
    sampler.start()
    m = None
    if N == 1:
        r = measure_ctx(ctx, torch, stream, flush, objs, cams, mode, rflags, args.steps, args.warmup, sample_e2e_clocks=True)
    else:
        gather = rtc_b200.GATHER_P2P if args.gather == "p2p" else rtc_b200.GATHER_HOST
        m = rtc_b200.MultiGpu(devices, gather)
        r = measure_mgpu(m, objs, cams, mode, rflags, args.steps, args.warmup, sample_e2e_clocks=True)
    clocks = sampler.stop()
    ms_per_step = r["ms_per_step"]
    if dist is not None:
        dist.barrier(group=gloo)
        torch.cuda.set_device(local_rank)
        tt = torch.tensor([ms_per_step, r["e2e_ms"]], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)                # max over ranks (the other ranks time nothing: 0)
        ms_per_step = float(tt[0].item())
    value = frame_rays / (ms_per_step * 1e-3) / 1e6

    def mr(ms, rays=frame_rays):
        return rays / (ms * 1e-3) / 1e6

    e2e = {"value": mr(r["e2e_ms"]), "unit": "Mrays/s", "ms_per_step": r["e2e_ms"], "h2d_bytes_per_step": r["h2d_bytes"],
           "d2h_bytes_per_step": r["d2h_bytes"],
           "api": ("rtc_scene_set_objects + rtc_submit / rtc_collect (pipelined RayTracingManager::Update), stream returned in pinned host memory"
                   if N == 1 else
                   "rtc_mgpu_scene_set_objects + rtc_mgpu_submit / rtc_mgpu_collect (RayTracingManager::Update across %d GPUs, one process, "
                   "three frames in flight), the assembled stream returned in one pinned host buffer" % N),
           "synchronous_update": {"value": mr(r["sync_ms"]), "ms_per_step": r["sync_ms"]}}
    if "host_us_per_frame" in r:
        e2e["worker_threads_us_per_frame"] = r["host_us_per_frame"]
        e2e["device_us_per_frame_in_this_loop"] = r["device_us_in_pipelined_loop"]
    if "e2e_clocks" in r:
        e2e["clocks"] = r["e2e_clocks"]

    pk = peaks()
    sm_count = ctx.device_info()["sm_count"]
    fp32_peak = sm_count * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12          # TFLOP/s at the max SM clock
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": N, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "x": x, "y": y, "rays_per_frame": frame_rays, "spheres": n_spheres,
                   "objects": int(len(objs)), "mode": args.mode, "parallelism": "rowband%d" % N,
                   "gather": args.gather if N > 1 else None, "driver": "rtc_ctx" if N == 1 else "rtc_mgpu (one process, worker thread per device)",
                   "devices": devices, "bands": r.get("bands", [[0, y]]), "camera_orbit_frames": args.orbit,
                   "shadow_rays": bool(args.shadows), "sphere_culling": bool(args.cull),
                   "l2": "flushed between timed steps (256 MiB memset per device, untimed)"},
        "frames_per_s": 1e3 / ms_per_step,
        "clocks": clocks, "e2e": e2e,
    }
    traffic = ncu_traffic()
    if N == 1:
        trace_ms = r["stages_ms"]["trace_ms"]
        enc_ms = r["stages_ms"]["encode_ms"]
        n_passes = 2 if args.shadows else 1                    # the shadow pass runs the same packed test over every tile with a shaded pixel
        tests_executed = r["sphere_tests"]                     # of the last frame (tile-granular: >= rays x spheres per pass)
        achieved = 7.0 * (tests_executed if args.cull else frame_rays * n_spheres * n_passes) / (trace_ms * 1e-3) / 1e12
        try:
            measured_ffma = max(ctx.fp32_peak(0, 3000)[0] for _ in range(2))
            measured_ffma2 = max(ctx.fp32_peak(1, 3000)[0] for _ in range(2))
        except Exception:
            measured_ffma = measured_ffma2 = None
        ctx.render(p, mode, rflags)
        _, n_stream = ctx.frame_ansi_device()
        tk = traffic.get("trace_kernel", {})
        line["roofline"] = {"bound": "fp32", "kernel": "trace_kernel (ray generation + nearest hit + shade/quantise epilogue, one launch)",
                            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                            "traffic": tk.get("traffic") if name == tk.get("workload") else None,
                            "traffic_source": tk.get("source"),
                            "peak_source": "%d SMs x 128 lanes x 2 FLOP x %.0f MHz (%s sm_max_mhz); not in MEASURED_PEAKS.json, which has HBM and bf16 tensor only"
                                           % (sm_count, pk["sm_max_mhz"], pk["source"]),
                            "algorithmic_flops_per_launch": 7.0 * frame_rays * n_spheres, "launches_in_kernel_ms": n_passes, "kernel_ms": trace_ms,
                            "sphere_tests_executed": tests_executed, "sphere_tests_brute_force": frame_rays * n_spheres * n_passes,
                            "measured_ffma_tflops": measured_ffma, "measured_ffma2_tflops": measured_ffma2,
                            "frac_of_measured_ffma": (achieved / measured_ffma) if measured_ffma else None}
        enc_bytes = bpp * frame_rays + n_stream
        # The encoder's roofline is quoted on BASELINE config 5 (8K worst case: every cell emits 20 bytes); the
        # frame rendered above is mostly background runs (1 byte per cell), i.e. cell-rate- not byte-bound.
        try:
            st_ms, st_out, st_in = encoder_stress(ctx, stream, flush)
            st_gbs = (st_in + st_out) / (st_ms * 1e-3) / 1e9
            line["roofline_encoder"] = {"bound": "hbm", "kernel": "count_kernel + emit_kernel (2 launches)",
                                        "workload": "config5_encode_8k: 7681x4320, i.i.d. random RGB",
                                        "achieved": st_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": st_gbs / pk["hbm_gbs"],
                                        "traffic": sum(traffic.get(k, {}).get("traffic", 0.0) for k in ("count_kernel<3>", "emit_kernel<3, 0>")) or None,
                                        "traffic_source": traffic.get("emit_kernel<3, 0>", {}).get("source"),
                                        "algorithmic_bytes_per_launch": st_in + st_out, "kernel_ms": st_ms,
                                        "peak_source": pk["source"] + " (MEASURED_PEAKS.json hbm_gbs)"}
        except Exception as e:
            line["roofline_encoder"] = {"error": repr(e)}
        line["encoder_on_rendered_frame"] = {"workload": name, "algorithmic_bytes": enc_bytes, "kernel_ms": enc_ms,
                                             "achieved_gbs": enc_bytes / (enc_ms * 1e-3) / 1e9,
                                             "cells_per_ns": frame_rays / (enc_ms * 1e6)}
        line["stages_ms"] = r["stages_ms"]
        line["gpu_launches"] = int(r["launches_per_step"] * args.steps)
        # What the shade + quantise epilogue costs inside the ray kernel: the same frames in SDL mode, which traces and
        # writes nothing (reference RayTracing.cu:755-795) -- same kernel, epilogue off.
        try:
            ctx.set_objects(objs)
            t_sdl = 0.0
            for i in range(13):
                flush.zero_()
                ctx.render(cams[i % len(cams)], rtc_b200.SDL, rflags)
                torch.cuda.synchronize()
                if i >= 3:
                    t_sdl += ctx.timings()["trace_ms"]
            t_sdl /= 10
            line["shade_epilogue"] = {"trace_ms_with_epilogue": trace_ms, "trace_ms_sdl_mode_no_epilogue": t_sdl, "epilogue_ms": trace_ms - t_sdl,
                                      "roofline_frac_of_the_bare_trace": 7.0 * frame_rays * n_spheres * n_passes / (t_sdl * 1e-3) / 1e12 / fp32_peak}
        except Exception as e:
            line["shade_epilogue"] = {"error": repr(e)}
    else:
        line["per_device_ms"] = r["per_device_ms"]
        line["encode_ms_on_gpu0"] = r["encode_ms_on_gpu0"]
        # per device: trace (scene hoist as its prologue, shade as its tile epilogue; with shadow rays: + shadow trace + shade), then
        # count + emit per device (host gather) or once on GPU 0 (p2p)
        tr = 3 if args.shadows else 1
        line["gpu_launches"] = int(((tr + 2) * N if args.gather == "host" else tr * N + 2) * args.steps)

    if not args.headline_only:
        sub_steps = max(8, min(args.steps, 40))
        # ---- the same frames with per-tile sphere culling (RTC_FLAG_CULL): identical output, fewer tests executed --
        if not args.cull:
            try:
                cf = rflags | rtc_b200.FLAG_CULL
                rc = (measure_ctx(ctx, torch, stream, flush, objs, cams, mode, cf, sub_steps, 3) if N == 1 else
                      measure_mgpu(m, objs, cams, mode, cf, sub_steps, 3))
                line["with_culling"] = {"value": mr(rc["ms_per_step"]), "unit": "Mrays/s", "ms_per_step": rc["ms_per_step"],
                                        "e2e": {"value": mr(rc["e2e_ms"]), "ms_per_step": rc["e2e_ms"]},
                                        "stages_ms": rc.get("stages_ms"), "sphere_tests_executed": rc.get("sphere_tests"),
                                        "output": "bit-identical to the brute-force frame (tests/test_gpu_parity.py::test_culling_is_invisible)"}
            except Exception as e:
                line["with_culling"] = {"error": repr(e)}
            # ---- and with the packet filter (RTC_FLAG_PACKET: the ray-sphere filter once per 8-ray packet), alone and on top
            #      of the culling -- what the facade's RayTracingManager::Update runs by default; identical output ----------
            for key, extra in (("with_packet_filter", rtc_b200.FLAG_PACKET), ("with_culling_and_packet_filter", rtc_b200.FLAG_PACKET | rtc_b200.FLAG_CULL)):
                try:
                    pf = rflags | extra
                    rc = (measure_ctx(ctx, torch, stream, flush, objs, cams, mode, pf, sub_steps, 3) if N == 1 else
                          measure_mgpu(m, objs, cams, mode, pf, sub_steps, 3))
                    same = None
                    if N == 1:
                        same = parity_streams(ctx, objs, cams[:2], mode, pf) == parity_streams(ctx, objs, cams[:2], mode, rflags)
                    line[key] = {"value": mr(rc["ms_per_step"]), "unit": "Mrays/s", "ms_per_step": rc["ms_per_step"],
                                 "e2e": {"value": mr(rc["e2e_ms"]), "ms_per_step": rc["e2e_ms"]}, "stages_ms": rc.get("stages_ms"),
                                 "stream_equals_per_ray_filter": same,
                                 "output": "bit-identical to the per-ray filter (tests/test_gpu_parity.py::test_packet_filter_is_invisible)"}
                except Exception as e:
                    line[key] = {"error": repr(e)}
        # ---- N > 1: the other gather, and the N-GPU bytes against a single-GPU frame (outside every timed region) ----
        parity = {}
        if N > 1:
            other = "p2p" if args.gather == "host" else "host"
            try:
                want = parity_streams(ctx, objs, cams[:3], mode, rflags)
                parity[args.gather] = parity_mgpu(m, objs, cams[:3], mode, rflags) == want
                m2 = rtc_b200.MultiGpu(devices, rtc_b200.GATHER_P2P if other == "p2p" else rtc_b200.GATHER_HOST)
                ro = measure_mgpu(m2, objs, cams, mode, rflags, sub_steps, 3)
                parity[other] = parity_mgpu(m2, objs, cams[:3], mode, rflags) == want
                m2.close()
                line["gather_" + other] = {"value": mr(ro["ms_per_step"]), "unit": "Mrays/s", "ms_per_step": ro["ms_per_step"],
                                           "e2e": {"value": mr(ro["e2e_ms"]), "ms_per_step": ro["e2e_ms"]}, "bands": ro["bands"],
                                           "per_device_ms": ro["per_device_ms"], "encode_ms_on_gpu0": ro["encode_ms_on_gpu0"]}
            except Exception as e:
                line["gather_" + other] = {"error": repr(e)}
                parity[other] = False
        # ---- config 4 (7681x4320, 4096 spheres, camera orbit: the configuration the >= 7x scaling target is stated on)
        if name != "config4_8k_4096" and not args.no_config4:
            try:
                o4 = scenes.config_scene("config4_8k_4096")
                n_orbit = 24
                c4 = [scenes.config_camera("config4_8k_4096", frame=k * 5, n_frames=120) for k in range(n_orbit)]   # every 5th of the 120 orbit frames
                rays4 = (c4[0].x - 1) * c4[0].y
                if N == 1:
                    r4 = measure_ctx(ctx, torch, stream, flush, o4, c4, rtc_b200.RGB_PIXEL, 0, n_orbit, 3, e2e_steps=n_orbit)
                else:
                    r4 = measure_mgpu(m, o4, c4, rtc_b200.RGB_PIXEL, 0, n_orbit, 3, e2e_steps=n_orbit)
                    want4 = parity_streams(ctx, o4, c4[:2], rtc_b200.RGB_PIXEL, 0)
                    parity["config4_" + args.gather] = parity_mgpu(m, o4, c4[:2], rtc_b200.RGB_PIXEL, 0) == want4
                line["config4_orbit"] = {"workload": "config4_8k_4096: 7681x4320, 4096 spheres, 24 frames of the 120-frame orbit (every 5th)",
                                         "value": mr(r4["ms_per_step"], rays4), "unit": "Mrays/s", "ms_per_step": r4["ms_per_step"],
                                         "frames_per_s": 1e3 / r4["ms_per_step"], "steps": n_orbit,
                                         "e2e": {"value": mr(r4["e2e_ms"], rays4), "ms_per_step": r4["e2e_ms"], "frames_per_s": 1e3 / r4["e2e_ms"],
                                                 "d2h_bytes_per_step": r4["d2h_bytes"], "h2d_bytes_per_step": r4["h2d_bytes"]},
                                         "bands": r4.get("bands", [[0, c4[0].y]]), "per_device_ms": r4.get("per_device_ms"),
                                         "stages_ms": r4.get("stages_ms"),
                                         "roofline_frac": (7.0 * rays4 * 4096 / (r4["stages_ms"]["trace_ms"] * 1e-3) / 1e12 / fp32_peak) if N == 1 else None}
            except Exception as e:
                line["config4_orbit"] = {"error": repr(e)}
        if N > 1:
            line["parity_check"] = {"n_gpu_equals_1_gpu": bool(parity) and all(parity.values()), "cases": parity,
                                    "how": "sha256 of the assembled N-GPU stream (rtc_mgpu_update) == sha256 of rtc_update on one context, same cameras; outside the timed regions"}
        # ---- config 2 with shadow rays (1921x1080, 64 spheres + plane, primary + shadow rays) on one GPU --------------
        if N == 1 and not args.shadows:
            try:
                o2 = scenes.config_scene("config2_1080p_64")
                c2 = [scenes.config_camera("config2_1080p_64")]
                rays2 = (c2[0].x - 1) * c2[0].y
                r2 = measure_ctx(ctx, torch, stream, flush, o2, c2, rtc_b200.RGB_PIXEL, rtc_b200.FLAG_SHADOWS, sub_steps, 3)
                line["config2_shadows"] = {"workload": "config2_1080p_64: 1921x1080, 64 spheres + plane, primary + shadow rays (RTC_FLAG_SHADOWS: an extension, the reference casts none)",
                                           "value": mr(r2["ms_per_step"], rays2), "unit": "Mrays/s (primary rays)", "ms_per_step": r2["ms_per_step"],
                                           "e2e": {"value": mr(r2["e2e_ms"], rays2), "ms_per_step": r2["e2e_ms"]}, "stages_ms": r2["stages_ms"]}
            except Exception as e:
                line["config2_shadows"] = {"error": repr(e)}
        if N == 1 and not args.no_cpu_baseline:
            line["ref_cuda_sm100"] = ref_cuda_sample(name, mode)
            try:
                cb = cpu_reference_sample(name, mode, args.cpu_seconds)
                line["cpu_baseline"] = {"value": cb["mrays_s"], "unit": "Mrays/s", "cores": cb["cores"], "kind": cb["kind"],
                                        "sample": cb["sample"], "cpu_model": cpu_model()}
            except Exception as e:  # the checker being absent must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line))
    sys.stdout.flush()
    if m is not None:
        m.close()
    ctx.close()
    ok = line.get("parity_check", {}).get("n_gpu_equals_1_gpu", True)
    if dist is not None:
        dist.barrier(group=gloo)
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("N-GPU stream differs from the single-GPU stream: %r" % (line["parity_check"],))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
