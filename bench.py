#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native console ray tracer hot path.

Metric (BASELINE.json): Mrays/s (and frames/s) at 3840x2160, 1024 random spheres + plane,
RGB_PIXEL mode, on 1/2/4/8 B200 with a row-band split gathered to GPU 0; ray-kernel fraction of
the FP32 roofline; encoder fraction of HBM bandwidth.

A "step" is one whole frame of the hot path: scene hoist -> ray kernel -> shade+quantise ->
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
    ap.add_argument("--gather", default="host", choices=["host", "nccl", "ipc"],
                    help="N > 1: host = every rank encodes its band and copies its piece of the stream into one shared pinned "
                         "host frame (no data-path collective); ipc / nccl = planes gathered to GPU 0 over NVLink, encoded there")
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
    path = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    try:
        return json.load(open(path))
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
        self.index = index
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
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
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
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(name, mode, seconds, threads=None):
    """Time the reference's own CPU code (oracle/_ref, built from the unmodified sources) on a
    bounded sample of the workload: a window of 16-row block rows of the frame, all host threads.
    Falls back to the restated oracle (kind 'port') where oracle/_ref is absent."""
    from oracle.oracle import Oracle, Reference
    from rtc_b200 import scenes
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
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
    """--impl reference: the reference's own CPU implementation on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import rtc_b200
    mode = rtc_b200.MODE_NAMES.index(args.mode)
    from rtc_b200 import scenes
    p = scenes.config_camera(args.workload)
    per_step = max(1.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(args.workload, mode, per_step)
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
        "config": {"workload": args.workload, "mode": args.mode, "note": "ms_per_step = whole-frame time extrapolated from the sample rate"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                         "cpu_model": cpu_model()},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": value * 1e6 / frame_rays,
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import rtc_b200
    from rtc_b200 import multigpu, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    mode = rtc_b200.MODE_NAMES.index(args.mode)
    name = args.workload
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    # --orbit N: frame i is seen from camera i mod N of an N-frame orbit about the scene centre
    cams = [scenes.config_camera(name, frame=k, n_frames=args.orbit) for k in range(args.orbit)] if args.orbit > 0 else [p]
    frame_no = [0]
    rflags = (rtc_b200.FLAG_SHADOWS if args.shadows else 0) | (rtc_b200.FLAG_CULL if args.cull else 0)

    def next_cam():
        c = cams[frame_no[0] % len(cams)]
        frame_no[0] += 1
        return c
    x, y = p.x, p.y
    W = x - 1
    bpp = rtc_b200.mode_bpp(mode)
    has_glyph = rtc_b200.mode_has_glyph(mode)
    n_spheres = int((objs["type"] == 2).sum())
    frame_rays = W * y

    ctx = rtc_b200.Context(local_rank)            # raises without a GPU: no CPU fallback
    stream = torch.cuda.Stream()                  # one explicit stream for torch, NCCL and the rtc kernels
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_objects(objs)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    cap = rtc_b200.encode_capacity(x, y, mode)
    launches_per_step = 0
    renderer = None
    deficit = 0.0
    if world == 1:
        def step():
            ctx.render(next_cam(), mode, rflags)
    else:
        # Row bands: every rank traces + shades its band straight into (ipc) or followed by NCCL send/recv into (nccl)
        # rank 0's frame planes; rank 0 encodes the assembled frame.  Rank 0 also pays for the encoder, so it gets a
        # smaller band: the deficit (in rows) is measured on rank 0 from one whole frame's stage timings.
        if args.gather == "host":
            # needs a POSIX shared-memory segment that can be page-locked; if any rank cannot have it, every rank
            # falls back to gathering the planes on GPU 0 over NVLink
            try:                                   # collective: raises on every rank or on none
                renderer = multigpu.HostAssembledRenderer(ctx, dist, rank, world, x, y, mode)
            except RuntimeError as e:
                sys.stderr.write("rank %d: %s -- falling back to --gather ipc\n" % (rank, e))
                renderer = None
                args.gather = "ipc"
        if renderer is None:
            hdr = [0.0]
            if rank == 0:
                for _ in range(3):
                    ctx.render(p, mode)
                torch.cuda.synchronize()
                t = ctx.timings()
                hdr = [t["encode_ms"] / max(1e-9, (t["trace_ms"] + t["shade_ms"]) / y)]
            dist.broadcast_object_list(hdr, src=0)
            deficit = float(hdr[0])
            renderer = multigpu.BandRenderer(ctx, dist, rank, world, x, y, mode, gather=args.gather, deficit_rows=deficit)

        def step():
            return renderer.step(next_cam(), rflags)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-timed throughput ---------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = {"prep_ms": 0.0, "trace_ms": 0.0, "shade_ms": 0.0, "encode_ms": 0.0}
    sync_all()
    for i in range(args.steps):
        flush.zero_()                              # evict L2 between timed iterations (untimed)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        if world == 1:
            torch.cuda.synchronize()
            t = ctx.timings()
            for k in stage:
                stage[k] += t[k]
            launches_per_step = t["launches"]
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    my_ms = sum(a.elapsed_time(b) for a, b in ev)
    if dist is not None:
        tt = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    else:
        total_ms = my_ms
    ms_per_step = total_ms / args.steps
    value = frame_rays / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the C-ABI with host buffers (N == 1 path; N > 1: rank 0 reads the stream back)
    e2e = None
    if world == 1:
        # Pipelined public API (rtc_submit / rtc_collect, the asynchronous form of RayTracingManager::Update): every
        # step uploads the scene + camera block from (pinned-staged) host memory and brings that step's stream back
        # to pinned host memory; the D2H of frame k overlaps the kernels of frame k+1.
        upd_flags = rtc_b200.FLAG_UPDATE_REF_LAUNCH_LIMIT | rflags
        for _ in range(2):
            ctx.set_objects(objs)
            s = ctx.update(p, mode, dt=0.0, flags=upd_flags)
        torch.cuda.synchronize()
        ctx.set_objects(objs)
        ctx.submit(p, mode, 0.0, upd_flags)
        t0 = time.perf_counter()
        nbytes = 0
        for _ in range(args.steps):
            ctx.set_objects(objs)                  # scene + camera block from host memory every frame
            ctx.submit(next_cam(), mode, 0.0, upd_flags)    # frame k+1
            s = ctx.collect()                      # frame k: stream in pinned host memory
            nbytes = len(s)
        t1 = time.perf_counter()
        ctx.collect()
        dev_ms_in_pipeline = ctx.timings()["total_ms"]
        e2e_ms = (t1 - t0) * 1e3 / args.steps
        # the synchronous form (one rtc_update per frame, as the reference's Update), for comparison
        t2 = time.perf_counter()
        for _ in range(args.steps):
            ctx.set_objects(objs)
            s = ctx.update(p, mode, dt=0.0, flags=upd_flags)
        t3 = time.perf_counter()
        sync_ms = (t3 - t2) * 1e3 / args.steps
        e2e = {"value": frame_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(objs.nbytes + 96), "d2h_bytes_per_step": int(nbytes + 8),
               "api": "rtc_scene_set_objects + rtc_submit / rtc_collect (pipelined RayTracingManager::Update), stream returned in pinned host memory",
               "device_ms_of_last_pipelined_frame": dev_ms_in_pipeline,
               "synchronous_rtc_update": {"value": frame_rays / (sync_ms * 1e-3) / 1e6, "ms_per_step": sync_ms}}
    else:
        # Pipelined like rtc_submit / rtc_collect: frame k+1 is enqueued on every rank before rank 0 waits for frame k's
        # stream length and copies the stream to pinned host memory on a separate copy stream.
        host = ([torch.empty(cap, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
                if (rank == 0 and args.gather != "host") else None)
        copy_stream = torch.cuda.Stream()
        done = [torch.cuda.Event(), torch.cuda.Event()]
        h_total = torch.zeros(2, dtype=torch.int64, pin_memory=True) if rank == 0 else None

        def submit():
            ctx.set_objects(objs)
            sl = renderer.step(next_cam(), rflags)
            if rank == 0:
                h_total[sl:sl + 1].copy_(renderer.total[sl:sl + 1], non_blocking=True)
                done[sl].record(stream)
            return sl

        def collect(sl):
            if rank != 0:
                return 0
            done[sl].synchronize()
            n = int(h_total[sl])
            with torch.cuda.stream(copy_stream):
                host[sl][:n].copy_(renderer.out[sl][:n], non_blocking=True)
            copy_stream.synchronize()
            return n

        if args.gather == "host":
            def submit():                                  # noqa: F811
                ctx.set_objects(objs)
                return renderer.submit(next_cam(), rflags)

            def collect(sl):                               # noqa: F811
                return renderer.collect()[1]
        sync_all()
        if args.gather == "host":
            # three frames in flight: the copy of frame k+1 is issued while frame k is returned
            submit(); submit()
            t0 = time.perf_counter()
            nbytes = 0
            for _ in range(args.steps):
                submit()
                nbytes = collect(None)
            t1 = time.perf_counter()
            collect(None); collect(None)
        else:
            prev = submit()
            t0 = time.perf_counter()
            nbytes = 0
            for _ in range(args.steps):
                cur = submit()
                nbytes = collect(prev)
                prev = cur
            t1 = time.perf_counter()
            collect(prev)
        sync_all()
        tt = torch.tensor([(t1 - t0) * 1e3 / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
        host_breakdown = None
        if args.gather == "host":
            kk = max(1, renderer.k_col)
            host_breakdown = {"wait_gpu_ms": renderer.t_wait_gpu * 1e3 / kk, "wait_lengths_ms": renderer.t_wait_len * 1e3 / kk,
                              "wait_copy_ms": renderer.t_copy * 1e3 / kk, "wait_ranks_ms": renderer.t_wait_done * 1e3 / kk,
                              "submit_ms": renderer.t_submit * 1e3 / kk}
        e2e = {"rank0_collect_breakdown": host_breakdown,"value": frame_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(objs.nbytes + 96) * world, "d2h_bytes_per_step": int(nbytes + 8),
               "api": ("per rank rtc_scene_set_objects + rtc_trace_band + rtc_encode_band, every rank copies its piece of the stream into one shared pinned host frame (pipelined three deep)"
                       if args.gather == "host" else
                       "per rank rtc_scene_set_objects + rtc_trace_band, bands gathered to GPU 0, rtc_encode, stream copied to pinned host memory (pipelined two deep)")}

    if rank != 0:
        if renderer is not None:
            renderer.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    sm_count = ctx.device_info()["sm_count"]
    fp32_peak = sm_count * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12          # TFLOP/s at the max SM clock
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "x": x, "y": y, "rays_per_frame": frame_rays, "spheres": n_spheres,
                   "objects": int(len(objs)), "mode": args.mode, "parallelism": "rowband%d" % world,
                   "gather": args.gather if world > 1 else None,
                   "bands": renderer.bands if renderer is not None else [[0, y]], "camera_orbit_frames": args.orbit, "shadow_rays": bool(args.shadows), "sphere_culling": bool(args.cull),
                   "l2": "flushed between timed steps (256 MiB memset, untimed)"},
        "frames_per_s": 1e3 / ms_per_step,
        "clocks": clocks, "e2e": e2e,
    }
    if world == 1:
        trace_ms = stage["trace_ms"] / args.steps
        enc_ms = stage["encode_ms"] / args.steps
        n_passes = 2 if args.shadows else 1                    # the shadow pass runs the same packed test over every tile with a shaded pixel
        tests_executed = ctx.timings()["sphere_tests"]         # of the last frame (tile-granular: >= rays x spheres per pass)
        # FLOPs of the tests actually executed: rays x spheres per pass without culling, fewer with --cull
        achieved = 7.0 * (tests_executed if args.cull else frame_rays * n_spheres * n_passes) / (trace_ms * 1e-3) / 1e12
        try:
            measured_ffma = max(ctx.fp32_peak(0, 3000)[0] for _ in range(2))
            measured_ffma2 = max(ctx.fp32_peak(1, 3000)[0] for _ in range(2))
        except Exception:
            measured_ffma = measured_ffma2 = None
        _, n_stream = ctx.frame_ansi_device()
        line["roofline"] = {"bound": "fp32", "kernel": "trace_kernel", "achieved": achieved, "peak": fp32_peak,
                            "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                            "traffic": (ncu_traffic().get("trace_kernel", {}).get("traffic") if name == "config3_4k_1024" else None),
                            "traffic_note": "DRAM bytes per launch from profiles/r01_ncu_traffic.json (ncu --set full): the kernel is FP32-pipe bound, 0.1 % of HBM bandwidth",
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
                                        "traffic": sum(ncu_traffic().get(k, {}).get("traffic", 0.0) for k in ("count_kernel<3>", "emit_kernel<3, 0>")) or None,
                                        "algorithmic_bytes_per_launch": st_in + st_out, "kernel_ms": st_ms,
                                        "peak_source": pk["source"] + " (MEASURED_PEAKS.json hbm_gbs)"}
        except Exception as e:
            line["roofline_encoder"] = {"error": repr(e)}
        line["encoder_on_rendered_frame"] = {"workload": name, "algorithmic_bytes": enc_bytes, "kernel_ms": enc_ms,
                                             "achieved_gbs": enc_bytes / (enc_ms * 1e-3) / 1e9,
                                             "cells_per_ns": frame_rays / (enc_ms * 1e6)}
        line["stages_ms"] = {k: v / args.steps for k, v in stage.items()}
        if not args.cull:
            # The same frames with per-tile sphere culling (RTC_FLAG_CULL): identical output (tests/test_gpu_parity.py::
            # test_culling_is_invisible), fewer ray-sphere tests executed -- reported beside the brute-force headline
            # because it changes the FLOP accounting the roofline above is defined on.
            try:
                cflags = rflags | rtc_b200.FLAG_CULL
                for _ in range(3):
                    ctx.render(next_cam(), mode, cflags)
                torch.cuda.synchronize()
                k = max(10, min(args.steps, 50))
                cms, cst = 0.0, {"trace_ms": 0.0, "shade_ms": 0.0, "encode_ms": 0.0}
                for _ in range(k):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    ctx.render(next_cam(), mode, cflags)
                    b.record(stream)
                    torch.cuda.synchronize()
                    cms += a.elapsed_time(b)
                    t = ctx.timings()
                    for kk in cst:
                        cst[kk] += t[kk]
                cms /= k
                line["with_culling"] = {"value": frame_rays / (cms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": cms,
                                        "frames_per_s": 1e3 / cms, "stages_ms": {kk: v / k for kk, v in cst.items()},
                                        "sphere_tests_executed": t["sphere_tests"],
                                        "fraction_of_brute_force_tests": t["sphere_tests"] / max(1, frame_rays * n_spheres * n_passes),
                                        "output": "bit-identical to the brute-force frame"}
            except Exception as e:
                line["with_culling"] = {"error": repr(e)}
        line["gpu_launches"] = int(launches_per_step * args.steps)
        if not args.no_cpu_baseline:
            line["ref_cuda_sm100"] = ref_cuda_sample(name, mode)
            try:
                cb = cpu_reference_sample(name, mode, args.cpu_seconds)
                line["cpu_baseline"] = {"value": cb["mrays_s"], "unit": "Mrays/s", "cores": cb["cores"], "kind": cb["kind"],
                                        "sample": cb["sample"], "cpu_model": cpu_model()}
            except Exception as e:  # the checker being absent must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    else:
        # hoist + trace + shade per rank, + encode on rank 0
        # hoist + trace + shade per rank; count + emit per rank (host) or on rank 0 only (ipc / nccl)
        line["gpu_launches"] = int(((5 * world) if args.gather == "host" else (3 * world + 2)) * args.steps)
    print(json.dumps(line))
    if renderer is not None:
        renderer.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
