"""rtc_b200 -- Python plumbing over the C-ABI of include/rtc.h (librtc_b200.so).

The product is the CUDA library; this module only binds it with ctypes so tests, bench.py and
multi-GPU drivers (torch.distributed) can call it.  There is no CPU fallback: creating a
Context without a CUDA device raises RtcError.
"""
import ctypes
import os

import numpy as np

from . import _types, scenes  # noqa: F401
from ._types import (BIT_ASCII, BIT_PIXEL, FLAG_CULL, FLAG_KEEP_HITS, FLAG_NORMALS_SATURATE, FLAG_PACKET, FLAG_SHADOWS, FLAG_UPDATE_REF_LAUNCH_LIMIT, MODE_NAMES,  # noqa: F401
                     OBJECT_DTYPE, RGB_ASCII, RGB_NORMALS, RGB_PIXEL, SDL, RtcParams, RtcTimings,
                     mode_bpp, mode_cell, mode_has_glyph, obj_ptr)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTC_B200_LIB") or os.path.join(_HERE, "librtc_b200.so")   # override: kernel experiments

# every symbol include/rtc.h declares (tests check the library exports all of them)
EXPORTS = [
    "rtc_create", "rtc_destroy", "rtc_last_error", "rtc_version", "rtc_set_stream", "rtc_device_info", "rtc_synchronize", "rtc_set_light",
    "rtc_resize", "rtc_scene_clear", "rtc_scene_add_sphere", "rtc_scene_add_plane", "rtc_scene_set_objects",
    "rtc_scene_get_objects", "rtc_scene_count", "rtc_update_objects", "rtc_render", "rtc_frame_ansi",
    "rtc_frame_ansi_device", "rtc_frame_color", "rtc_frame_hits", "rtc_debug_ansi256_cube", "rtc_update", "rtc_submit", "rtc_collect", "rtc_last_timings",
    "rtc_trace_band", "rtc_trace_raw", "rtc_raw_size", "rtc_encode", "rtc_encode_band", "rtc_encode_capacity", "rtc_mode_bpp", "rtc_mode_has_glyph",
    "rtc_camera_params", "rtc_fp32_peak",
    "rtc_mgpu_create", "rtc_mgpu_destroy", "rtc_mgpu_count", "rtc_mgpu_context", "rtc_mgpu_scene_clear",
    "rtc_mgpu_scene_add_sphere", "rtc_mgpu_scene_add_plane", "rtc_mgpu_scene_set_objects", "rtc_mgpu_scene_get_objects",
    "rtc_mgpu_set_light", "rtc_mgpu_submit", "rtc_mgpu_collect", "rtc_mgpu_update", "rtc_mgpu_last_frame",
    "rtc_mgpu_set_bands", "rtc_mgpu_flush_l2", "rtc_mgpu_host_stats", "rtc_mgpu_debug_trace", "rtc_plan_bands",
]

GATHER_HOST, GATHER_P2P = 0, 1


class RtcError(RuntimeError):
    pass


_lib = None


def load_library(build_if_missing=True):
    """dlopen librtc_b200.so (building it in-tree with nvcc first if needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from .build import build
        build()
    if not os.path.exists(LIB_PATH):
        raise RtcError("librtc_b200.so is missing: run __graft_entry__.build() (nvcc, sm_100a)")
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, u32, i32, f32, f64, sz = c.c_void_p, c.c_uint32, c.c_int, c.c_float, c.c_double, c.c_size_t
    L.rtc_last_error.restype = c.c_char_p
    L.rtc_version.restype = c.c_char_p
    L.rtc_encode_capacity.restype = sz
    L.rtc_encode_capacity.argtypes = [u32, u32, i32]
    L.rtc_mode_bpp.restype = u32
    L.rtc_mode_has_glyph.restype = u32
    L.rtc_create.argtypes = [c.POINTER(vp), i32]
    L.rtc_destroy.argtypes = [vp]
    L.rtc_destroy.restype = None
    L.rtc_set_stream.argtypes = [vp, vp]
    L.rtc_device_info.argtypes = [vp, c.POINTER(i32), c.POINTER(i32), c.POINTER(sz)]
    L.rtc_resize.argtypes = [vp, u32, u32]
    L.rtc_set_light.argtypes = [vp, vp]
    L.rtc_synchronize.argtypes = [vp]
    L.rtc_scene_clear.argtypes = [vp]
    L.rtc_scene_add_sphere.argtypes = [vp, vp, f32, vp, f32, i32]
    L.rtc_scene_add_plane.argtypes = [vp, vp, vp, vp, f32, f32]
    L.rtc_scene_set_objects.argtypes = [vp, vp, u32]
    L.rtc_scene_get_objects.argtypes = [vp, vp, u32, c.POINTER(u32)]
    L.rtc_scene_count.argtypes = [vp, c.POINTER(u32)]
    L.rtc_update_objects.argtypes = [vp, f64, u32]
    L.rtc_render.argtypes = [vp, vp, i32, u32]
    L.rtc_frame_ansi.argtypes = [vp, c.POINTER(vp), c.POINTER(sz)]
    L.rtc_frame_ansi_device.argtypes = [vp, c.POINTER(vp), c.POINTER(sz)]
    L.rtc_frame_color.argtypes = [vp, c.POINTER(vp), c.POINTER(u32), c.POINTER(vp)]
    L.rtc_frame_hits.argtypes = [vp, c.POINTER(vp), c.POINTER(vp)]
    L.rtc_debug_ansi256_cube.argtypes = [vp, vp]
    L.rtc_update.argtypes = [vp, vp, i32, f64, u32, c.POINTER(vp), c.POINTER(sz)]
    L.rtc_submit.argtypes = [vp, vp, i32, f64, u32]
    L.rtc_collect.argtypes = [vp, c.POINTER(vp), c.POINTER(sz)]
    L.rtc_last_timings.argtypes = [vp, vp]
    L.rtc_trace_band.argtypes = [vp, vp, i32, u32, u32, u32, vp, vp]
    L.rtc_encode.argtypes = [vp, vp, vp, u32, u32, i32, vp, sz, vp]
    L.rtc_trace_raw.argtypes = [vp, vp, i32, u32, vp]
    L.rtc_raw_size.restype = sz
    L.rtc_raw_size.argtypes = [u32, u32]
    L.rtc_encode_band.argtypes = [vp, vp, vp, u32, u32, i32, i32, vp, sz, vp]
    L.rtc_camera_params.argtypes = [u32, u32, vp, vp, f32, vp]
    L.rtc_fp32_peak.argtypes = [vp, i32, i32, c.POINTER(f32), c.POINTER(f32)]
    L.rtc_mgpu_create.argtypes = [c.POINTER(vp), i32, vp, i32]
    L.rtc_mgpu_destroy.argtypes = [vp]
    L.rtc_mgpu_destroy.restype = None
    L.rtc_mgpu_count.argtypes = [vp]
    L.rtc_mgpu_context.argtypes = [vp, i32, c.POINTER(vp)]
    L.rtc_mgpu_scene_clear.argtypes = [vp]
    L.rtc_mgpu_scene_add_sphere.argtypes = [vp, vp, f32, vp, f32, i32]
    L.rtc_mgpu_scene_add_plane.argtypes = [vp, vp, vp, vp, f32, f32]
    L.rtc_mgpu_scene_set_objects.argtypes = [vp, vp, u32]
    L.rtc_mgpu_scene_get_objects.argtypes = [vp, vp, u32, c.POINTER(u32)]
    L.rtc_mgpu_set_light.argtypes = [vp, vp]
    L.rtc_mgpu_submit.argtypes = [vp, vp, i32, f64, u32]
    L.rtc_mgpu_collect.argtypes = [vp, c.POINTER(vp), c.POINTER(sz)]
    L.rtc_mgpu_update.argtypes = [vp, vp, i32, f64, u32, c.POINTER(vp), c.POINTER(sz)]
    L.rtc_mgpu_last_frame.argtypes = [vp, vp, vp, vp]
    L.rtc_mgpu_set_bands.argtypes = [vp, u32, vp]
    L.rtc_mgpu_flush_l2.argtypes = [vp]
    L.rtc_mgpu_host_stats.argtypes = [vp, vp]
    L.rtc_mgpu_debug_trace.argtypes = [vp, vp]
    L.rtc_plan_bands.argtypes = [u32, i32, u32, f64, u32, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RtcError("rtc error %d: %s" % (rc, load_library().rtc_last_error().decode(errors="replace")))


def encode_capacity(x, y, mode):
    return int(load_library().rtc_encode_capacity(x, y, mode))


def camera_params(x, y, pos, rot, pixel_aspect=0.0):
    """Camera3D::Init/Update/GetInverseVMatrix + Engine3D::Render's block, on the host (no GPU)."""
    p = RtcParams()
    pos = np.ascontiguousarray(pos, np.float32)
    rot = np.ascontiguousarray(rot, np.float32)
    _check(load_library().rtc_camera_params(x, y, pos.ctypes.data, rot.ctypes.data, float(pixel_aspect), ctypes.byref(p)))
    return p


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


def plan_bands(y, n, align=1, deficit_rows=0.0, wave_units=0):
    """rtc_plan_bands (host only): the row bands [(r0, r1)] * n of a y-row frame."""
    rows = (ctypes.c_uint32 * (n + 1))()
    _check(load_library().rtc_plan_bands(y, n, align, float(deficit_rows), wave_units, rows))
    return [(int(rows[g]), int(rows[g + 1])) for g in range(n)]


class MultiGpu:
    """rtc_mgpu: RayTracingManager::Update across the GPUs of one box, one process (row bands)."""

    def __init__(self, devices, gather=GATHER_HOST):
        self.L = load_library()
        self._h = ctypes.c_void_p()
        devices = list(range(devices)) if isinstance(devices, int) else list(devices)
        ids = (ctypes.c_int * len(devices))(*devices)
        _check(self.L.rtc_mgpu_create(ctypes.byref(self._h), len(devices), ids, gather))
        self.devices, self.n, self.gather = devices, len(devices), gather

    def close(self):
        if self._h:
            self.L.rtc_mgpu_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_objects(self, objs):
        objs = np.ascontiguousarray(objs, OBJECT_DTYPE)
        _check(self.L.rtc_mgpu_scene_set_objects(self._h, obj_ptr(objs), len(objs)))

    def clear(self):
        _check(self.L.rtc_mgpu_scene_clear(self._h))

    def add_sphere(self, center, radius, rgb, speed=1.0, mover=-1):
        c = np.ascontiguousarray(center, np.float32)
        k = np.ascontiguousarray(rgb, np.float32)
        _check(self.L.rtc_mgpu_scene_add_sphere(self._h, c.ctypes.data, float(radius), k.ctypes.data, float(speed), int(mover)))

    def add_plane(self, center, normal, rgb, width, height):
        c = np.ascontiguousarray(center, np.float32)
        n = np.ascontiguousarray(normal, np.float32)
        k = np.ascontiguousarray(rgb, np.float32)
        _check(self.L.rtc_mgpu_scene_add_plane(self._h, c.ctypes.data, n.ctypes.data, k.ctypes.data, float(width), float(height)))

    def get_objects(self):
        n = ctypes.c_uint32()
        _check(self.L.rtc_mgpu_scene_get_objects(self._h, None, 0, ctypes.byref(n)))
        out = np.zeros(n.value, OBJECT_DTYPE)
        _check(self.L.rtc_mgpu_scene_get_objects(self._h, obj_ptr(out), n.value, ctypes.byref(n)))
        return out

    def set_light(self, light=None):
        if light is None:
            _check(self.L.rtc_mgpu_set_light(self._h, None))
        else:
            a = np.ascontiguousarray(light, np.float32)
            assert a.size == 11
            _check(self.L.rtc_mgpu_set_light(self._h, a.ctypes.data))

    def submit(self, params, mode, dt=0.0, flags=0):
        _check(self.L.rtc_mgpu_submit(self._h, ctypes.byref(params), mode, float(dt), flags))

    def collect(self, copy=False):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.rtc_mgpu_collect(self._h, ctypes.byref(p), ctypes.byref(n)))
        v = _view(p.value, n.value, np.uint8)
        return v.copy() if copy else v

    def update(self, params, mode, dt=0.0, flags=0, copy=False):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.rtc_mgpu_update(self._h, ctypes.byref(params), mode, float(dt), flags, ctypes.byref(p), ctypes.byref(n)))
        v = _view(p.value, n.value, np.uint8)
        return v.copy() if copy else v

    def last_frame(self):
        ms = (ctypes.c_float * self.n)()
        rows = (ctypes.c_uint32 * (self.n + 1))()
        enc = ctypes.c_float()
        _check(self.L.rtc_mgpu_last_frame(self._h, ms, rows, ctypes.byref(enc)))
        return dict(device_ms=[float(v) for v in ms], bands=[[int(rows[g]), int(rows[g + 1])] for g in range(self.n)],
                    encode_ms=float(enc.value))

    def set_bands(self, y, rows=None):
        if rows is None:
            _check(self.L.rtc_mgpu_set_bands(self._h, y, None))
        else:
            a = (ctypes.c_uint32 * (self.n + 1))(*rows)
            _check(self.L.rtc_mgpu_set_bands(self._h, y, a))

    def flush_l2(self):
        _check(self.L.rtc_mgpu_flush_l2(self._h))

    def debug_trace(self):
        """(workers[n][64][6], main[64][2]) host-clock timeline of the last 64 frames, in us."""
        a = np.zeros((self.n * 6 + 2) * 64, np.float64)
        _check(self.L.rtc_mgpu_debug_trace(self._h, a.ctypes.data))
        return a[:self.n * 64 * 6].reshape(self.n, 64, 6), a[self.n * 64 * 6:].reshape(64, 2)

    def host_stats(self):
        """Per device, average us per frame since the last call: enqueue, wait for lengths, D2H copy."""
        a = (ctypes.c_float * (3 * self.n))()
        _check(self.L.rtc_mgpu_host_stats(self._h, a))
        return {"enqueue_us": [round(float(a[3 * g]), 1) for g in range(self.n)],
                "wait_lengths_us": [round(float(a[3 * g + 1]), 1) for g in range(self.n)],
                "copy_us": [round(float(a[3 * g + 2]), 1) for g in range(self.n)]}

    def device_info(self, i=0):
        h = ctypes.c_void_p()
        _check(self.L.rtc_mgpu_context(self._h, i, ctypes.byref(h)))
        sm, clk, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        _check(self.L.rtc_device_info(h, ctypes.byref(sm), ctypes.byref(clk), ctypes.byref(smem)))
        return dict(sm_count=sm.value, clock_khz=clk.value, smem_optin=smem.value)


class Context:
    """One rtc_ctx == one GPU.  Mirrors the reference's RayTracingManager + Scene3D pair."""

    def __init__(self, device=0):
        self.L = load_library()
        self._h = ctypes.c_void_p()
        _check(self.L.rtc_create(ctypes.byref(self._h), device))
        self.device = device

    def close(self):
        if self._h:
            self.L.rtc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing --------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        _check(self.L.rtc_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        _check(self.L.rtc_synchronize(self._h))

    def device_info(self):
        sm, clk, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        _check(self.L.rtc_device_info(self._h, ctypes.byref(sm), ctypes.byref(clk), ctypes.byref(smem)))
        return dict(sm_count=sm.value, clock_khz=clk.value, smem_optin=smem.value)

    def set_light(self, light=None):
        """light: 11 floats (pos[3], diffuse colour, diffuse power, specular colour, specular power, ambient[3], object
        specular) or None for the reference's constants."""
        if light is None:
            _check(self.L.rtc_set_light(self._h, None))
        else:
            a = np.ascontiguousarray(light, np.float32)
            assert a.size == 11
            _check(self.L.rtc_set_light(self._h, a.ctypes.data))

    # -- scene (Scene3D) ---------------------------------------------------------------------
    def set_objects(self, objs):
        objs = np.ascontiguousarray(objs, OBJECT_DTYPE)
        _check(self.L.rtc_scene_set_objects(self._h, obj_ptr(objs), len(objs)))

    def clear(self):
        _check(self.L.rtc_scene_clear(self._h))

    def add_sphere(self, center, radius, rgb, speed=1.0, mover=-1):
        c = np.ascontiguousarray(center, np.float32)
        k = np.ascontiguousarray(rgb, np.float32)
        _check(self.L.rtc_scene_add_sphere(self._h, c.ctypes.data, float(radius), k.ctypes.data, float(speed), int(mover)))

    def add_plane(self, center, normal, rgb, width, height):
        c = np.ascontiguousarray(center, np.float32)
        n = np.ascontiguousarray(normal, np.float32)
        k = np.ascontiguousarray(rgb, np.float32)
        _check(self.L.rtc_scene_add_plane(self._h, c.ctypes.data, n.ctypes.data, k.ctypes.data, float(width), float(height)))

    def get_objects(self):
        n = ctypes.c_uint32()
        _check(self.L.rtc_scene_count(self._h, ctypes.byref(n)))
        out = np.zeros(n.value, OBJECT_DTYPE)
        _check(self.L.rtc_scene_get_objects(self._h, obj_ptr(out), n.value, ctypes.byref(n)))
        return out

    def update_objects(self, dt, flags=0):
        _check(self.L.rtc_update_objects(self._h, float(dt), flags))

    # -- frame (RayTracingManager::Update) -----------------------------------------------------
    def render(self, params, mode, flags=0):
        _check(self.L.rtc_render(self._h, ctypes.byref(params), mode, flags))

    def frame_ansi(self, copy=True):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.rtc_frame_ansi(self._h, ctypes.byref(p), ctypes.byref(n)))
        v = _view(p.value, n.value, np.uint8)
        return v.copy() if copy else v

    def frame_ansi_device(self):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.rtc_frame_ansi_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def frame_color(self, n_px):
        pc, pg, bpp = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint32()
        _check(self.L.rtc_frame_color(self._h, ctypes.byref(pc), ctypes.byref(bpp), ctypes.byref(pg)))
        color = _view(pc.value, n_px * bpp.value, np.uint8).copy()
        glyph = _view(pg.value, n_px, np.uint8).copy() if pg.value else None
        return color, glyph

    def frame_hits(self, n_px):
        pd, pi = ctypes.c_void_p(), ctypes.c_void_p()
        _check(self.L.rtc_frame_hits(self._h, ctypes.byref(pd), ctypes.byref(pi)))
        return _view(pd.value, n_px, np.float32).copy(), _view(pi.value, n_px, np.int32).copy()

    def ansi256_cube(self, dev_out):
        _check(self.L.rtc_debug_ansi256_cube(self._h, ctypes.c_void_p(dev_out)))

    def update(self, params, mode, dt=0.0, flags=0):
        """RayTracingManager::Update: physics step + render + stream to host."""
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.rtc_update(self._h, ctypes.byref(params), mode, float(dt), flags, ctypes.byref(p), ctypes.byref(n)))
        return _view(p.value, n.value, np.uint8)

    def submit(self, params, mode, dt=0.0, flags=0):
        """Pipelined Update, part 1: enqueue physics + render (returns at once; at most two frames in flight)."""
        _check(self.L.rtc_submit(self._h, ctypes.byref(params), mode, float(dt), flags))

    def collect(self, copy=False):
        """Pipelined Update, part 2: the oldest submitted frame's stream, in pinned host memory."""
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(self.L.rtc_collect(self._h, ctypes.byref(p), ctypes.byref(n)))
        v = _view(p.value, n.value, np.uint8)
        return v.copy() if copy else v

    def timings(self):
        t = RtcTimings()
        _check(self.L.rtc_last_timings(self._h, ctypes.byref(t)))
        return dict(prep_ms=t.prep_ms, trace_ms=t.trace_ms, shade_ms=t.shade_ms, encode_ms=t.encode_ms,
                    total_ms=t.total_ms, launches=t.launches, sphere_tests=int(t.sphere_tests))

    # -- stage level, caller-owned device memory (raw pointers, e.g. torch tensors' data_ptr()) ---
    def trace_band(self, params, mode, row0, row1, dev_color, dev_glyph=0, flags=0):
        _check(self.L.rtc_trace_band(self._h, ctypes.byref(params), mode, flags, row0, row1,
                                     ctypes.c_void_p(dev_color), ctypes.c_void_p(dev_glyph or 0)))

    def trace_raw(self, params, mode, dev_result, flags=0):
        """RayTracing::RayTrace's raw 20*x*y-byte cell buffer into caller-owned device memory (asynchronous)."""
        _check(self.L.rtc_trace_raw(self._h, ctypes.byref(params), mode, flags, ctypes.c_void_p(dev_result)))

    def encode(self, dev_color, dev_glyph, x, y, mode, dev_out, cap, dev_total):
        _check(self.L.rtc_encode(self._h, ctypes.c_void_p(dev_color), ctypes.c_void_p(dev_glyph or 0), x, y, mode,
                                 ctypes.c_void_p(dev_out), cap, ctypes.c_void_p(dev_total)))

    def encode_band(self, dev_color, dev_glyph, x, rows, mode, continues, dev_out, cap, dev_total):
        _check(self.L.rtc_encode_band(self._h, ctypes.c_void_p(dev_color), ctypes.c_void_p(dev_glyph or 0), x, rows, mode,
                                      1 if continues else 0, ctypes.c_void_p(dev_out), cap, ctypes.c_void_p(dev_total)))

    def fp32_peak(self, variant, iters=2000):
        tf, ms = ctypes.c_float(), ctypes.c_float()
        _check(self.L.rtc_fp32_peak(self._h, variant, iters, ctypes.byref(tf), ctypes.byref(ms)))
        return tf.value, ms.value
