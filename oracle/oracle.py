"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle.

Two libraries:
  * librt_oracle.so  -- the plain-C restatement (oracle/rt_oracle.c), always available;
  * _ref/libref_cpu.so -- the reference's OWN sources compiled for CPU through shims
    (oracle/ref_build), available where it was built (it travels to the GPU box as a
    prebuilt file; /root/reference itself does not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


sys.path.insert(0, _ROOT) if _ROOT not in sys.path else None
import rtc_b200  # noqa: E402,F401  (root-level import shim for the hyphenated package directory)
from rtc_b200._types import (OBJECT_DTYPE, RtcParams, mode_bpp, mode_cell,  # noqa: E402
                             mode_has_glyph, obj_ptr)

_c = ctypes
_u8p = _c.POINTER(_c.c_uint8)


def build(force=False):
    """Compile the restated oracle and (where the reference tree exists) the reference."""
    so = os.path.join(_HERE, "librt_oracle.so")
    src = os.path.join(_HERE, "rt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "librt_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/ConsoleProject"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _arr(a, dtype=np.uint8):
    return a.ctypes.data_as(_c.c_void_p)


class Oracle:
    """Front-end of librt_oracle.so (the restatement)."""

    def __init__(self):
        build()
        L = _c.CDLL(os.path.join(_HERE, "librt_oracle.so"))
        self.L = L
        L.orc_minimize.restype = _c.c_size_t
        L.orc_encode_planes.restype = _c.c_size_t
        L.orc_render.restype = _c.c_size_t
        L.orc_default_scene.restype = _c.c_uint32
        L.orc_time_trace.restype = _c.c_double
        L.orc_minimize.argtypes = [_c.c_void_p, _c.c_size_t, _c.c_uint32, _c.c_uint32, _c.c_int, _c.c_void_p]
        L.orc_encode_planes.argtypes = [_c.c_void_p, _c.c_void_p, _c.c_uint32, _c.c_uint32, _c.c_int, _c.c_void_p]
        L.orc_trace_planes.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_int, _c.c_uint32, _c.c_uint32,
                                       _c.c_uint32, _c.c_int] + [_c.c_void_p] * 5
        L.orc_trace_raw.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_int, _c.c_uint32, _c.c_int, _c.c_void_p]
        L.orc_update_objects.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_double, _c.c_uint32]
        L.orc_camera_params.argtypes = [_c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_float, _c.c_void_p]
        L.orc_time_trace.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_int, _c.c_uint32, _c.c_uint32,
                                     _c.c_uint32, _c.c_int]
        L.orc_sphere_trace.argtypes = [_c.c_void_p, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]
        L.orc_plane_trace.argtypes = [_c.c_void_p, _c.c_void_p, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p,
                                      _c.c_void_p, _c.c_void_p]
        L.orc_ascii_char.argtypes = [_c.c_float, _c.c_float, _c.c_float]
        L.orc_raytrace.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_uint32] + [_c.c_void_p] * 5

    # -- frames ---------------------------------------------------------------------------
    def trace_planes(self, objs, params, mode, flags=0, row0=0, row1=None, nthreads=8):
        W = params.x - 1
        row1 = params.y if row1 is None else row1
        n = (row1 - row0) * W
        color = np.zeros(n * mode_bpp(mode), np.uint8)
        glyph = np.zeros(n, np.uint8)
        hit = np.zeros(n, np.uint8)
        dist = np.zeros(n, np.float32)
        index = np.zeros(n, np.int32)
        rc = self.L.orc_trace_planes(obj_ptr(objs), len(objs), _c.byref(params), mode, flags, row0, row1, nthreads,
                                     _arr(color), _arr(glyph), _arr(hit), _arr(dist), _arr(index))
        assert rc == 0
        return dict(color=color, glyph=glyph, hit=hit, dist=dist, index=index)

    def trace_raw(self, objs, params, mode, flags=0, nthreads=8):
        raw = np.zeros(20 * params.x * params.y, np.uint8)
        assert self.L.orc_trace_raw(obj_ptr(objs), len(objs), _c.byref(params), mode, flags, nthreads, _arr(raw)) == 0
        return raw

    def minimize(self, raw, x, y, mode):
        out = np.zeros(raw.size + y + 16, np.uint8)
        n = self.L.orc_minimize(_arr(raw), raw.size, x, y, mode, _arr(out))
        return out[:n].copy()

    def encode_planes(self, color, glyph, x, y, mode):
        out = np.zeros((x - 1) * y * mode_cell(mode) + y + 16, np.uint8)
        n = self.L.orc_encode_planes(_arr(color), _arr(glyph) if glyph is not None else None, x, y, mode, _arr(out))
        return out[:n].copy()

    def render(self, objs, params, mode, flags=0, nthreads=8):
        """RayTracingManager::Update (dt = 0): the minimised ANSI stream."""
        raw = self.trace_raw(objs, params, mode, flags, nthreads)
        return self.minimize(raw, params.x, params.y, mode)

    def update_objects(self, objs, dt, flags=0):
        o = objs.copy()
        self.L.orc_update_objects(obj_ptr(o), len(o), float(dt), flags)
        return o

    def camera_params(self, x, y, pos, rot, pixel_aspect=0.0):
        p = RtcParams()
        pos = np.asarray(pos, np.float32)
        rot = np.asarray(rot, np.float32)
        assert self.L.orc_camera_params(x, y, _arr(pos), _arr(rot), pixel_aspect, _c.byref(p)) == 0
        return p

    def default_scene(self):
        o = np.zeros(6, OBJECT_DTYPE)
        assert self.L.orc_default_scene(obj_ptr(o), 6) == 6
        return o

    def time_trace(self, objs, params, mode, row0, row1, nthreads, flags=0):
        return self.L.orc_time_trace(obj_ptr(objs), len(objs), _c.byref(params), mode, flags, row0, row1, nthreads)

    def ansi256_range(self, first, count):
        out = np.zeros(count, np.uint8)
        self.L.orc_ansi256_range(first, count, _arr(out))
        return out

    def set_light(self, light=None):
        """The 11 floats of rtc_light, or None for the reference's constants (process-wide state of the checker)."""
        if light is None:
            self.L.orc_set_light(None)
        else:
            a = np.ascontiguousarray(light, np.float32)
            assert a.size == 11
            self.L.orc_set_light(_arr(a))


class Reference:
    """Front-end of oracle/_ref/libref_cpu.so: the reference's own code on the CPU."""

    @staticmethod
    def path():
        return os.path.join(_HERE, "_ref", "libref_cpu.so")

    @staticmethod
    def available():
        return os.path.exists(Reference.path())

    def __init__(self):
        L = _c.CDLL(self.path())
        self.L = L
        L.ref_update.restype = _c.c_double
        L.ref_update.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_int, _c.c_void_p, _c.c_int, _c.c_double, _c.c_int,
                                 _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_void_p]
        L.ref_trace_blockrows.restype = _c.c_double
        L.ref_trace_blockrows.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_int, _c.c_void_p, _c.c_int, _c.c_int,
                                          _c.c_uint32, _c.c_uint32, _c.c_void_p]
        L.ref_camera_params.argtypes = [_c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_void_p, _c.c_void_p]
        L.ref_sphere_trace.argtypes = [_c.c_void_p, _c.c_float, _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_void_p]
        L.ref_plane_trace.argtypes = [_c.c_void_p, _c.c_void_p, _c.c_float, _c.c_float, _c.c_void_p, _c.c_void_p,
                                      _c.c_void_p, _c.c_void_p]
        L.ref_ascii_char.argtypes = [_c.c_float, _c.c_float, _c.c_float]
        L.ref_raytrace.argtypes = [_c.c_void_p, _c.c_uint32, _c.c_void_p, _c.c_void_p] + [_c.c_void_p] * 4

    def update(self, objs, params, mode, dt=0.0, nthreads=1, want_raw=False, default_scene=False, want_objs=False):
        x, y = params.x, params.y
        raw = np.zeros(20 * x * y, np.uint8) if want_raw else None
        mn = np.zeros(20 * x * y + 16, np.uint8)
        n = _c.c_size_t()
        n_obj = 0 if default_scene else len(objs)
        after = np.zeros(max(n_obj, 6), OBJECT_DTYPE) if want_objs else None
        secs = self.L.ref_update(None if default_scene else obj_ptr(objs), n_obj, 1 if default_scene else 0,
                                 _c.byref(params), mode, float(dt), nthreads,
                                 _arr(raw) if want_raw else None, raw.size if want_raw else 0,
                                 _arr(mn), mn.size, _c.byref(n), _arr(after) if want_objs else None)
        res = dict(stream=mn[: n.value].copy(), secs=secs)
        if want_raw:
            res["raw"] = raw
        if want_objs:
            res["objs"] = after[: max(n_obj, 6 if default_scene else 0)]
        return res

    def trace_blockrows(self, objs, params, mode, brow0, brow1, nthreads):
        rays = _c.c_ulonglong()
        secs = self.L.ref_trace_blockrows(obj_ptr(objs), len(objs), 0, _c.byref(params), mode, nthreads,
                                          brow0, brow1, _c.byref(rays))
        return secs, rays.value

    def camera_params(self, x, y, pos, rot):
        p = RtcParams()
        pos = np.asarray(pos, np.float32)
        rot = np.asarray(rot, np.float32)
        self.L.ref_camera_params(x, y, _arr(pos), _arr(rot), _c.byref(p))
        return p

    def ansi256_range(self, first, count):
        out = np.zeros(count, np.uint8)
        self.L.ref_ansi256_range(first, count, _arr(out))
        return out


def planes_from_raw(raw, x, y, mode):
    """Parse the reference's raw cell buffer into colour / glyph / selector planes
    (SURVEY appendix A.5): cell at (row*x + col)*SIZE; channel = decimal of the non-NUL
    digit bytes."""
    cs = mode_cell(mode)
    W = x - 1
    cells = raw[: cs * x * y].reshape(y, x, cs)[:, :W, :]

    def dec(b3):
        d = np.where(b3 == 0, 0, b3.astype(np.int32) - 48)
        return (d[..., 0] * 100 + d[..., 1] * 10 + d[..., 2]).astype(np.uint8)

    if cs == 20:
        color = np.stack([dec(cells[..., 7:10]), dec(cells[..., 11:14]), dec(cells[..., 15:18])], -1)
    else:
        color = dec(cells[..., 7:10])[..., None]
    glyph = cells[..., cs - 1].copy()
    fg = (cells[..., 2] == ord("3")).astype(np.uint8)
    return color.reshape(-1), glyph.reshape(-1), fg.reshape(-1)
