"""ctypes / numpy mirrors of the PODs in include/rtc.h (no GPU needed)."""
import ctypes

import numpy as np


class RtcParams(ctypes.Structure):
    """== rtc_params == RayTracingCPUToGPUData minus vptrs (reference RayTracingManager.h:9-19)."""
    _fields_ = [
        ("inv_view", ctypes.c_float * 16),
        ("cam_pos", ctypes.c_float * 3),
        ("x", ctypes.c_uint32),
        ("y", ctypes.c_uint32),
        ("element1", ctypes.c_float),
        ("element2", ctypes.c_float),
        ("cam_far", ctypes.c_float),
    ]

    def copy(self):
        q = RtcParams()
        ctypes.memmove(ctypes.byref(q), ctypes.byref(self), ctypes.sizeof(self))
        return q


class RtcTimings(ctypes.Structure):
    _fields_ = [
        ("prep_ms", ctypes.c_float),
        ("trace_ms", ctypes.c_float),
        ("shade_ms", ctypes.c_float),
        ("encode_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float),
        ("launches", ctypes.c_uint32),
        ("reserved_", ctypes.c_uint32),
        ("sphere_tests", ctypes.c_uint64),
    ]


# == rtc_object (64 bytes): union of the reference's Sphere / Plane state.
OBJECT_DTYPE = np.dtype(
    [
        ("type", "<i4"),
        ("center", "<f4", (3,)),
        ("color", "<f4", (3,)),
        ("radius", "<f4"),
        ("normal", "<f4", (3,)),
        ("width", "<f4"),
        ("height", "<f4"),
        ("speed", "<f4"),
        ("mover", "<i4"),
        ("reserved_", "<i4"),
    ],
    align=False,
)
assert OBJECT_DTYPE.itemsize == 64
assert ctypes.sizeof(RtcParams) == 96

OBJ_NONE, OBJ_PLANE, OBJ_SPHERE = 0, 1, 2

# == RenderingMode order (reference RayTracingManager.h:21)
BIT_ASCII, BIT_PIXEL, RGB_ASCII, RGB_PIXEL, RGB_NORMALS, SDL = range(6)
MODE_NAMES = ["BIT_ASCII", "BIT_PIXEL", "RGB_ASCII", "RGB_PIXEL", "RGB_NORMALS", "SDL"]

FLAG_SHADOWS = 1
FLAG_UPDATE_REF_LAUNCH_LIMIT = 2
FLAG_CULL = 4
FLAG_KEEP_HITS = 8
FLAG_NORMALS_SATURATE = 16
FLAG_PACKET = 32


def mode_bpp(mode):
    return 1 if mode in (BIT_ASCII, BIT_PIXEL) else 3


def mode_cell(mode):
    """SIZE_8BIT / SIZE_RGB (reference RayTracing.h:120-123)."""
    return 12 if mode in (BIT_ASCII, BIT_PIXEL) else 20


def mode_has_glyph(mode):
    return mode in (BIT_ASCII, RGB_ASCII)


def obj_ptr(objs):
    return objs.ctypes.data_as(ctypes.c_void_p) if objs is not None and len(objs) else None
