"""Synthetic scenes and cameras of the BASELINE.json configs.

Host-side only (numpy); no oracle, no GPU.  The sphere distribution is the reference's own
"random sphere per second" generator (reference Engine3D.cpp:63: radius rand()%10, centre
rand()%100-50 per axis, colour rand()%255 per channel; Sphere.cu:11-12: speed
(rand()%300+100)/100), driven by a portable SplitMix64 instead of rand().
"""
import math

import numpy as np

from ._types import OBJECT_DTYPE, OBJ_PLANE, OBJ_SPHERE, RtcParams

_MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        return z ^ (z >> 31)


def make_sphere(center, radius, color, speed=1.0, mover=-1):
    o = np.zeros((), OBJECT_DTYPE)
    o["type"] = OBJ_SPHERE
    o["center"] = center
    o["radius"] = radius
    o["color"] = color
    o["speed"] = speed
    o["mover"] = mover
    return o


def normalize_host(n):
    """MyMath::Vector3::Normalize (reference MyMath.h:117-123) in float32."""
    n = np.asarray(n, np.float32)
    length = np.sqrt(np.float32(n[0] * n[0] + n[1] * n[1]) + n[2] * n[2], dtype=np.float32)
    div = np.float32(0.0) if length < np.float32(0.000001) else np.float32(1.0) / length
    return (n * div).astype(np.float32)


def make_plane(center, normal, color, width, height):
    """Plane ctor (reference Plane.cu:6-12): the stored normal is normalised."""
    o = np.zeros((), OBJECT_DTYPE)
    o["type"] = OBJ_PLANE
    o["center"] = center
    o["normal"] = normalize_host(normal)
    o["color"] = color
    o["width"] = width
    o["height"] = height
    return o


def default_scene():
    """Scene3D::Init's temporary scene (reference Scene3D.cpp:28-33)."""
    objs = [
        make_sphere((0.0, 10.0, 20.0), 7.0, (255.0, 1.0, 1.0)),
        make_sphere((5.0, 10.0, 20.0), 6.0, (1.0, 255.0, 1.0)),
        make_sphere((10.0, 10.0, 40.0), 10.0, (1.0, 1.0, 255.0)),
        make_sphere((5.0, 10.0, 20.0), 3.0, (225.0, 210.0, 20.0)),
        make_sphere((-5.0, 10.0, 40.0), 4.0, (225.0, 10.0, 220.0)),
        make_plane((0.0, -3.0, 30.0), (0.0, 1.0, 0.0), (100.0, 100.0, 100.0), 10.0, 20.0),
    ]
    return np.array(objs, OBJECT_DTYPE)


def random_spheres(n, seed):
    rng = SplitMix64(seed)
    objs = np.zeros(n, OBJECT_DTYPE)
    for i in range(n):
        r = float(rng.next() % 10)
        c = [float(rng.next() % 100) - 50.0 for _ in range(3)]
        col = [float(rng.next() % 255) for _ in range(3)]
        speed = np.float32(float(rng.next() % 300 + 100)) / np.float32(100.0)
        objs[i] = make_sphere(c, r, col, speed, -1)
    return objs


def bench_plane():
    return make_plane((0.0, -60.0, 0.0), (0.0, 1.0, 0.0), (100.0, 100.0, 100.0), 400.0, 400.0)


# name -> (x, y, n_spheres, with_plane, seed)      x = W + 1 (newline column)
CONFIGS = {
    "config1_240x64": (240, 64, 0, False, 0),                    # reference default scene
    "config1_400x150": (400, 150, 0, False, 0),
    "config2_1080p_64": (1921, 1080, 64, True, 0x5EED0002),
    "config3_4k_1024": (3841, 2160, 1024, True, 0x5EED0003),
    "config4_8k_4096": (7681, 4320, 4096, False, 0x5EED0004),
}


def config_scene(name):
    x, y, n, with_plane, seed = CONFIGS[name]
    if n == 0:
        return default_scene()
    objs = random_spheres(n, seed)
    if with_plane:
        objs = np.concatenate([objs, np.array([bench_plane()], OBJECT_DTYPE)])
    return objs


def camera_params(x, y, pos, rot, pixel_aspect=0.0):
    """Camera3D::Init/Update/GetInverseVMatrix + Engine3D::Render's parameter block (reference
    Camera3D.cpp:8-48,:51-98,:207-376; Engine3D.cpp:88-97) -- computed by the library's host code
    (csrc/rtc_camera.cpp via rtc_camera_params; no GPU needed)."""
    from . import camera_params as _camera_params
    return _camera_params(x, y, pos, rot, pixel_aspect)


def config_camera(name, frame=0, n_frames=120, camera_fn=None):
    """Cameras of SURVEY 8d.  config1: reference default camera (origin, rot (0,pi,0),
    pixel aspect 0.01).  configs 2-3: pos (0,0,-120), rot (0,pi,0), pixel aspect 1/W.
    config 4: orbit R=120 about the origin, phi_k = pi + 2*pi*k/n_frames."""
    cam = camera_fn or camera_params               # camera_fn: another restatement of Camera3D with the same signature (the oracle's)
    x, y, n, _, _ = CONFIGS[name]
    if n == 0:
        return cam(x, y, (0.0, 0.0, 0.0), (0.0, np.float32(math.pi), 0.0), 0.0)
    k = 1.0 / float(x - 1)
    if name.startswith("config4"):
        phi = math.pi + 2.0 * math.pi * frame / n_frames
        pos = (120.0 * math.sin(phi), 0.0, 120.0 * math.cos(phi))
        return cam(x, y, pos, (0.0, phi, 0.0), k)
    return cam(x, y, (0.0, 0.0, -120.0), (0.0, np.float32(math.pi), 0.0), k)
