"""Generate tests/golden/golden.npz from the REFERENCE ITSELF (oracle/_ref/libref_cpu.so, i.e.
the unmodified sources under /root/reference compiled for CPU by oracle/ref_build).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.
"""
import ctypes
import hashlib
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, planes_from_raw  # noqa: E402
import rtc_b200  # noqa: E402
from rtc_b200 import scenes  # noqa: E402
from rtc_b200._types import OBJECT_DTYPE, RtcParams  # noqa: E402


def params_bytes(p):
    return np.frombuffer(bytes(p), np.uint8).copy()


def main():
    R = Reference()
    out = {}
    pi32 = np.float32(math.pi)
    # 1. the reference default scene, default camera, every mode, two console sizes
    for (x, y) in [(240, 64), (400, 150)]:
        p = R.camera_params(x, y, (0, 0, 0), (0, pi32, 0))
        out[f"default_{x}x{y}_params"] = params_bytes(p)
        for mode in range(6):
            r = R.update(None, p, mode, dt=0.0, want_raw=True, default_scene=True)
            out[f"default_{x}x{y}_m{mode}_stream"] = r["stream"]
            out[f"default_{x}x{y}_m{mode}_rawsha"] = np.frombuffer(hashlib.sha256(r["raw"].tobytes()).digest(), np.uint8).copy()
            if mode in (0, 2, 3, 4) and (x, y) == (240, 64):
                color, glyph, fg = planes_from_raw(r["raw"], x, y, mode)
                out[f"default_{x}x{y}_m{mode}_color"] = color
                out[f"default_{x}x{y}_m{mode}_glyph"] = glyph
                out[f"default_{x}x{y}_m{mode}_fg"] = fg
    # 2. random scenes (spheres + planes, shuffled order), random cameras, incl. physics step
    rng = np.random.default_rng(20261018)
    n_cases = 8
    out["n_cases"] = np.array([n_cases])
    for k in range(n_cases):
        n = int(rng.integers(3, 90))
        objs = scenes.random_spheres(n, 777 + k)
        extra = []
        for j in range(int(rng.integers(0, 3))):
            nrm = np.array([0.0, 1.0, 0.0]) if j == 0 else rng.normal(size=3)
            extra.append(scenes.make_plane(rng.uniform(-60, 60, 3), nrm, rng.uniform(0, 255, 3),
                                           float(rng.uniform(20, 400)), float(rng.uniform(20, 400))))
        if extra:
            objs = np.concatenate([objs, np.array(extra, OBJECT_DTYPE)])
        objs = objs[rng.permutation(len(objs))]
        x, y = int(rng.integers(30, 180)), int(rng.integers(12, 80))
        if k % 2 == 0:
            pos, rot = (0.0, 0.0, -120.0), (0.0, pi32, 0.0)
        else:
            pos, rot = rng.uniform(-120, 120, 3), (rng.uniform(-1.2, 1.2), rng.uniform(-3.1, 3.1), 0.0)
        if k == 5:
            pos = objs[0]["center"] + np.float32(0.25)      # camera inside / next to a sphere
        p = R.camera_params(x, y, pos, rot)
        if k % 2 == 0:                                        # bench-style pixel aspect 1/W
            p = rtc_b200.camera_params(x, y, pos, rot, 1.0 / (x - 1))
        dt = 0.0 if k % 3 else float(rng.uniform(0.0, 2.5))
        out[f"case{k}_objs"] = objs.view(np.uint8).reshape(-1).copy()
        out[f"case{k}_params"] = params_bytes(p)
        out[f"case{k}_dt"] = np.array([dt])
        for mode in range(5):
            r = R.update(objs, p, mode, dt=dt, want_raw=True, want_objs=(mode == 3))
            out[f"case{k}_m{mode}_stream"] = r["stream"]
            if mode == 3:
                out[f"case{k}_objs_after"] = r["objs"].view(np.uint8).reshape(-1).copy()
            if mode in (0, 2, 4):
                color, glyph, fg = planes_from_raw(r["raw"], x, y, mode)
                out[f"case{k}_m{mode}_color"] = color
                out[f"case{k}_m{mode}_glyph"] = glyph
                out[f"case{k}_m{mode}_fg"] = fg
    # 3. ansi256_from_rgb on a fixed pseudo-random sample + all greys
    sample = np.concatenate([rng.integers(0, 1 << 24, 20000, dtype=np.int64),
                             np.array([(v << 16) | (v << 8) | v for v in range(256)], np.int64)]).astype(np.uint32)
    allv = R.ansi256_range(0, 1 << 24)
    out["ansi_sample_rgb"] = sample
    out["ansi_sample_idx"] = allv[sample]
    out["ansi_all_sha"] = np.frombuffer(hashlib.sha256(allv.tobytes()).digest(), np.uint8).copy()
    # 4. per-function KATs
    L = R.L
    n_k = 400
    kat_o = rng.uniform(-80, 80, (n_k, 3)).astype(np.float32)
    kat_c = rng.uniform(-60, 60, (n_k, 3)).astype(np.float32)
    kat_r = rng.integers(0, 10, n_k).astype(np.float32)
    kat_d = rng.normal(size=(n_k, 3)).astype(np.float32)
    aim = (kat_c - kat_o) + rng.normal(scale=3.0, size=(n_k, 3)).astype(np.float32)
    kat_d[: n_k // 2] = aim[: n_k // 2]
    kat_d /= np.linalg.norm(kat_d, axis=1, keepdims=True).astype(np.float32)
    sp_hit = np.zeros(n_k, np.int32); sp_t = np.zeros(n_k, np.float32); sp_n = np.zeros((n_k, 3), np.float32)
    pl_hit = np.zeros(n_k, np.int32); pl_t = np.zeros(n_k, np.float32)
    for i in range(n_k):
        t = ctypes.c_float(); nn = np.zeros(3, np.float32)
        sp_hit[i] = L.ref_sphere_trace(kat_c[i].ctypes.data, float(kat_r[i]), kat_o[i].ctypes.data, kat_d[i].ctypes.data,
                                       ctypes.byref(t), nn.ctypes.data)
        sp_t[i] = t.value; sp_n[i] = nn
        nrm = np.array([0, 1, 0], np.float32) if i % 2 else rng.normal(size=3).astype(np.float32)
        t2 = ctypes.c_float(); n2 = np.zeros(3, np.float32)
        pl_hit[i] = L.ref_plane_trace(kat_c[i].ctypes.data, nrm.ctypes.data, 300.0, 200.0, kat_o[i].ctypes.data,
                                      kat_d[i].ctypes.data, ctypes.byref(t2), n2.ctypes.data)
        pl_t[i] = t2.value
        out.setdefault("kat_plane_n", np.zeros((n_k, 3), np.float32))[i] = nrm
    out.update(kat_o=kat_o, kat_c=kat_c, kat_r=kat_r, kat_d=kat_d, kat_sp_hit=sp_hit, kat_sp_t=sp_t, kat_sp_n=sp_n,
               kat_pl_hit=pl_hit, kat_pl_t=pl_t)
    # GetASCIICharacter
    sv = np.linspace(-1.0, 1.0, 401).astype(np.float32)
    out["kat_ascii_sv"] = sv
    out["kat_ascii_ch"] = np.array([L.ref_ascii_char(ctypes.c_float(10.0), ctypes.c_float(250.0), ctypes.c_float(float(v))) for v in sv], np.uint8)
    # camera blocks
    cams = []
    for i in range(64):
        pos = rng.uniform(-100, 100, 3).astype(np.float32); rot = rng.uniform(-3.2, 3.2, 3).astype(np.float32)
        cams.append(np.concatenate([pos.view(np.uint8), rot.view(np.uint8), params_bytes(R.camera_params(400, 150, pos, rot))]))
    out["kat_cameras"] = np.stack(cams)
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
