"""CPU, build container only: pin the restated oracle against the reference's OWN sources
(oracle/_ref/libref_cpu.so, compiled unmodified from /root/reference by oracle/ref_build)."""
import numpy as np

from rtc_b200 import scenes
from rtc_b200._types import FLAG_UPDATE_REF_LAUNCH_LIMIT, OBJECT_DTYPE
from util import PI32


def test_ansi256_exhaustive(oracle, reference):
    assert np.array_equal(oracle.ansi256_range(0, 1 << 24), reference.ansi256_range(0, 1 << 24))


def test_abi_sizes(reference):
    import ctypes
    sz = (ctypes.c_int * 8)()
    reference.L.ref_sizes(sz)
    assert list(sz)[:5] == [24, 64, 96, 96, 160]


def test_frames_random(oracle, reference):
    rng = np.random.default_rng(42)
    for trial in range(120):
        n = int(rng.integers(1, 70))
        objs = scenes.random_spheres(n, 4242 + trial)
        if trial % 2:
            pl = scenes.make_plane(rng.uniform(-60, 60, 3), rng.normal(size=3), rng.uniform(0, 255, 3), 300.0, 250.0)
            objs = np.concatenate([objs, np.array([pl], OBJECT_DTYPE)])[rng.permutation(n + 1)]
        x, y = int(rng.integers(20, 140)), int(rng.integers(10, 60))
        p = oracle.camera_params(x, y, rng.uniform(-120, 120, 3), (rng.uniform(-1, 1), rng.uniform(-3, 3), 0.0),
                                 0.0 if trial % 2 else 1.0 / (x - 1))
        dt = float(rng.uniform(0, 2)) if trial % 3 == 0 else 0.0
        after = oracle.update_objects(objs, dt, FLAG_UPDATE_REF_LAUNCH_LIMIT)
        for mode in range(6):
            r = reference.update(objs, p, mode, dt=dt, want_raw=True, want_objs=True)
            raw = oracle.trace_raw(after, p, mode)
            assert np.array_equal(raw, r["raw"]), (trial, mode)
            assert np.array_equal(oracle.minimize(raw, x, y, mode), r["stream"]), (trial, mode)
            assert r["objs"].tobytes() == after.tobytes()


def test_camera(oracle, reference, rtc):
    rng = np.random.default_rng(3)
    for i in range(500):
        pos = rng.uniform(-100, 100, 3).astype(np.float32)
        rot = rng.uniform(-3.2, 3.2, 3).astype(np.float32) if i else np.array([0, PI32, 0], np.float32)
        want = bytes(reference.camera_params(400, 150, pos, rot))
        assert bytes(oracle.camera_params(400, 150, pos, rot)) == want
        assert bytes(rtc.camera_params(400, 150, pos, rot)) == want


def test_more_than_1024_objects_do_not_move(oracle, reference):
    """The reference's UpdateObjects launch is invalid for count > 1024."""
    objs = scenes.random_spheres(1030, 11)
    p = oracle.camera_params(40, 12, (0, 0, -120), (0, PI32, 0), 1.0 / 39)
    r = reference.update(objs, p, 3, dt=1.0, want_objs=True)
    assert r["objs"].tobytes() == objs.tobytes()
    assert np.array_equal(oracle.render(objs, p, 3), r["stream"])
