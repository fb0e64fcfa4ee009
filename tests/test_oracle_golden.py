"""CPU: the restated oracle (oracle/rt_oracle.c) against the committed golden vectors, which were
produced by the reference's own sources compiled for CPU (tests/golden/make_golden.py)."""
import ctypes
import hashlib

import numpy as np
import pytest

from rtc_b200._types import FLAG_UPDATE_REF_LAUNCH_LIMIT, MODE_NAMES, mode_cell
from util import objs_from_bytes, params_from_bytes, parse_stream


@pytest.mark.parametrize("size", [(240, 64), (400, 150)])
@pytest.mark.parametrize("mode", range(6))
def test_default_scene_frame(oracle, golden, size, mode):
    x, y = size
    p = params_from_bytes(golden[f"default_{x}x{y}_params"])
    objs = oracle.update_objects(oracle.default_scene(), 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
    raw = oracle.trace_raw(objs, p, mode)
    assert hashlib.sha256(raw.tobytes()).digest() == golden[f"default_{x}x{y}_m{mode}_rawsha"].tobytes(), \
        f"raw cell buffer differs from the reference in {MODE_NAMES[mode]}"
    stream = oracle.minimize(raw, x, y, mode)
    assert np.array_equal(stream, golden[f"default_{x}x{y}_m{mode}_stream"])
    # the local-rule encoder (what the GPU implements) must give the same bytes as the serial scan
    pl = oracle.trace_planes(objs, p, mode)
    stream2 = oracle.encode_planes(pl["color"], pl["glyph"], x, y, mode)
    assert np.array_equal(stream2, stream)


def test_default_scene_anchor_sizes(golden):
    """Sanity anchors the survey observed on the reference."""
    assert len(golden["default_400x150_m0_stream"]) == 62860
    assert len(golden["default_400x150_m3_stream"]) == 76682
    assert len(golden["default_400x150_m4_stream"]) == 301452
    assert len(golden["default_240x64_m3_stream"]) == 18666


def test_random_scenes(oracle, golden):
    for k in range(int(golden["n_cases"][0])):
        objs = objs_from_bytes(golden[f"case{k}_objs"])
        p = params_from_bytes(golden[f"case{k}_params"])
        dt = float(golden[f"case{k}_dt"][0])
        after = oracle.update_objects(objs, dt, FLAG_UPDATE_REF_LAUNCH_LIMIT)   # Update always runs the physics step
        assert after.tobytes() == golden[f"case{k}_objs_after"].tobytes(), f"case {k}: physics step differs"
        for mode in range(5):
            stream = oracle.render(after, p, mode)
            assert np.array_equal(stream, golden[f"case{k}_m{mode}_stream"]), f"case {k} mode {mode}"
            if mode in (0, 2, 4):
                pl = oracle.trace_planes(after, p, mode)
                assert np.array_equal(pl["color"], golden[f"case{k}_m{mode}_color"])
                if mode != 4:
                    assert np.array_equal(pl["glyph"], golden[f"case{k}_m{mode}_glyph"])
                    assert np.array_equal((pl["hit"] != 0), golden[f"case{k}_m{mode}_fg"] != 0)


def test_ansi256(oracle, golden):
    rgb = golden["ansi_sample_rgb"]
    allv = oracle.ansi256_range(0, 1 << 24)
    assert np.array_equal(allv[rgb], golden["ansi_sample_idx"])
    assert hashlib.sha256(allv.tobytes()).digest() == golden["ansi_all_sha"].tobytes()


def test_trace_kats(oracle, golden):
    L = oracle.L
    o, c, r, d = golden["kat_o"], golden["kat_c"], golden["kat_r"], golden["kat_d"]
    for i in range(len(o)):
        t = ctypes.c_float(); n = np.zeros(3, np.float32)
        oi, ci, di = o[i].copy(), c[i].copy(), d[i].copy()
        hit = L.orc_sphere_trace(ci.ctypes.data, float(r[i]), oi.ctypes.data, di.ctypes.data, ctypes.byref(t), n.ctypes.data)
        assert hit == golden["kat_sp_hit"][i]
        if hit:
            assert np.float32(t.value).tobytes() == golden["kat_sp_t"][i].tobytes()
            assert n.tobytes() == golden["kat_sp_n"][i].tobytes()
        pn = golden["kat_plane_n"][i].copy(); t2 = ctypes.c_float(); n2 = np.zeros(3, np.float32)
        hit2 = L.orc_plane_trace(ci.ctypes.data, pn.ctypes.data, 300.0, 200.0, oi.ctypes.data, di.ctypes.data,
                                 ctypes.byref(t2), n2.ctypes.data)
        assert hit2 == golden["kat_pl_hit"][i]
        if hit2:
            assert np.float32(t2.value).tobytes() == golden["kat_pl_t"][i].tobytes()
    assert golden["kat_sp_hit"].sum() > 20 and golden["kat_pl_hit"].sum() > 20   # the KATs exercise both outcomes


def test_ascii_ramp(oracle, golden):
    got = np.array([oracle.L.orc_ascii_char(10.0, 250.0, float(v)) for v in golden["kat_ascii_sv"]], np.uint8)
    assert np.array_equal(got, golden["kat_ascii_ch"])
    assert oracle.L.orc_ascii_char(300.0, 250.0, 0.5) == ord(" ")          # beyond the far plane


def test_camera_blocks(oracle, golden, rtc):
    for row in golden["kat_cameras"]:
        pos = np.frombuffer(row[:12].tobytes(), np.float32); rot = np.frombuffer(row[12:24].tobytes(), np.float32)
        want = row[24:].tobytes()
        assert bytes(oracle.camera_params(400, 150, pos, rot)) == want
        assert bytes(rtc.camera_params(400, 150, pos, rot)) == want          # the library's host code (no GPU)


def test_digits_and_cells(oracle):
    buf = ctypes.create_string_buffer(3)
    for v in range(256):
        oracle.L.orc_digits3(v, buf)
        s = str(v)
        want = b"\0" * (3 - len(s)) + s.encode()
        assert buf.raw == want                                              # NUL padding, not '0' or ' '
    cell = ctypes.create_string_buffer(20)
    c = (ctypes.c_uint8 * 3)(7, 45, 255)
    assert oracle.L.orc_make_cell(3, c, ord(" "), 1, cell) == 20
    assert cell.raw == b"\x1b[48;2;\0\0007;\00045;255m "
    assert oracle.L.orc_make_cell(2, c, ord("#"), 1, cell) == 20
    assert cell.raw == b"\x1b[38;2;\0\0007;\00045;255m#"
    assert oracle.L.orc_make_cell(0, c, ord("#"), 1, cell) == 12
    assert cell.raw[:12] == b"\x1b[38;5;\0\0007m#"


def test_minimiser_properties(oracle):
    """Hand-built planes: run structure, row carry-over, first cell, ragged sizes."""
    rng = np.random.default_rng(5)
    for (x, y) in [(2, 1), (2, 5), (3, 3), (17, 4), (65, 7), (130, 3)]:
        W = x - 1
        for mode in (3, 1, 2, 0):
            bpp = 1 if mode in (0, 1) else 3
            keys = rng.integers(0, 3, (y * W, 1)).astype(np.uint8).repeat(bpp, 1) * 90   # long runs
            glyph = None
            if mode in (0, 2):
                glyph = rng.choice(np.frombuffer(b" .#@", np.uint8), y * W)
            stream = oracle.encode_planes(keys.reshape(-1), glyph, x, y, mode)
            k2, g2, full = parse_stream(stream, x, y, mode)
            assert np.array_equal(k2, keys.reshape(-1))
            if glyph is not None:
                assert np.array_equal(g2, glyph)
            kk = keys.reshape(y * W, bpp)
            want_full = np.ones(y * W, np.uint8)
            want_full[1:] = (kk[1:] != kk[:-1]).any(1)
            assert np.array_equal(full, want_full)                          # colour carried across rows
            assert len(stream) == int(want_full.sum()) * mode_cell(mode) + int((1 - want_full).sum()) + y


def test_update_objects_launch_limit(oracle):
    from rtc_b200 import scenes
    objs = scenes.random_spheres(1025, 9)
    same = oracle.update_objects(objs, 0.5, FLAG_UPDATE_REF_LAUNCH_LIMIT)
    assert same.tobytes() == objs.tobytes()                 # reference: block=count>1024 -> launch rejected, nothing moves
    moved = oracle.update_objects(objs, 0.5, 0)
    assert moved.tobytes() != objs.tobytes()
    assert np.all(np.abs(moved["center"][:, 1]) <= 10.0)    # Sphere::Update clamps y into [-10,10]


def test_encode_planes_equals_raw_minimiser_property(oracle):
    """Property (hypothesis): for arbitrary colour / glyph / hit planes, the plane-based encoder the GPU is compared
    against (orc_encode_planes: "emit the full cell iff the colour key differs from the previous traced cell") produces
    exactly the bytes of the reference's own pipeline -- 20/12-byte cells laid out with row stride SIZE*x in the
    20*x*y raw buffer (RayTracing.cu:585-608, :231-251), then the serial MinimizeRGB / Minimize8bit scan
    (RayTracingManager.cu:251-319, :181-249) -- including its quirks: NUL-padded digits are bytes like any other, the
    fg/bg selector and the glyph do not take part in the comparison, latestColor survives row ends."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None)
    @given(st.integers(2, 40), st.integers(1, 12), st.sampled_from([0, 1, 2, 3]), st.integers(0, 2 ** 32 - 1),
           st.sampled_from([1, 2, 3, 255]))
    def check(x, y, mode, seed, n_levels):
        rng = np.random.default_rng(seed)
        W = x - 1
        bpp = 1 if mode in (0, 1) else 3
        cs = mode_cell(mode)
        has_glyph = mode in (0, 2)
        levels = rng.integers(0, 256, n_levels).astype(np.uint8)
        keys = levels[rng.integers(0, n_levels, (y * W, bpp))]
        if has_glyph:                                              # ' ' == a miss (background selector), others == a hit
            glyph = rng.choice(np.frombuffer(b"  .#@", np.uint8), y * W)
        else:
            glyph = None
        raw = np.zeros(20 * x * y, np.uint8)
        cell = ctypes.create_string_buffer(20)
        for r in range(y):
            for c in range(W):
                i = r * W + c
                g = int(glyph[i]) if has_glyph else 32
                hit = 1 if (not has_glyph or g != 32) else 0
                k = keys[i]
                if has_glyph and not hit:                           # the reference's miss cell is black / index 16
                    k = np.array([16], np.uint8) if bpp == 1 else np.zeros(3, np.uint8)
                    keys[i] = k
                col = (ctypes.c_uint8 * 3)(*[int(v) for v in (list(k) + [0, 0])[:3]])
                n = oracle.L.orc_make_cell(mode, col, g, hit, cell)
                assert n == cs
                off = r * x * cs + c * cs
                raw[off:off + cs] = np.frombuffer(cell.raw[:cs], np.uint8)
        want = oracle.minimize(raw, x, y, mode)
        got = oracle.encode_planes(keys.reshape(-1), glyph, x, y, mode)
        assert np.array_equal(got, want)

    check()


def test_unit_rsqrt_formula():
    """csrc/rtc_device.cuh: rsqrt_near_one_exact replaces fl(1 / fl(sqrt(s))) for s within 128 ulp of 1 (the reference
    re-normalises already-normalised vectors: Sphere.cu:67 -> RayTracing.cu:129 -> :56, :150 -> :57) by integer
    arithmetic on the bit pattern.  The two IEEE operations (numpy float32 sqrt and divide are correctly rounded) and the
    formula must agree on every such s."""
    one = 0x3F800000
    for m in range(-128, 129):
        s = np.array([one + m], np.int32).view(np.float32)[0]
        inv = np.float32(1.0) / np.sqrt(s, dtype=np.float32)
        want = int(np.array([inv], np.float32).view(np.int32)[0])
        if m >= 0:
            got = one - (m & ~1)
        else:
            k = -m
            got = one + ((((k + 1) >> 1) + 1) >> 1)
        assert got == want, (m, hex(want), hex(got))
