"""GPU (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle and the golden
vectors made by the reference itself.

Bars (BASELINE.json north_star):
  * hit/miss decisions, accepted object index and hit distance: IDENTICAL (bit-exact);
  * quantised colour: within +-1 LSB per channel (the only non-bit-exact operation is pow(x,32),
    computed correctly rounded on the GPU vs the host libm in the oracle); in practice identical;
  * ANSI stream: bit-exact given identical colour planes.
"""
import numpy as np
import pytest

from rtc_b200 import scenes
from rtc_b200._types import (BIT_ASCII, BIT_PIXEL, FLAG_CULL, FLAG_KEEP_HITS, FLAG_PACKET, FLAG_SHADOWS, FLAG_UPDATE_REF_LAUNCH_LIMIT, MODE_NAMES,
                             OBJECT_DTYPE, RGB_ASCII, RGB_NORMALS, RGB_PIXEL, SDL, mode_bpp, mode_cell, mode_has_glyph)
from util import PI32, bind_stream, objs_from_bytes, params_from_bytes, unbind_stream

pytestmark = pytest.mark.gpu

RGB_TOL = 1          # LSB per channel, stated tolerance for the floating-point part of the path
MAX_OFF_FRACTION = 1e-4


def check_frame(ctx, oracle, objs, p, mode, flags=0, expect_stream=None):
    x, y = p.x, p.y
    n_px = (x - 1) * y
    ctx.set_objects(objs)
    # the product path: the ray kernel shades in its tile epilogue, hit records never leave the SM
    ctx.render(p, mode, flags)
    stream = ctx.frame_ansi()
    planes = ctx.frame_color(n_px) if mode != SDL else None
    # the same frame with the parity hook on: hit records kept; planes and stream must not change
    ctx.render(p, mode, flags | FLAG_KEEP_HITS)
    assert np.array_equal(ctx.frame_ansi(), stream), "RTC_FLAG_KEEP_HITS changed the stream"
    if mode != SDL:
        c2, g2 = ctx.frame_color(n_px)
        assert np.array_equal(c2, planes[0]) and (g2 is None or np.array_equal(g2, planes[1])), "RTC_FLAG_KEEP_HITS changed the planes"
    dist, index = ctx.frame_hits(n_px)
    o = oracle.trace_planes(objs, p, mode, flags)
    # --- hit records: bit-exact
    assert np.array_equal(index, o["index"]), f"{MODE_NAMES[mode]}: accepted object differs on {(index != o['index']).sum()} rays"
    assert dist.tobytes() == o["dist"].tobytes(), f"{MODE_NAMES[mode]}: hit distance differs"
    if mode == SDL:
        assert stream.tobytes() == b"\n" * y
        return stream
    color, glyph = ctx.frame_color(n_px)
    # --- colour planes: +-1 LSB on at most a sliver of pixels (expected: identical)
    if mode in (BIT_ASCII, BIT_PIXEL):
        n_off = int((color != o["color"]).sum())     # an xterm index may flip when RGB moves by 1 LSB
    else:
        diff = np.abs(color.astype(np.int16) - o["color"].astype(np.int16))
        assert diff.max(initial=0) <= RGB_TOL, f"{MODE_NAMES[mode]}: colour off by {diff.max()} LSB"
        n_off = int((diff != 0).sum())
    assert n_off <= max(2, MAX_OFF_FRACTION * n_px), f"{MODE_NAMES[mode]}: {n_off} colour bytes differ from the oracle"
    if mode_has_glyph(mode):
        assert np.array_equal(glyph, o["glyph"])
    # --- stream: bit-exact given the GPU's own planes ...
    assert np.array_equal(stream, oracle.encode_planes(color, glyph, x, y, mode)), "ANSI stream differs given identical planes"
    # ... and, when the planes are identical (the expected case), equal to the oracle's full path
    if n_off == 0:
        want = oracle.render(objs, p, mode, flags) if expect_stream is None else expect_stream
        assert np.array_equal(stream, want), "ANSI stream differs from the reference path"
    return stream


@pytest.mark.parametrize("size", [(240, 64), (400, 150)])
@pytest.mark.parametrize("mode", range(6))
def test_default_scene_vs_golden(ctx, oracle, golden, size, mode):
    x, y = size
    p = params_from_bytes(golden[f"default_{x}x{y}_params"])
    objs = scenes.default_scene()
    ctx.set_objects(objs)
    stream = np.array(ctx.update(p, mode, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))   # == RayTracingManager::Update
    assert np.array_equal(stream, golden[f"default_{x}x{y}_m{mode}_stream"]), \
        f"{MODE_NAMES[mode]} {x}x{y}: stream differs from the reference's"
    after = ctx.get_objects()
    check_frame(ctx, oracle, after, p, mode, expect_stream=golden[f"default_{x}x{y}_m{mode}_stream"])


def test_random_scenes_vs_golden(ctx, oracle, golden):
    for k in range(int(golden["n_cases"][0])):
        objs = objs_from_bytes(golden[f"case{k}_objs"])
        p = params_from_bytes(golden[f"case{k}_params"])
        dt = float(golden[f"case{k}_dt"][0])
        for mode in range(5):
            ctx.set_objects(objs)
            stream = np.array(ctx.update(p, mode, dt=dt, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))
            assert ctx.get_objects().tobytes() == golden[f"case{k}_objs_after"].tobytes(), f"case {k}: physics differs"
            assert np.array_equal(stream, golden[f"case{k}_m{mode}_stream"]), f"case {k} {MODE_NAMES[mode]}"
            if mode in (0, 2, 4):
                color, glyph = ctx.frame_color((p.x - 1) * p.y)
                assert np.array_equal(color, golden[f"case{k}_m{mode}_color"])
                if mode != 4:
                    assert np.array_equal(glyph, golden[f"case{k}_m{mode}_glyph"])


@pytest.mark.parametrize("name", ["config2_1080p_64"])
def test_bench_config_small(ctx, oracle, name):
    """Config 2 (1921x1080, 64 spheres + plane) against the oracle, whole frame."""
    objs = scenes.config_scene(name)
    p = scenes.config_camera(name)
    check_frame(ctx, oracle, objs, p, RGB_PIXEL)


def test_many_spheres_band(ctx, oracle):
    """Config-3 scene (1024 spheres + plane): a band of rows against the oracle (full frame is too slow on CPU)."""
    objs = scenes.config_scene("config3_4k_1024")
    p = scenes.config_camera("config3_4k_1024")
    ctx.set_objects(objs)
    n_px = (p.x - 1) * p.y
    ctx.render(p, RGB_PIXEL)
    color0, _ = ctx.frame_color(n_px)
    with pytest.raises(Exception, match="KEEP_HITS"):
        ctx.frame_hits(n_px)
    ctx.render(p, RGB_PIXEL, FLAG_KEEP_HITS)
    dist, index = ctx.frame_hits(n_px)
    color, _ = ctx.frame_color(n_px)
    assert np.array_equal(color, color0)
    W = p.x - 1
    for (r0, r1) in [(0, 4), (1078, 1084), (2150, 2160)]:
        o = oracle.trace_planes(objs, p, RGB_PIXEL, row0=r0, row1=r1)
        sl = slice(r0 * W, r1 * W)
        assert np.array_equal(index[sl], o["index"])
        assert dist[sl].tobytes() == o["dist"].tobytes()
        d = np.abs(color[r0 * W * 3:r1 * W * 3].astype(np.int16) - o["color"].astype(np.int16))
        assert d.max() <= RGB_TOL and (d != 0).sum() <= 2
    # size-independent properties of the full-size stream
    stream = ctx.frame_ansi()
    assert stream[-1] == 10 and int((stream == 10).sum()) >= p.y
    assert np.array_equal(stream, oracle.encode_planes(color, None, p.x, p.y, RGB_PIXEL))


def test_trace_band_matches_full_frame(ctx, rtc):
    """Row-band partition (multi-GPU building block): bands written at their offsets == the full frame."""
    import torch
    objs = scenes.config_scene("config2_1080p_64")
    x, y = 481, 270
    p = rtc.camera_params(x, y, (0, 0, -120), (0, PI32, 0), 1.0 / (x - 1))
    W = x - 1
    ctx.set_objects(objs)
    for mode in (RGB_PIXEL, RGB_ASCII, BIT_ASCII):
        ctx.render(p, mode)
        full_color, full_glyph = ctx.frame_color(W * y)
        want = ctx.frame_ansi()
        bpp = mode_bpp(mode)
        color = torch.zeros(W * y * bpp, dtype=torch.uint8, device="cuda")
        glyph = torch.zeros(W * y, dtype=torch.uint8, device="cuda")
        bind_stream(ctx)
        for (r0, r1) in [(0, 67), (67, 135), (135, 136), (136, 270)]:      # ragged bands
            ctx.trace_band(p, mode, r0, r1, color.data_ptr() + r0 * W * bpp, glyph.data_ptr() + r0 * W)
        cap = rtc.encode_capacity(x, y, mode)
        out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.encode(color.data_ptr(), glyph.data_ptr() if mode_has_glyph(mode) else 0, x, y, mode, out.data_ptr(), cap, total.data_ptr())
        torch.cuda.synchronize()
        unbind_stream(ctx)
        assert np.array_equal(color.cpu().numpy(), full_color)
        n = int(total.item())
        assert np.array_equal(out[:n].cpu().numpy(), want)


def test_encoder_edge_cases(ctx, oracle, rtc):
    """Encoder alone on hand-made planes: ragged sizes, 1-cell rows, long runs, every-cell-differs,
    slice/tile boundaries (160-cell warp slices, 1280-cell tiles), unaligned plane pointers."""
    import torch
    rng = np.random.default_rng(11)
    bind_stream(ctx)
    cases = [(2, 1), (2, 9), (3, 3), (5, 2), (6, 7), (18, 5), (161, 1), (162, 3), (321, 7), (1281, 1), (1282, 2), (2049, 1),
             (2562, 2), (4097, 3), (700, 37), (1025, 16)]
    for (x, y) in cases:
        W = x - 1
        for mode in (RGB_PIXEL, RGB_ASCII, BIT_PIXEL, BIT_ASCII):
            bpp = mode_bpp(mode)
            for pattern in ("noise", "runs", "constant"):
                if pattern == "noise":
                    keys = rng.integers(0, 256, W * y * bpp).astype(np.uint8)
                elif pattern == "runs":
                    keys = np.repeat(rng.integers(0, 256, (W * y + 36) // 37 * bpp).astype(np.uint8).reshape(-1, bpp), 37, 0)[:W * y].reshape(-1)
                else:
                    keys = np.full(W * y * bpp, 200, np.uint8)
                glyph = rng.choice(np.frombuffer(b"  .:#@", np.uint8), W * y) if mode_has_glyph(mode) else None
                for off in (0, 5):                                  # 5: deliberately unaligned device pointers
                    dk = torch.zeros(keys.size + 16, dtype=torch.uint8, device="cuda")
                    dk[off:off + keys.size] = torch.from_numpy(keys).cuda()
                    dg = None
                    if glyph is not None:
                        dg = torch.zeros(glyph.size + 16, dtype=torch.uint8, device="cuda")
                        dg[off:off + glyph.size] = torch.from_numpy(glyph).cuda()
                    cap = rtc.encode_capacity(x, y, mode)
                    out = torch.zeros(cap + 16, dtype=torch.uint8, device="cuda")
                    total = torch.zeros(1, dtype=torch.int64, device="cuda")
                    ctx.encode(dk.data_ptr() + off, (dg.data_ptr() + off) if dg is not None else 0, x, y, mode,
                               out.data_ptr() + (off % 3), cap, total.data_ptr())
                    torch.cuda.synchronize()
                    n = int(total.item())
                    got = out[(off % 3):(off % 3) + n].cpu().numpy()
                    want = oracle.encode_planes(keys, glyph, x, y, mode)
                    assert np.array_equal(got, want), (x, y, MODE_NAMES[mode], pattern, off)
    unbind_stream(ctx)


def test_encoder_full_size_properties(ctx, oracle, rtc):
    """Config 5 (7681x4320) worst case: i.i.d. random RGB -> almost every cell emits 20 bytes.
    Size-independent checks: length formula, newline count/positions, random windows decoded."""
    import torch
    x, y = 7681, 4320
    W = x - 1
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    rgb = torch.randint(0, 256, (W * y * 3,), dtype=torch.uint8, device="cuda", generator=g)
    cap = rtc.encode_capacity(x, y, RGB_PIXEL)
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    bind_stream(ctx)
    ctx.encode(rgb.data_ptr(), 0, x, y, RGB_PIXEL, out.data_ptr(), cap, total.data_ptr())
    torch.cuda.synchronize()
    unbind_stream(ctx)
    n = int(total.item())
    px = rgb.view(-1, 3)
    full = torch.ones(W * y, dtype=torch.bool, device="cuda")
    full[1:] = (px[1:] != px[:-1]).any(1)
    n_full = int(full.sum().item())
    assert n == 20 * n_full + (W * y - n_full) + y                          # length formula
    lens = torch.where(full, 20, 1).to(torch.int64)
    lens.view(y, W)[:, -1] += 1
    ends = torch.cumsum(lens, 0)
    row_end = ends.view(y, W)[:, -1] - 1
    assert bool((out[row_end] == 10).all())                                 # one newline per row, in place
    # decode random windows and the first/last rows against the oracle on the same cells
    host = out[:n].cpu().numpy()
    starts = (ends - lens).cpu().numpy()
    rgb_h = rgb.cpu().numpy()
    for row in [0, 1, y // 2, y - 1]:
        a, b = int(starts[row * W]), int(ends[(row + 1) * W - 1])
        want = oracle.encode_planes(rgb_h[(row * W - (1 if row else 0)) * 3:(row + 1) * W * 3], None,
                                    W + (2 if row else 1), 1, RGB_PIXEL)
        if row:                                                             # drop the predecessor cell used as context
            want = want[20:]
        assert np.array_equal(host[a:b], want), f"row {row}"


def test_update_objects(ctx, oracle):
    objs = scenes.random_spheres(300, 21)
    ctx.set_objects(objs)
    cur = objs
    for dt in (0.0, 0.37, 1.9, 7.0):
        ctx.update_objects(dt)
        cur = oracle.update_objects(cur, dt)
        assert ctx.get_objects().tobytes() == cur.tobytes()
    big = scenes.random_spheres(1500, 22)
    ctx.set_objects(big)
    ctx.update_objects(0.5, FLAG_UPDATE_REF_LAUNCH_LIMIT)                   # reference launch bug mirrored
    assert ctx.get_objects().tobytes() == big.tobytes()
    ctx.update_objects(0.5)                                                 # fixed behaviour
    assert ctx.get_objects().tobytes() == oracle.update_objects(big, 0.5).tobytes()


def test_scene_api(ctx, oracle, rtc):
    """rtc_scene_add_* (Scene3D::CreateSphere/CreatePlane) == bulk upload."""
    objs = scenes.default_scene()
    ctx.clear()
    for o in objs:
        if o["type"] == 2:
            ctx.add_sphere(o["center"], o["radius"], o["color"], o["speed"], o["mover"])
        else:
            ctx.add_plane(o["center"], (0.0, 3.0, 0.0), o["color"], o["width"], o["height"])   # un-normalised normal
    assert ctx.get_objects().tobytes() == objs.tobytes()
    p = rtc.camera_params(120, 40, (0, 0, 0), (0, PI32, 0))
    check_frame(ctx, oracle, objs, p, RGB_ASCII)


def test_edge_scenes(ctx, oracle, rtc):
    p = rtc.camera_params(64, 20, (0, 0, 0), (0, PI32, 0))
    empty = np.zeros(0, OBJECT_DTYPE)
    for mode in range(6):
        check_frame(ctx, oracle, empty, p, mode)                            # empty scene: all misses
    one = np.array([scenes.make_sphere((0, 0, 30), 0.0, (10, 20, 30))], OBJECT_DTYPE)   # zero radius
    check_frame(ctx, oracle, one, p, RGB_PIXEL)
    inside = np.array([scenes.make_sphere((0, 0, 1), 50.0, (200, 20, 30))], OBJECT_DTYPE)   # camera inside: invisible (Sphere.cu:57)
    s = check_frame(ctx, oracle, inside, p, RGB_PIXEL)
    assert len(s) == 20 + (63 * 20 - 1) + 20                                # one escape, then characters only
    dup = np.array([scenes.make_sphere((0, 0, 30), 5.0, (10, 200, 30)), scenes.make_sphere((0, 0, 30), 5.0, (200, 10, 30))],
                   OBJECT_DTYPE)                                            # exact tie: lowest index wins
    check_frame(ctx, oracle, dup, p, RGB_PIXEL)
    planes = np.array([scenes.make_plane((0, -5, 30), (0, 1, 0), (50, 60, 70), 80, 80),
                       scenes.make_plane((0, -5, 30), (0, 1, 0), (250, 60, 70), 80, 80)], OBJECT_DTYPE)
    check_frame(ctx, oracle, planes, p, RGB_ASCII)
    tiny = rtc.camera_params(2, 1, (0, 0, 0), (0, PI32, 0))                 # 1 traced cell
    check_frame(ctx, oracle, scenes.default_scene(), tiny, RGB_PIXEL)
    ragged = rtc.camera_params(38, 23, (0, 3, -10), (0.2, PI32, 0))         # not a multiple of the 16x16 warp tile
    for mode in range(5):
        check_frame(ctx, oracle, scenes.default_scene(), ragged, mode)


def test_shadow_extension(ctx, oracle, rtc):
    """Opt-in extension (not in the reference): shadow rays, against the oracle's definition."""
    objs = np.concatenate([scenes.default_scene(),
                           np.array([scenes.make_plane((0, -3, 30), (0, 1, 0), (100, 100, 100), 200, 200)], OBJECT_DTYPE)])
    p = rtc.camera_params(160, 60, (0, 5, -10), (0.1, PI32, 0))
    s_on = check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_SHADOWS)
    s_off = check_frame(ctx, oracle, objs, p, RGB_PIXEL)
    assert not np.array_equal(s_on, s_off)


def test_many_spheres_chunked(ctx, oracle, rtc):
    """More spheres than one shared-memory chunk (3032): multi-launch carry of the running best (and, for the
    shadow pass, of the occlusion mask)."""
    objs = scenes.random_spheres(9000, 31)
    p = rtc.camera_params(49, 20, (0, 0, -120), (0, PI32, 0), 1.0 / 48)
    check_frame(ctx, oracle, objs, p, RGB_PIXEL)
    check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_SHADOWS)


def test_reference_scene_capacity(ctx, oracle, rtc):
    """The reference's typed sphere array holds 52,083 spheres (5 MB / 96 B, Scene3D.cpp:131-141): a scene of that size
    goes through ~21 shared-memory chunks per pass."""
    objs = scenes.random_spheres(52083, 41)
    p = rtc.camera_params(34, 9, (0, 0, -120), (0, PI32, 0), 1.0 / 33)
    check_frame(ctx, oracle, objs, p, RGB_PIXEL)
    check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_CULL | FLAG_SHADOWS)


def test_shadow_pass_config2(ctx, oracle):
    """BASELINE config 2 (1921x1080, 64 spheres + plane, primary + shadow rays): the light-origin shadow pass
    (trace_kernel<true>: packed filter + exact path) against the oracle's definition, whole frame, two modes."""
    objs = scenes.config_scene("config2_1080p_64")
    p = scenes.config_camera("config2_1080p_64")
    s_on = check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_SHADOWS)
    s_off = check_frame(ctx, oracle, objs, p, RGB_PIXEL)
    assert not np.array_equal(s_on, s_off)
    check_frame(ctx, oracle, objs, p, BIT_ASCII, flags=FLAG_SHADOWS)


def test_pipelined_submit_collect(ctx, rtc):
    """rtc_submit / rtc_collect (two frames in flight, stream copied out on the copy stream while the next
    frame renders) == the synchronous rtc_update, frame by frame, with physics and a moving camera."""
    objs = scenes.random_spheres(200, 5)
    ps = [rtc.camera_params(321, 100, (0.7 * k, 0, -120), (0, PI32, 0), 1.0 / 320) for k in range(6)]
    modes = [RGB_PIXEL, RGB_ASCII, BIT_PIXEL, RGB_PIXEL, BIT_ASCII, RGB_PIXEL]
    ctx.set_objects(objs)
    want = [np.array(ctx.update(p, m, dt=0.11)) for p, m in zip(ps, modes)]
    ctx.set_objects(objs)
    got = []
    ctx.submit(ps[0], modes[0], dt=0.11)
    for k in range(len(ps)):
        if k + 1 < len(ps):
            ctx.submit(ps[k + 1], modes[k + 1], dt=0.11)
        got.append(ctx.collect(copy=True))
    for k in range(len(ps)):
        assert np.array_equal(got[k], want[k]), f"frame {k}"
    with pytest.raises(rtc.RtcError, match="no frame in flight"):
        ctx.collect()
    ctx.submit(ps[0], RGB_PIXEL)
    ctx.submit(ps[1], RGB_PIXEL)
    with pytest.raises(rtc.RtcError, match="in flight"):
        ctx.submit(ps[2], RGB_PIXEL)
    ctx.collect(); ctx.collect()


def test_band_encode_concatenates(ctx, rtc):
    """Per-band encoding (multi-GPU without gathering planes): every band is traced with one context row above it and
    encoded with rtc_encode_band(continues=1); the band streams concatenate to the whole frame's stream, bit for bit."""
    import torch
    objs = scenes.config_scene("config2_1080p_64")
    x, y = 481, 270
    p = rtc.camera_params(x, y, (0, 0, -120), (0, PI32, 0), 1.0 / (x - 1))
    W = x - 1
    ctx.set_objects(objs)
    bind_stream(ctx)
    for mode in (RGB_PIXEL, RGB_ASCII, BIT_ASCII, BIT_PIXEL):
        ctx.render(p, mode)
        want = ctx.frame_ansi()
        bpp, gl = mode_bpp(mode), mode_has_glyph(mode)
        pieces = []
        for (r0, r1) in [(0, 67), (67, 135), (135, 136), (136, 136), (136, 270)]:       # ragged, 1-row and empty bands
            c0 = r0 - 1 if r0 > 0 else 0                                              # context row
            color = torch.zeros((r1 - c0) * W * bpp + 64, dtype=torch.uint8, device="cuda")
            glyph = torch.zeros((r1 - c0) * W + 64, dtype=torch.uint8, device="cuda")
            ctx.trace_band(p, mode, c0, r1, color.data_ptr(), glyph.data_ptr() if gl else 0)
            skip = (r0 - c0) * W
            cap = rtc.encode_capacity(x, max(1, r1 - r0), mode)
            out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
            total = torch.zeros(1, dtype=torch.int64, device="cuda")
            ctx.encode_band(color.data_ptr() + skip * bpp, (glyph.data_ptr() + skip) if gl else 0, x, r1 - r0, mode,
                            r0 > 0, out.data_ptr(), cap, total.data_ptr())
            torch.cuda.synchronize()
            pieces.append(out[:int(total.item())].cpu().numpy())
        assert np.array_equal(np.concatenate(pieces), want), MODE_NAMES[mode]
    unbind_stream(ctx)


def test_culling_is_invisible(ctx, oracle, rtc):
    """RTC_FLAG_CULL (per-tile sphere-group culling) must not change a single hit record, colour or stream byte:
    whole frames against the oracle with the flag on, primary and shadow passes, narrow and wide tile cones, chunked
    sphere lists, camera inside the cloud; and it must actually skip work."""
    objs = scenes.config_scene("config2_1080p_64")
    p = scenes.config_camera("config2_1080p_64")
    check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_CULL)
    check_frame(ctx, oracle, objs, p, RGB_ASCII, flags=FLAG_CULL | FLAG_SHADOWS)
    # wide cones: tiny consoles take the keep-all path
    small = rtc.camera_params(64, 20, (0, 0, 0), (0, PI32, 0))
    check_frame(ctx, oracle, scenes.default_scene(), small, RGB_PIXEL, flags=FLAG_CULL)
    check_frame(ctx, oracle, scenes.default_scene(), small, BIT_ASCII, flags=FLAG_CULL | FLAG_SHADOWS)
    # many spheres: several shared-memory chunks, ragged last group, camera inside the cloud and inside a sphere
    many = scenes.random_spheres(9001, 33)
    q = rtc.camera_params(241, 120, (0, 0, -120), (0, PI32, 0), 1.0 / 240)
    check_frame(ctx, oracle, many, q, RGB_PIXEL, flags=FLAG_CULL)
    inside = rtc.camera_params(161, 90, (3, -2, 5), (0.3, 1.0, 0), 1.0 / 160)
    check_frame(ctx, oracle, scenes.random_spheres(700, 34), inside, RGB_PIXEL, flags=FLAG_CULL)
    check_frame(ctx, oracle, scenes.random_spheres(700, 34), inside, RGB_PIXEL, flags=FLAG_CULL | FLAG_SHADOWS)
    # full-size config 3: bands against the oracle, whole frame against the brute-force GPU path, and the work saved
    objs = scenes.config_scene("config3_4k_1024")
    p = scenes.config_camera("config3_4k_1024")
    n_px = (p.x - 1) * p.y
    ctx.set_objects(objs)
    ctx.render(p, RGB_PIXEL, FLAG_KEEP_HITS)
    d0, i0 = ctx.frame_hits(n_px)
    s0 = ctx.frame_ansi()
    brute = ctx.timings()["sphere_tests"]
    ctx.render(p, RGB_PIXEL, FLAG_CULL | FLAG_KEEP_HITS)
    d1, i1 = ctx.frame_hits(n_px)
    s1 = ctx.frame_ansi()
    culled = ctx.timings()["sphere_tests"]
    assert np.array_equal(i0, i1) and d0.tobytes() == d1.tobytes() and np.array_equal(s0, s1)
    assert brute >= n_px * 1024 and culled < brute // 4, (brute, culled)


def test_packet_filter_is_invisible(ctx, oracle, rtc):
    """RTC_FLAG_PACKET (the ray-sphere filter on the first and last ray of every thread's 8-ray packet instead of on every
    ray) must not change a hit record, colour or stream byte: whole frames against the oracle, alone and with culling /
    shadow rays, 8-ray and 4-ray packets (console-sized frames), chunked sphere lists, camera inside the cloud, and --
    the case the chord bound exists for -- spheres much smaller than a packet, sitting between its end rays."""
    objs = scenes.config_scene("config2_1080p_64")
    p = scenes.config_camera("config2_1080p_64")
    check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_PACKET)
    check_frame(ctx, oracle, objs, p, RGB_ASCII, flags=FLAG_PACKET | FLAG_CULL | FLAG_SHADOWS)
    check_frame(ctx, oracle, objs, p, BIT_PIXEL, flags=FLAG_PACKET | FLAG_CULL)
    small = rtc.camera_params(64, 20, (0, 0, 0), (0, PI32, 0))
    for mode in (RGB_PIXEL, BIT_ASCII):
        check_frame(ctx, oracle, scenes.default_scene(), small, mode, flags=FLAG_PACKET)
        check_frame(ctx, oracle, scenes.default_scene(), small, mode, flags=FLAG_PACKET | FLAG_CULL)
    check_frame(ctx, oracle, scenes.default_scene(), rtc.camera_params(400, 150, (0, 0, 0), (0, PI32, 0)), RGB_ASCII, flags=FLAG_PACKET)
    many = scenes.random_spheres(9001, 33)
    q = rtc.camera_params(241, 120, (0, 0, -120), (0, PI32, 0), 1.0 / 240)
    check_frame(ctx, oracle, many, q, RGB_PIXEL, flags=FLAG_PACKET)
    check_frame(ctx, oracle, many, q, RGB_PIXEL, flags=FLAG_PACKET | FLAG_CULL)
    inside = rtc.camera_params(161, 90, (3, -2, 5), (0.3, 1.0, 0), 1.0 / 160)
    check_frame(ctx, oracle, scenes.random_spheres(700, 34), inside, RGB_PIXEL, flags=FLAG_PACKET)
    check_frame(ctx, oracle, scenes.random_spheres(700, 34), inside, RGB_PIXEL, flags=FLAG_PACKET | FLAG_CULL | FLAG_SHADOWS)
    # degenerate frames (a packet as tall as the frame: the deflation saturates and every sphere is a candidate), edge scenes,
    # the reference's sphere capacity (21 chunks), a non-camera matrix (the flag is ignored: dot-product filter)
    for (x, y) in ((2, 1), (3, 2), (18, 9), (38, 23)):
        check_frame(ctx, oracle, scenes.default_scene(), rtc.camera_params(x, y, (0, 3, -10), (0.2, PI32, 0)), RGB_PIXEL, flags=FLAG_PACKET)
    e = rtc.camera_params(64, 20, (0, 0, 0), (0, PI32, 0))
    check_frame(ctx, oracle, np.zeros(0, OBJECT_DTYPE), e, RGB_PIXEL, flags=FLAG_PACKET)
    check_frame(ctx, oracle, np.array([scenes.make_sphere((0, 0, 1), 50.0, (200, 20, 30))], OBJECT_DTYPE), e, RGB_PIXEL, flags=FLAG_PACKET)
    check_frame(ctx, oracle, np.array([scenes.make_sphere((0, 0, 30), 5.0, (10, 200, 30)), scenes.make_sphere((0, 0, 30), 5.0, (200, 10, 30))],
                                      OBJECT_DTYPE), e, RGB_PIXEL, flags=FLAG_PACKET)
    check_frame(ctx, oracle, scenes.random_spheres(52083, 41), rtc.camera_params(34, 9, (0, 0, -120), (0, PI32, 0), 1.0 / 33), RGB_PIXEL,
                flags=FLAG_PACKET | FLAG_CULL)
    skew = rtc.camera_params(161, 90, (0, 0, -120), (0, PI32, 0), 1.0 / 160)
    skew.inv_view[1] = 0.3                                                    # 3x3 no longer orthonormal
    check_frame(ctx, oracle, scenes.random_spheres(300, 35), skew, RGB_PIXEL, flags=FLAG_PACKET)
    # sub-pixel to few-pixel spheres all over a tall frame (a pixel is 1.07e-3 rad high; radii 2e-4 .. 4e-3 rad), rolled and
    # pitched camera, ragged width: many of them lie strictly between the end rays of a packet
    rng = np.random.default_rng(11)
    n_s = 600
    u = rng.normal(size=(n_s, 3)); u[:, 2] = np.abs(u[:, 2]) + 0.6
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    dist = rng.uniform(30.0, 200.0, n_s)
    rad = dist * rng.choice([2e-4, 5e-4, 1e-3, 2e-3, 4e-3], n_s)
    tiny = np.array([scenes.make_sphere(tuple(float(v) for v in (u[i] * dist[i])), float(rad[i]),
                                        (int(rng.integers(0, 255)), int(rng.integers(0, 255)), int(rng.integers(0, 255)))) for i in range(n_s)], OBJECT_DTYPE)
    for rot in ((0.0, float(PI32), 0.0), (0.21, float(PI32) + 0.4, 0.0)):
        t = rtc.camera_params(598, 1080, (0.0, 0.0, 0.0), rot, 1.0 / 597)
        s_plain = check_frame(ctx, oracle, tiny, t, RGB_PIXEL)
        s_pack = check_frame(ctx, oracle, tiny, t, RGB_PIXEL, flags=FLAG_PACKET)
        s_both = check_frame(ctx, oracle, tiny, t, RGB_PIXEL, flags=FLAG_PACKET | FLAG_CULL)
        assert np.array_equal(s_plain, s_pack) and np.array_equal(s_plain, s_both)
        assert len(s_plain) > 1080 + 20 * 597                               # (something is hit)
    # full-size configs 3 and 4 (two sphere chunks): hit records and stream against the per-ray filter
    for name in ("config3_4k_1024", "config4_8k_4096"):
        objs = scenes.config_scene(name)
        p = scenes.config_camera(name, frame=7) if name == "config4_8k_4096" else scenes.config_camera(name)
        n_px = (p.x - 1) * p.y
        ctx.set_objects(objs)
        ctx.render(p, RGB_PIXEL, FLAG_KEEP_HITS)
        d0, i0 = ctx.frame_hits(n_px)
        s0 = np.array(ctx.frame_ansi())
        for fl in (FLAG_PACKET, FLAG_PACKET | FLAG_CULL):
            ctx.render(p, RGB_PIXEL, fl | FLAG_KEEP_HITS)
            d1, i1 = ctx.frame_hits(n_px)
            assert np.array_equal(i0, i1) and d0.tobytes() == d1.tobytes() and np.array_equal(s0, ctx.frame_ansi()), (name, fl)
            ctx.render(p, RGB_PIXEL, fl)
            assert np.array_equal(s0, ctx.frame_ansi()), (name, fl)


def test_quantisers_exhaustive(ctx, oracle, rtc):
    """The integer quantisers over their whole domains: xterm-256 index of all 2^24 RGB values (shade kernel) and the
    NUL-padded decimal digits of all 256 byte values in every channel position (encoder), against the oracle (which
    tests/test_oracle_vs_reference.py pins exhaustively against the reference)."""
    import torch
    bind_stream(ctx)
    cube = torch.empty(1 << 24, dtype=torch.uint8, device="cuda")
    ctx.ansi256_cube(cube.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(cube.cpu().numpy(), oracle.ansi256_range(0, 1 << 24))
    # every byte value in every channel, each cell different from its neighbour: 3 x 256 cells + 8-bit mode 256 cells
    v = np.arange(256, dtype=np.uint8)
    z = np.zeros(256, np.uint8)
    rgb = np.concatenate([np.stack([v, z, z], 1), np.stack([z + 7, v, z], 1), np.stack([z, z + 9, v], 1)]).reshape(-1)
    for mode, keys, x, y in ((RGB_PIXEL, rgb, 257, 3), (RGB_ASCII, rgb, 129, 6), (BIT_PIXEL, v, 65, 4), (BIT_ASCII, v, 257, 1)):
        W = x - 1
        glyph = (np.arange(W * y) % 5 == 0).astype(np.uint8) * 3 + 32 if mode_has_glyph(mode) else None
        dk = torch.from_numpy(keys.copy()).cuda()
        dg = torch.from_numpy(glyph).cuda() if glyph is not None else None
        cap = rtc.encode_capacity(x, y, mode)
        out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.encode(dk.data_ptr(), dg.data_ptr() if dg is not None else 0, x, y, mode, out.data_ptr(), cap, total.data_ptr())
        torch.cuda.synchronize()
        got = out[:int(total.item())].cpu().numpy()
        assert np.array_equal(got, oracle.encode_planes(keys, glyph, x, y, mode)), MODE_NAMES[mode]
    unbind_stream(ctx)


def test_random_scenes_sweep(ctx, oracle, rtc):
    """Seeded sweep of small frames: random sphere clouds (radius 0 included, overlapping and nested spheres), planes
    of either facing, cameras outside / inside the cloud / inside a sphere, arbitrary rotations and pixel aspects,
    every flag combination (each also with the packet filter) -- hit records bit-exact, colours and stream as in check_frame."""
    rng = np.random.default_rng(2024)
    for case in range(72):
        n = int(rng.integers(1, 400))
        objs = scenes.random_spheres(n, 1000 + case)
        if case % 3 == 0:                                            # nested / coincident spheres, exact distance ties
            objs[: n // 2]["center"] = objs[n // 2: 2 * (n // 2)]["center"]
        extra = []
        if case % 2 == 0:
            extra.append(scenes.make_plane((0, float(rng.integers(-70, -20)), 0), (0, 1, 0), (100, 100, 100), 300, 300))
        if case % 4 == 1:
            extra.append(scenes.make_plane((0, 0, 40), (0.3, 0.2, -1.0), (10, 200, 100), 120, 90))
        if extra:
            objs = np.concatenate([objs, np.array(extra, OBJECT_DTYPE)])
        x, y = int(rng.integers(18, 150)), int(rng.integers(5, 70))
        if case % 5 == 0:
            pos = tuple(float(v) for v in rng.uniform(-40, 40, 3))   # inside the cloud
        elif case % 5 == 1:
            c0 = objs[0]["center"]; pos = (float(c0[0]) + 0.1, float(c0[1]), float(c0[2]))   # inside / on a sphere
        else:
            pos = tuple(float(v) for v in rng.uniform(-1, 1, 3) * 60 + np.array([0, 0, -130]))
        rot = (float(rng.uniform(-0.6, 0.6)), float(PI32 + rng.uniform(-0.8, 0.8)), 0.0)
        p = rtc.camera_params(x, y, pos, rot, float(rng.choice([0.0, 1.0 / (x - 1), 0.02])))
        mode = [RGB_PIXEL, RGB_ASCII, BIT_PIXEL, BIT_ASCII, RGB_NORMALS][case % 5]
        flags = [0, FLAG_CULL, FLAG_SHADOWS, FLAG_CULL | FLAG_SHADOWS][case % 4]
        check_frame(ctx, oracle, objs, p, mode, flags=flags)
        check_frame(ctx, oracle, objs, p, mode, flags=flags | FLAG_PACKET)


def test_encoder_state_survives_skipped_launches(ctx, oracle, rtc):
    """The encoder's group accumulators are zeroed by the PREVIOUS count launch (two parities).  Frames that launch no
    count kernel (SDL, 1-column consoles, renders that fail validation) must not disturb that: a large frame after an
    odd number of them used to add onto stale sums (round-1 advisor finding)."""
    objs = scenes.config_scene("config2_1080p_64")
    p = rtc.camera_params(401, 150, (0, 0, -120), (0, PI32, 0), 1.0 / 400)      # 60000 cells = 47 tiles (> one 32-tile group)
    one_col = rtc.camera_params(1, 7, (0, 0, -120), (0, PI32, 0), 1.0)
    ctx.set_objects(objs)
    ctx.render(p, RGB_PIXEL)
    want = ctx.frame_ansi()
    assert np.array_equal(want, oracle.render(objs, p, RGB_PIXEL))
    for skipped in ("sdl", "one_column", "bad_mode", "sdl+sdl+sdl"):
        for _ in range(skipped.count("+") + 1):
            if skipped.startswith("sdl"):
                ctx.render(p, SDL)
                assert ctx.frame_ansi().tobytes() == b"\n" * p.y
            elif skipped == "one_column":
                ctx.render(one_col, RGB_PIXEL)
                assert ctx.frame_ansi().tobytes() == b"\n" * 7
            else:
                with pytest.raises(rtc.RtcError):
                    ctx.render(p, 17)
        ctx.render(p, RGB_PIXEL)
        assert np.array_equal(ctx.frame_ansi(), want), f"frame after a skipped encoder launch ({skipped}) differs"
        ctx.render(p, BIT_ASCII)                                             # and the other cell size, same scratch
        assert np.array_equal(ctx.frame_ansi(), oracle.render(objs, p, BIT_ASCII))


def test_group_reject_bound_tiny_spheres(ctx, oracle, rtc):
    """Radius-0 and tiny spheres sitting ON a larger surface: the reference's rounded hit distance of a near-tangent hit
    comes out up to ~3e-3 |oc| nearer than the geometric |oc| - r (rounding of b*b - 4ac amplified by the square root),
    so the reference sometimes accepts the tiny sphere in front of the big one.  The per-group reject bound must not
    drop those candidates (round-1 advisor finding: the geometric bound did).  Both camera facings, so that the big
    sphere is tested before (low Morton code) and after the points."""
    rng = np.random.default_rng(7)
    n_pts = 700
    for facing in (+1.0, -1.0):
        cam = (0.0, 0.0, 120.0 * facing)
        cz = -40.0 * facing
        u = rng.normal(size=(n_pts, 3))
        u[:, 2] = np.abs(u[:, 2]) * facing                                  # the hemisphere facing the camera
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        centers = (np.array([0.0, 0.0, cz]) + 60.0 * u).astype(np.float32)
        radii = np.where(np.arange(n_pts) % 3 == 0, 0.0, rng.choice([0.01, 0.05, 0.1], n_pts)).astype(np.float32)
        big = [scenes.make_sphere((0.0, 0.0, cz), 60.0, (200, 40, 40)) for _ in range(4)]
        pts = [scenes.make_sphere(tuple(float(v) for v in centers[i]), float(radii[i]), (20, 250, 20)) for i in range(n_pts)]
        objs = np.array(big + pts * 4, OBJECT_DTYPE)                        # x4: packed groups of identical spheres
        rot = (0.0, 0.0 if facing > 0 else float(PI32), 0.0)
        p = rtc.camera_params(961, 540, cam, rot, 1.0 / 960)
        check_frame(ctx, oracle, objs, p, RGB_PIXEL)
        check_frame(ctx, oracle, objs, p, RGB_PIXEL, flags=FLAG_CULL)


def test_encoder_full_size_bytes(ctx, oracle, rtc):
    """Config 5 at full size, every byte: 7681x4320 i.i.d. random RGB (663 MB of stream) against the oracle's serial
    MinimizeRGB restatement; and the three other cell formats (RGB_ASCII, BIT_PIXEL, BIT_ASCII) on 3841x2160 planes with
    runs and random glyphs."""
    import torch
    bind_stream(ctx)
    cases = [(7681, 4320, RGB_PIXEL, "noise"), (3841, 2160, RGB_ASCII, "runs"), (3841, 2160, BIT_PIXEL, "noise"), (3841, 2160, BIT_ASCII, "runs")]
    for (x, y, mode, pattern) in cases:
        W, bpp = x - 1, mode_bpp(mode)
        g = torch.Generator(device="cuda"); g.manual_seed(5 + mode)
        if pattern == "noise":
            keys = torch.randint(0, 256, (W * y * bpp,), dtype=torch.uint8, device="cuda", generator=g)
        else:                                                         # runs of 1..8 equal cells, as a rendered frame has
            n_runs = W * y // 3
            base = torch.randint(0, 256, (n_runs, bpp), dtype=torch.uint8, device="cuda", generator=g)
            lens = torch.randint(1, 9, (n_runs,), device="cuda", generator=g)
            keys = torch.repeat_interleave(base, lens, dim=0)[:W * y].contiguous().view(-1)
            assert keys.numel() == W * y * bpp
        glyph = None
        if mode_has_glyph(mode):
            table = torch.tensor(list(b"  .:#@"), dtype=torch.uint8, device="cuda")
            glyph = table[torch.randint(0, 6, (W * y,), device="cuda", generator=g)]
        cap = rtc.encode_capacity(x, y, mode)
        out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        ctx.encode(keys.data_ptr(), glyph.data_ptr() if glyph is not None else 0, x, y, mode, out.data_ptr(), cap, total.data_ptr())
        torch.cuda.synchronize()
        got = out[:int(total.item())].cpu().numpy()
        want = oracle.encode_planes(keys.cpu().numpy(), glyph.cpu().numpy() if glyph is not None else None, x, y, mode)
        assert got.size == want.size and np.array_equal(got, want), (x, y, MODE_NAMES[mode])
        del out, keys, got, want
    unbind_stream(ctx)


def test_config4_full_size(ctx, oracle):
    """Config 4 at full size (7681x4320, 4096 spheres = two shared-memory chunks of the sphere list, orbit cameras): three
    orbit frames, three row windows each -- hit index and distance '==', colour +-1 LSB -- brute force and with per-tile
    culling; and the culled frame's stream == the brute-force frame's stream."""
    name = "config4_8k_4096"
    objs = scenes.config_scene(name)
    ctx.set_objects(objs)
    for frame in (0, 37, 95):
        p = scenes.config_camera(name, frame=frame, n_frames=120)
        W, y = p.x - 1, p.y
        n_px = W * y
        streams = []
        for flags in (0, FLAG_CULL):
            ctx.render(p, RGB_PIXEL, flags)
            streams.append(ctx.frame_ansi())
            ctx.render(p, RGB_PIXEL, flags | FLAG_KEEP_HITS)
            dist, index = ctx.frame_hits(n_px)
            color, _ = ctx.frame_color(n_px)
            for (r0, r1) in [(0, 2), (y // 2 - 1, y // 2 + 2), (y - 3, y)]:
                o = oracle.trace_planes(objs, p, RGB_PIXEL, row0=r0, row1=r1, nthreads=16)
                sl = slice(r0 * W, r1 * W)
                assert np.array_equal(index[sl], o["index"]), (frame, flags, r0)
                assert dist[sl].tobytes() == o["dist"].tobytes(), (frame, flags, r0)
                d = np.abs(color[r0 * W * 3:r1 * W * 3].astype(np.int16) - o["color"].astype(np.int16))
                assert d.max(initial=0) <= RGB_TOL and (d != 0).sum() <= 2, (frame, flags, r0)
            del dist, index, color
        assert np.array_equal(streams[0], streams[1]), f"frame {frame}: culling changed the stream"
        assert np.array_equal(streams[0][:200000], oracle.encode_planes(ctx.frame_color(n_px)[0], None, p.x, p.y, RGB_PIXEL)[:200000])


def test_light_parameters(ctx, oracle, rtc):
    """rtc_set_light: the reference's hard-coded light / material constants (RayTracing.cu:143-152, :77) as context state.
    NULL restores them; other values change the frame and agree with the oracle evaluated with the same constants."""
    objs = scenes.default_scene()
    p = rtc.camera_params(161, 60, (0, 0, 0), (0, PI32, 0))
    ctx.set_objects(objs)
    ctx.set_light(None)
    ref = np.array(ctx.update(p, RGB_PIXEL, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))
    light = [-20.0, 30.0, 5.0, 1.0, 1500.0, 0.5, 2500.0, 0.1, 0.3, 0.2, 1.0]
    try:
        ctx.set_light(light)
        oracle.set_light(light)
        for mode, flags in ((RGB_PIXEL, 0), (RGB_ASCII, FLAG_SHADOWS), (BIT_PIXEL, 0)):
            s = check_frame(ctx, oracle, objs, p, mode, flags=flags)
            if mode == RGB_PIXEL:
                assert not np.array_equal(s, ref)
    finally:
        ctx.set_light(None)
        oracle.set_light(None)
    ctx.set_objects(objs)
    assert np.array_equal(np.array(ctx.update(p, RGB_PIXEL, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT)), ref)
