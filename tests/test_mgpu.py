"""GPU (-m gpu): the multi-GPU frame driver behind the C-ABI (rtc_mgpu_*, csrc/rtc_mgpu.cu) -- one process, one
worker thread + stream per device, row bands -- against single-context frames, byte for byte.

On the 1-GPU test box the "devices" are several contexts on cuda:0 (device_ids = [0, 0, ...]: every code path of the
driver except the physical peer link); with >= 2 GPUs visible the same tests also run across real devices."""
import numpy as np
import pytest

from rtc_b200 import scenes
from rtc_b200._types import (BIT_ASCII, BIT_PIXEL, FLAG_CULL, FLAG_PACKET, FLAG_SHADOWS, FLAG_UPDATE_REF_LAUNCH_LIMIT, MODE_NAMES,
                             RGB_ASCII, RGB_NORMALS, RGB_PIXEL, SDL)
from util import PI32

pytestmark = pytest.mark.gpu


def n_devices():
    import torch
    return torch.cuda.device_count()


def device_sets():
    sets = [[0], [0, 0], [0, 0, 0], [0, 0, 0, 0, 0]]
    return sets


def moving_frames(rtc, x, y, n):
    ps = [rtc.camera_params(x, y, (0.9 * k - 2.0, 0.3 * k, -120 + 0.5 * k), (0.01 * k, PI32 + 0.02 * k, 0), 1.0 / (x - 1)) for k in range(n)]
    modes = [RGB_PIXEL, RGB_ASCII, BIT_PIXEL, RGB_PIXEL, BIT_ASCII, RGB_NORMALS, RGB_PIXEL, SDL, RGB_PIXEL]
    return ps, [modes[k % len(modes)] for k in range(n)]


def single_gpu_frames(ctx, objs, ps, modes, dt, flags):
    ctx.set_objects(objs)
    return [np.array(ctx.update(p, m, dt=dt, flags=flags)) for p, m in zip(ps, modes)]


@pytest.mark.parametrize("gather", ["host", "p2p"])
@pytest.mark.parametrize("devices", device_sets(), ids=lambda d: "x".join(map(str, d)))
def test_mgpu_pipelined_frames_equal_single_gpu(ctx, rtc, gather, devices):
    """>= 9 pipelined frames (three in flight), moving camera, physics step every frame, all rendering modes: the
    assembled N-band stream == the single-context stream, frame by frame (band seams carry MinimizeRGB's latestColor,
    RayTracingManager.cu:262-301)."""
    objs = np.concatenate([scenes.random_spheres(300, 77), np.array([scenes.bench_plane()], scenes.OBJECT_DTYPE)])
    x, y = 322, 131                                                    # W = 321: not a multiple of 16 (byte-store epilogue)
    ps, modes = moving_frames(rtc, x, y, 9)
    want = single_gpu_frames(ctx, objs, ps, modes, 0.07, 0)
    g = rtc.GATHER_HOST if gather == "host" else rtc.GATHER_P2P
    with rtc.MultiGpu(devices, g) as m:
        m.set_objects(objs)
        got = []
        m.submit(ps[0], modes[0], 0.07)
        m.submit(ps[1], modes[1], 0.07)
        for k in range(len(ps)):
            if k + 2 < len(ps):
                m.submit(ps[k + 2], modes[k + 2], 0.07)
            got.append(m.collect(copy=True))
        for k in range(len(ps)):
            assert np.array_equal(got[k], want[k]), f"frame {k} ({MODE_NAMES[modes[k]]}) differs on {len(devices)} bands, gather {gather}"
        info = m.last_frame()
        assert info["bands"][0][0] == 0 and info["bands"][-1][1] == y and len(info["device_ms"]) == len(devices)
        st = m.host_stats()                                          # host-side accounting of the driver's worker threads
        assert len(st["enqueue_us"]) == len(devices) and all(v > 0 for v in st["enqueue_us"])
        workers, main = m.debug_trace()
        assert workers.shape == (len(devices), 64, 6) and main.shape == (64, 2)
        # the replicas ran the same physics
        ctx.set_objects(objs)
        for _ in range(len(ps)):
            ctx.update_objects(0.07)
        assert m.get_objects().tobytes() == ctx.get_objects().tobytes()
        with pytest.raises(rtc.RtcError, match="no frame in flight"):
            m.collect()


@pytest.mark.parametrize("gather", ["host", "p2p"])
def test_mgpu_aligned_width_flags_and_scene_api(ctx, rtc, gather):
    """W % 16 == 0 (16-byte-store epilogue, the path peer stores take), shadows / culling flags, ragged explicit bands
    (1-row and empty bands), the incremental scene API, synchronous update."""
    g = rtc.GATHER_HOST if gather == "host" else rtc.GATHER_P2P
    objs = scenes.config_scene("config2_1080p_64")
    x, y = 481, 270
    p = rtc.camera_params(x, y, (0, 0, -120), (0, PI32, 0), 1.0 / (x - 1))
    with rtc.MultiGpu([0, 0, 0, 0], g) as m:
        m.set_objects(objs)
        for mode in (RGB_PIXEL, RGB_ASCII, BIT_ASCII, BIT_PIXEL):
            for flags in (0, FLAG_CULL, FLAG_SHADOWS, FLAG_CULL | FLAG_SHADOWS, FLAG_PACKET, FLAG_PACKET | FLAG_CULL | FLAG_SHADOWS):
                ctx.set_objects(objs)
                want = np.array(ctx.update(p, mode, dt=0.0, flags=flags | FLAG_UPDATE_REF_LAUNCH_LIMIT))
                got = m.update(p, mode, 0.0, flags | FLAG_UPDATE_REF_LAUNCH_LIMIT, copy=True)
                assert np.array_equal(got, want), (MODE_NAMES[mode], flags)
        m.set_bands(y, [0, 67, 68, 68, 270])                            # ragged: a 1-row band and an empty band
        ctx.set_objects(objs)
        want = np.array(ctx.update(p, RGB_ASCII, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))
        assert np.array_equal(m.update(p, RGB_ASCII, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT, copy=True), want)
        assert m.last_frame()["bands"] == [[0, 67], [67, 68], [68, 68], [68, 270]]
        m.set_bands(y, None)
        # incremental scene API == bulk upload
        m.clear()
        d = scenes.default_scene()
        for o in d:
            if o["type"] == 2:
                m.add_sphere(o["center"], o["radius"], o["color"], o["speed"], o["mover"])
            else:
                m.add_plane(o["center"], (0.0, 3.0, 0.0), o["color"], o["width"], o["height"])
        assert m.get_objects().tobytes() == d.tobytes()
        q = rtc.camera_params(240, 64, (0, 0, 0), (0, PI32, 0))
        ctx.set_objects(d)
        want = np.array(ctx.update(q, RGB_PIXEL, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))
        assert np.array_equal(m.update(q, RGB_PIXEL, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT, copy=True), want)
        # a new scene submitted while frames are in flight takes effect with the next submit, not earlier
        m.submit(q, RGB_PIXEL, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
        m.set_objects(objs)
        m.submit(p, RGB_PIXEL, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
        assert np.array_equal(m.collect(copy=True), want)
        ctx.set_objects(objs)
        assert np.array_equal(m.collect(copy=True), np.array(ctx.update(p, RGB_PIXEL, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT)))
        # errors come back as codes, not hangs
        with pytest.raises(rtc.RtcError):
            m.submit(p, 17)
        for _ in range(3):
            m.submit(p, RGB_PIXEL)
        with pytest.raises(rtc.RtcError, match="in flight"):
            m.submit(p, RGB_PIXEL)
        for _ in range(3):
            m.collect()


def test_mgpu_config3_full_size(ctx, rtc):
    """Config 3 (3841x2160, 1024 spheres + plane) on 8 bands, both gathers == the single-context frame; p2p band weights
    calibrate themselves after three frames without changing a byte."""
    objs = scenes.config_scene("config3_4k_1024")
    p = scenes.config_camera("config3_4k_1024")
    ctx.set_objects(objs)
    want = np.array(ctx.update(p, RGB_PIXEL, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))
    for g in (rtc.GATHER_HOST, rtc.GATHER_P2P):
        with rtc.MultiGpu([0] * 8, g) as m:
            m.set_objects(objs)
            for k in range(5):
                got = m.update(p, RGB_PIXEL, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
                assert np.array_equal(got, want), f"gather {g}, frame {k}"
            bands = m.last_frame()["bands"]
            assert bands[0][0] == 0 and bands[-1][1] == p.y
            if g == rtc.GATHER_P2P:
                assert bands[0][1] - bands[0][0] <= bands[1][1] - bands[1][0]      # device 0 also encodes


def test_mgpu_real_devices(ctx, rtc):
    """With >= 2 GPUs visible: the same comparison across real devices (peer stores over NVLink in the p2p gather)."""
    n = n_devices()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    objs = scenes.config_scene("config3_4k_1024")
    x, y = 1921, 1080
    ps, _ = moving_frames(rtc, x, y, 8)
    modes = [RGB_PIXEL, RGB_ASCII, BIT_ASCII, RGB_PIXEL] * 2
    want = single_gpu_frames(ctx, objs, ps, modes, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
    for g in (rtc.GATHER_HOST, rtc.GATHER_P2P):
        with rtc.MultiGpu(list(range(min(n, 8))), g) as m:
            m.set_objects(objs)
            got = []
            m.submit(ps[0], modes[0], 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
            m.submit(ps[1], modes[1], 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
            for k in range(len(ps)):
                if k + 2 < len(ps):
                    m.submit(ps[k + 2], modes[k + 2], 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT)
                got.append(m.collect(copy=True))
            for k in range(len(ps)):
                assert np.array_equal(got[k], want[k]), f"{n} GPUs, gather {g}, frame {k}"


@pytest.mark.parametrize("backing", ["thp", "interleave"])
def test_mgpu_alternative_host_buffers(ctx, rtc, monkeypatch, backing):
    """RTC_MGPU_HOSTBUF: the frame buffer as registered anonymous memory (transparent huge pages / NUMA interleaving) instead
    of cudaHostAlloc -- same bytes."""
    monkeypatch.setenv("RTC_MGPU_HOSTBUF", backing)
    objs = scenes.config_scene("config2_1080p_64")
    p = rtc.camera_params(481, 270, (0, 0, -120), (0, PI32, 0), 1.0 / 480)
    ctx.set_objects(objs)
    want = np.array(ctx.update(p, RGB_PIXEL, dt=0.0, flags=FLAG_UPDATE_REF_LAUNCH_LIMIT))
    with rtc.MultiGpu([0, 0, 0], rtc.GATHER_HOST) as m:
        m.set_objects(objs)
        for _ in range(4):
            assert np.array_equal(m.update(p, RGB_PIXEL, 0.0, FLAG_UPDATE_REF_LAUNCH_LIMIT, copy=True), want)
