"""CPU, world_size 2 and 3 over gloo: the host-side multi-GPU logic (row-band partition, gather of
the bands to rank 0, one encode over the assembled frame) reproduces the single-rank stream.
Per-band pixels come from the oracle here (no GPU); the GPU run of the same path is bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, x, y, out_path):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rtc_b200
    from rtc_b200 import multigpu, scenes
    from rtc_b200._types import mode_bpp, mode_has_glyph
    from oracle.oracle import Oracle
    orc = Oracle()
    objs = scenes.default_scene()
    p = rtc_b200.camera_params(x, y, (0, 0, 0), (0, np.float32(np.pi), 0))
    W, bpp, gl = x - 1, mode_bpp(mode), mode_has_glyph(mode)
    r0, r1 = multigpu.band(y, rank, world)
    pl = orc.trace_planes(objs, p, mode, row0=r0, row1=r1, nthreads=1)
    band_color = torch.from_numpy(pl["color"].copy())
    band_glyph = torch.from_numpy(pl["glyph"].copy()) if gl else None
    frame_color = torch.zeros(W * y * bpp, dtype=torch.uint8) if rank == 0 else None
    frame_glyph = torch.zeros(W * y, dtype=torch.uint8) if (rank == 0 and gl) else None
    if rank == 0:
        frame_color[r0 * W * bpp:r1 * W * bpp] = band_color
        if gl:
            frame_glyph[r0 * W:r1 * W] = band_glyph
    multigpu.gather_planes(dist, rank, world, y, W, bpp, band_color, frame_color, band_glyph, frame_glyph)
    if rank == 0:
        stream = orc.encode_planes(frame_color.numpy(), frame_glyph.numpy() if gl else None, x, y, mode)
        np.save(out_path, stream)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode,size", [(2, 3, (70, 23)), (3, 2, (41, 10)), (2, 0, (33, 7))])
def test_rowband_gather_matches_single_rank(tmp_path, oracle, rtc, world, mode, size):
    x, y = size
    out = str(tmp_path / "stream.npy")
    mp.spawn(_worker, args=(world, _free_port(), mode, x, y, out), nprocs=world, join=True)
    from rtc_b200 import scenes
    p = rtc.camera_params(x, y, (0, 0, 0), (0, np.float32(np.pi), 0))
    want = oracle.render(scenes.default_scene(), p, mode)
    assert np.array_equal(np.load(out), want)


def test_band_partition_properties(rtc):
    from rtc_b200 import multigpu
    for y in (1, 7, 64, 2160, 4320, 4321):
        for world in (1, 2, 3, 4, 8):
            bs = multigpu.bands(y, world)
            assert bs[0][0] == 0 and bs[-1][1] == y
            assert all(bs[i][1] == bs[i + 1][0] for i in range(world - 1))          # contiguous, no overlap
            sizes = [b - a for a, b in bs]
            assert max(sizes) - min(sizes) <= 1                                      # balanced to one row


def test_weighted_bands(rtc):
    """Rank 0 also encodes, so it gets fewer rows; the cover stays exact and contiguous."""
    from rtc_b200 import multigpu
    for y in (1, 7, 64, 2160, 4321):
        for world in (1, 2, 3, 8):
            assert multigpu.weighted_bands(y, world, 0.0) == multigpu.bands(y, world)
            for d in (0.5, 3.7, 62.0, 1e9):
                bs = multigpu.weighted_bands(y, world, d)
                assert len(bs) == world and bs[0][0] == 0 and bs[-1][1] == y
                assert all(bs[i][1] == bs[i + 1][0] and bs[i][0] <= bs[i][1] for i in range(world - 1))
                if world > 1:
                    others = [b - a for a, b in bs[1:]]
                    assert bs[0][1] - bs[0][0] <= max(others) and max(others) - min(others) <= 1
    assert multigpu.weighted_bands(2160, 8, 62.0)[0] == (0, 216)
    for y in (1, 7, 64, 270, 2160, 4321):
        for world in (2, 3, 8):
            for d in (0.0, 40.0, 62.0, 1e9):
                bs = multigpu.weighted_bands(y, world, d, align=16)
                assert len(bs) == world and bs[0][0] == 0 and bs[-1][1] == y
                assert all(bs[i][1] == bs[i + 1][0] and bs[i][0] <= bs[i][1] for i in range(world - 1))
                assert all(b[1] % 16 == 0 or b[1] == y for b in bs)
    # 4K frame on 8 GPUs: 17 tile rows x 240 tiles fit one wave of 148 SMs x 28 warps, 18 do not -> rank 0 keeps 16
    assert multigpu.weighted_bands(2160, 8, 62.0, align=16, wave_units=17) == [(0, 256)] + [(256 + 272 * g, 256 + 272 * (g + 1)) for g in range(7)]
    assert multigpu.weighted_bands(2160, 8, 62.0, align=16)[0] == (0, 208)


def test_plan_bands_matches_python_and_tiles_exactly(rtc):
    """rtc_plan_bands (the C++ band planner the multi-GPU driver uses; host only) tiles [0, y) exactly for any y, n,
    alignment and encoder deficit, and agrees with the Python restatement in rtc_b200.multigpu."""
    from rtc_b200 import multigpu
    rng = np.random.default_rng(3)
    for _ in range(400):
        y = int(rng.integers(1, 5000))
        n = int(rng.integers(1, 9))
        align = int(rng.choice([1, 16]))
        deficit = float(rng.choice([0.0, 0.0, rng.uniform(0, 200)]))
        wave = int(rng.choice([0, 17, 8]))
        got = rtc.plan_bands(y, n, align, deficit, wave)
        assert got[0][0] == 0 and got[-1][1] == y and len(got) == n
        assert all(a <= b for a, b in got) and all(got[g][1] == got[g + 1][0] for g in range(n - 1))
        assert got == [tuple(b) for b in multigpu.weighted_bands(y, n, deficit, align, wave)], (y, n, align, deficit, wave)
    assert rtc.plan_bands(2160, 8) == multigpu.bands(2160, 8)
    with pytest.raises(rtc.RtcError):
        rtc.plan_bands(100, 0)
