"""GPU (-m gpu): the multi-rank drivers with 2 and 3 ranks sharing cuda:0 (the test box has one GPU).  No kernel ever
waits on another rank here -- ranks only synchronise on the host -- so sharing a GPU is safe.  The assembled frame must
equal the single-context frame bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, x, y, n_frames, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rtc_b200
    from rtc_b200 import multigpu, scenes
    torch.cuda.set_device(0)
    ctx = rtc_b200.Context(0)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
    objs = scenes.config_scene("config2_1080p_64")
    cams = [rtc_b200.camera_params(x, y, (0.9 * k, 0.0, -120.0), (0.0, np.float32(np.pi), 0.0), 1.0 / (x - 1)) for k in range(n_frames)]
    ctx.set_objects(objs)
    r = multigpu.HostAssembledRenderer(ctx, dist, rank, world, x, y, mode)
    r.step(cams[0]); r.step(cams[0]); r.step(cams[0])          # free-running steps must not disturb submit/collect
    frames = []
    depth = 2 + (world % 2)                                    # two or three frames in flight
    sub = 0
    while sub < min(depth - 1, n_frames):
        r.submit(cams[sub]); sub += 1
    for k in range(n_frames):
        if sub < n_frames:
            r.submit(cams[sub]); sub += 1
        view, n = r.collect()
        if rank == 0:
            frames.append(view[:n].numpy().copy())
    if rank == 0:
        np.savez(os.path.join(out_dir, "frames.npz"), *frames)
    dist.barrier()
    r.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode,size", [(2, 3, (641, 271)), (3, 2, (322, 100)), (2, 0, (203, 37))])
def test_host_assembled_frames_match_single_gpu(tmp_path, ctx, rtc, world, mode, size):
    import torch.multiprocessing as mp
    from rtc_b200 import scenes
    x, y = size
    n_frames = 7
    mp.spawn(_worker, args=(world, _free_port(), mode, x, y, n_frames, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "frames.npz")
    ctx.set_objects(scenes.config_scene("config2_1080p_64"))
    for k in range(n_frames):
        p = rtc.camera_params(x, y, (0.9 * k, 0.0, -120.0), (0.0, np.float32(np.pi), 0.0), 1.0 / (x - 1))
        ctx.render(p, mode)
        want = ctx.frame_ansi()
        assert np.array_equal(got[f"arr_{k}"], want), f"frame {k}"
