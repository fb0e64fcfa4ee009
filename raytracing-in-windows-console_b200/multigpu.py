"""Row-band partitioning across ranks (one process per GPU) and the gather of the bands to rank 0.

The frame shards naturally: pixels are independent given the (small, replicated) scene and the
camera block, so rank g traces rows [y*g/G, y*(g+1)/G) with no data-path collective; the only
exchange step is the assembly of the quantised colour plane (and glyph plane) on rank 0, which
then runs the ANSI encoder over the whole frame (the minimiser's colour carry-over crosses band
boundaries, so the stream is encoded once, on the assembled planes).

Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests).
"""


def band(y, rank, world):
    """Contiguous rows [r0, r1) of rank `rank`; the bands tile [0, y) exactly for any y, world."""
    return (y * rank) // world, (y * (rank + 1)) // world


def bands(y, world):
    return [band(y, g, world) for g in range(world)]


def gather_planes(dist, rank, world, y, W, bpp, band_color, frame_color, band_glyph=None, frame_glyph=None):
    """Send every rank's band to rank 0's frame planes (rank 0's own band is expected to be written
    in place already).  Point-to-point, because bands may be ragged (y not divisible by world)."""
    if world == 1:
        return
    ops = []
    if rank == 0:
        for g in range(1, world):
            a, b = band(y, g, world)
            if b > a:
                ops.append(dist.P2POp(dist.irecv, frame_color[a * W * bpp:b * W * bpp], g))
                if frame_glyph is not None:
                    ops.append(dist.P2POp(dist.irecv, frame_glyph[a * W:b * W], g))
    else:
        a, b = band(y, rank, world)
        if b > a:
            ops.append(dist.P2POp(dist.isend, band_color[:(b - a) * W * bpp], 0))
            if band_glyph is not None:
                ops.append(dist.P2POp(dist.isend, band_glyph[:(b - a) * W], 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
