"""Row-band partitioning: the host-side arithmetic of the multi-GPU split, and a torch.distributed gather of band planes.

The product's multi-GPU frame driver is C++ behind the C-ABI (csrc/rtc_mgpu.cu: one process, a worker thread + stream per
device, bands planned by rtc_plan_bands).  This module is what is left on the Python side: the same band arithmetic
(checked against rtc_plan_bands in tests/test_multirank_cpu.py) and a point-to-point gather of ragged bands to rank 0
over any torch.distributed backend, used by the world-size-2/3 gloo tests of the band + seam logic on CPU.
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


def weighted_bands(y, world, deficit_rows=0.0, align=1, wave_units=0):
    """Bands for the case where rank 0 also runs the encoder: rank 0 gets `deficit_rows` fewer rows than the
    others (deficit = encode time / trace+shade time per row, measured), the remaining rows are split evenly.
    `align` > 1 puts the boundaries on multiples of `align` rows (16 = the ray kernel's tile height: a band that ends
    inside a tile row pays for the whole row, and one partial tile row too many can cost a whole extra tile wave).
    `wave_units` > 0: the number of `align`-row units one tile wave of the ray kernel covers (SMs x warps per CTA /
    tiles per row); if the deficit would push the other ranks just over one wave while everything fits into one wave
    per rank, rank 0 takes the excess instead (a second wave costs more than the imbalance).
    Contiguous, exact cover of [0, y) for any y, world; deficit 0 and align 1 reproduce bands()."""
    if world == 1:
        return [(0, y)]
    d = max(0.0, float(deficit_rows))
    if d == 0.0 and align <= 1:
        return bands(y, world)
    align = max(1, int(align))
    units = (y + align - 1) // align                      # rows in units of `align`
    du = d / align
    n0 = int(((units + du) / world - du) + 0.5) if (units + du) / world - du > 0 else 0     # round half up, as rtc_plan_bands
    n0 = max(0, min(units, n0))
    if wave_units > 0 and units <= wave_units * world:
        per_other = -(-(units - n0) // (world - 1))
        if per_other > wave_units:
            n0 = units - wave_units * (world - 1)
    edges = [0, n0]
    rest = units - n0
    for g in range(world - 1):
        edges.append(n0 + band(rest, g, world - 1)[1])
    rows = [min(y, e * align) for e in edges]
    rows[-1] = y
    return [(rows[g], rows[g + 1]) for g in range(world)]
