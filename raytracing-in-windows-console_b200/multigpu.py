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


def weighted_bands(y, world, deficit_rows=0.0):
    """Bands for the case where rank 0 also runs the encoder: rank 0 gets `deficit_rows` fewer rows than the
    others (deficit = encode time / trace+shade time per row, measured), the remaining rows are split evenly.
    Contiguous, exact cover of [0, y) for any y, world; deficit 0 reproduces bands()."""
    if world == 1:
        return [(0, y)]
    d = max(0.0, float(deficit_rows))
    n0 = int(round((y + d) / world - d))
    n0 = max(0, min(y, n0))
    if d == 0.0:
        return bands(y, world)
    out = [(0, n0)]
    rest = y - n0
    for g in range(world - 1):
        a, b = band(rest, g, world - 1)
        out.append((n0 + a, n0 + b))
    return out


class BandRenderer:
    """One frame across `world` ranks (one process per GPU): every rank traces + shades its row band, the bands
    are assembled on rank 0, rank 0 encodes the frame.  Everything is enqueued on the current torch stream (which the
    rtc context must share: ctx.set_stream(torch.cuda.current_stream().cuda_stream)); nothing blocks the host.

    gather="ipc":  the shade kernel's 128-bit stores go straight into rank 0's frame planes through a CUDA-IPC peer
                   mapping over NVLink -- the gather is fused into the producing kernel; a one-element all-reduce is
                   the "bands have landed" signal.  The planes are double-buffered: rank g may already write frame
                   k+1 while rank 0 still encodes frame k.
    gather="nccl": bands are written locally and moved with batched NCCL send/recv.
    """

    def __init__(self, ctx, dist, rank, world, x, y, mode, gather="ipc", deficit_rows=0.0):
        import torch
        from . import encode_capacity, mode_bpp, mode_has_glyph
        self.torch, self.ctx, self.dist, self.rank, self.world = torch, ctx, dist, rank, world
        self.x, self.y, self.W, self.mode, self.gather = x, y, x - 1, mode, gather
        self.bpp, self.gl = mode_bpp(mode), bool(mode_has_glyph(mode))
        self.bands = weighted_bands(y, world, deficit_rows)
        self.r0, self.r1 = self.bands[rank]
        self.cap = encode_capacity(x, y, mode)
        self.k = 0
        W, bpp = self.W, self.bpp
        u8 = dict(dtype=torch.uint8, device="cuda")
        nbuf = 2 if gather == "ipc" else 1
        self.frame_color = [torch.empty(W * y * bpp + 16, **u8) for _ in range(nbuf)] if rank == 0 else None
        self.frame_glyph = [torch.empty(W * y + 16, **u8) for _ in range(nbuf)] if (rank == 0 and self.gl) else None
        self.out = [torch.empty(self.cap, **u8) for _ in range(2)] if rank == 0 else None
        self.total = torch.zeros(2, dtype=torch.int64, device="cuda") if rank == 0 else None
        self.flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.band_color = self.band_glyph = None
        self.peer_color = self.peer_glyph = None
        if world > 1 and gather == "ipc":
            handles = [None]
            if rank == 0:
                handles = [[(ctx.ipc_export(self.frame_color[i].data_ptr()),
                             ctx.ipc_export(self.frame_glyph[i].data_ptr()) if self.gl else None) for i in range(2)]]
            dist.broadcast_object_list(handles, src=0)
            if rank == 0:
                self.peer_color = [t.data_ptr() for t in self.frame_color]
                self.peer_glyph = [t.data_ptr() for t in self.frame_glyph] if self.gl else [0, 0]
            else:
                self.peer_color = [ctx.ipc_open(h[0]) for h in handles[0]]
                self.peer_glyph = [ctx.ipc_open(h[1]) if self.gl else 0 for h in handles[0]]
        elif world > 1 and rank != 0:
            rows = self.r1 - self.r0
            self.band_color = torch.empty(max(1, rows * W * bpp), **u8)
            self.band_glyph = torch.empty(max(1, rows * W), **u8) if self.gl else None

    def step(self, params, flags=0):
        """Enqueue one frame.  Returns the slot (0/1) whose `out`/`total` will hold the stream on rank 0."""
        ctx, W, bpp, r0, r1, gl = self.ctx, self.W, self.bpp, self.r0, self.r1, self.gl
        slot = self.k & 1
        self.k += 1
        if self.world == 1:
            fc, fg = self.frame_color[0], (self.frame_glyph[0] if gl else None)
            ctx.trace_band(params, self.mode, 0, self.y, fc.data_ptr(), fg.data_ptr() if gl else 0, flags)
        elif self.gather == "ipc":
            ctx.trace_band(params, self.mode, r0, r1, self.peer_color[slot] + r0 * W * bpp,
                           (self.peer_glyph[slot] + r0 * W) if gl else 0, flags)
            self.dist.all_reduce(self.flag)            # stream-ordered: every band of this frame is in GPU 0's HBM
            fc, fg = (self.frame_color[slot], self.frame_glyph[slot] if gl else None) if self.rank == 0 else (None, None)
        else:
            if self.rank == 0:
                fc, fg = self.frame_color[0], (self.frame_glyph[0] if gl else None)
                dc, dg = fc[r0 * W * bpp:], (fg[r0 * W:] if gl else None)
            else:
                fc = fg = None
                dc, dg = self.band_color, self.band_glyph
            ctx.trace_band(params, self.mode, r0, r1, dc.data_ptr(), dg.data_ptr() if gl else 0, flags)
            self._gather_nccl(fc, fg)
        if self.rank == 0:
            ctx.encode(fc.data_ptr(), fg.data_ptr() if gl else 0, self.x, self.y, self.mode,
                       self.out[slot].data_ptr(), self.cap, self.total[slot:].data_ptr())
        return slot

    def _gather_nccl(self, frame_color, frame_glyph):
        dist, W, bpp = self.dist, self.W, self.bpp
        ops = []
        if self.rank == 0:
            for g in range(1, self.world):
                a, b = self.bands[g]
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, frame_color[a * W * bpp:b * W * bpp], g))
                    if frame_glyph is not None:
                        ops.append(dist.P2POp(dist.irecv, frame_glyph[a * W:b * W], g))
        else:
            a, b = self.r0, self.r1
            if b > a:
                ops.append(dist.P2POp(dist.isend, self.band_color[:(b - a) * W * bpp], 0))
                if self.band_glyph is not None:
                    ops.append(dist.P2POp(dist.isend, self.band_glyph[:(b - a) * W], 0))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def close(self):
        if self.world > 1 and self.gather == "ipc" and self.rank != 0:
            for p in self.peer_color:
                self.ctx.ipc_close(p)
            if self.gl:
                for p in self.peer_glyph:
                    self.ctx.ipc_close(p)
