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
    n0 = int(round((units + du) / world - du))
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

    def __init__(self, ctx, dist, rank, world, x, y, mode, gather="ipc", deficit_rows=0.0, align=16, wave_units=None):
        import torch
        from . import encode_capacity, mode_bpp, mode_has_glyph
        self.torch, self.ctx, self.dist, self.rank, self.world = torch, ctx, dist, rank, world
        self.x, self.y, self.W, self.mode, self.gather = x, y, x - 1, mode, gather
        self.bpp, self.gl = mode_bpp(mode), bool(mode_has_glyph(mode))
        if wave_units is None:                     # 28 warps per SM, one 16x16-pixel tile per warp and wave
            wave_units = (ctx.device_info()["sm_count"] * 28) // max(1, (x - 1 + 15) // 16) if align == 16 else 0
        self.bands = weighted_bands(y, world, deficit_rows, align, wave_units)
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


class HostAssembledRenderer:
    """Row bands with NO data-path collective: every rank traces its band (plus one context row above it), encodes
    the band itself (rtc_encode_band) and copies its piece of the stream over ITS OWN PCIe link straight into one
    frame buffer in shared, page-locked host memory, at the offset given by the stream lengths of the ranks before
    it.  The frame is assembled where the reference's sink wants it -- in host memory (PrintMachine::
    SetDataInBackBuffer) -- by 1/2/4/8 copy engines in parallel instead of one.

      submit(params)  enqueue trace + shade + encode of this rank's band (up to three frames in flight)
      collect()       oldest frame: publish this rank's stream length, wait for the lengths of the ranks before it,
                      D2H the band stream into the shared frame; on rank 0 also wait for every rank and return
                      (uint8 view of the frame, n_bytes), valid until the second collect after this one.
    With three frames in flight (submit, submit, then submit + collect per frame) the copy of frame k+1 is issued
    while frame k is being returned and overlaps both the kernels of frame k+2 and the host work of the next submit.
    Cross-process state lives in the shared segment (lengths and frame tags); ranks poll it from the host.
    """

    HDR = 4096
    NS = 3                      # frame slots: frame j lives in slot j % NS

    def __init__(self, ctx, dist, rank, world, x, y, mode, name=None):
        import ctypes
        from multiprocessing import shared_memory
        import numpy as np
        import torch
        from . import encode_capacity, mode_bpp, mode_has_glyph
        NS = self.NS
        self.torch, self.np, self.ctx, self.rank, self.world = torch, np, ctx, rank, world
        self.x, self.y, self.W, self.mode = x, y, x - 1, mode
        self.bpp, self.gl = mode_bpp(mode), bool(mode_has_glyph(mode))
        self.bands = bands(y, world)
        self.r0, self.r1 = self.bands[rank]
        self.c0 = self.r0 - 1 if self.r0 > 0 else 0                     # first traced row (context)
        rows_t = self.r1 - self.c0
        W, bpp = self.W, self.bpp
        u8 = dict(dtype=torch.uint8, device="cuda")
        self.color = torch.empty(rows_t * W * bpp + 64, **u8)
        self.glyph = torch.empty(rows_t * W + 64, **u8) if self.gl else None
        self.band_cap = encode_capacity(x, max(1, self.r1 - self.r0), mode)
        self.out = [torch.empty(self.band_cap, **u8) for _ in range(NS)]
        self.total = torch.zeros(NS, dtype=torch.int64, device="cuda")
        self.h_total = torch.zeros(NS, dtype=torch.int64).pin_memory()
        self.done_ev = [torch.cuda.Event() for _ in range(NS)]
        self.copy_ev = [torch.cuda.Event() for _ in range(NS)]
        self.copy_stream = torch.cuda.Stream()
        self.frame_cap = encode_capacity(x, y, mode)
        size = self.HDR + NS * self.frame_cap
        # Set-up is collective: rank 0 creates the segment, everybody maps and page-locks it, and either every rank
        # succeeds or every rank raises (so that callers can fall back to another gather mode together).
        self.shm, self._addr, err = None, None, None
        names = [None]
        if rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(create=True, size=size, name=name)
                names = [self.shm.name]
            except Exception as e:                                   # noqa: BLE001
                err = e
        if world > 1:
            dist.broadcast_object_list(names, src=0)
        if names[0] is None and err is None:
            err = RuntimeError("rank 0 could not create the shared frame segment")
        hdr = None
        if err is None:
            try:
                if rank != 0:
                    self.shm = shared_memory.SharedMemory(name=names[0])
                    try:                                 # only the creator unlinks; keep Python's tracker from doing it again
                        from multiprocessing import resource_tracker
                        resource_tracker.unregister(self.shm._name, "shared_memory")
                    except Exception:
                        pass
                self._addr = ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf))
                rc = torch.cuda.cudart().cudaHostRegister(self._addr, size, 1)   # portable; every rank pins its own mapping
                if int(rc) != 0:
                    self._addr = None
                    raise RuntimeError("cudaHostRegister failed: %s" % rc)
                hdr = np.frombuffer(self.shm.buf, dtype=np.int64, count=self.HDR // 8)
                if rank == 0:
                    hdr[:] = 0
            except Exception as e:                                   # noqa: BLE001
                err = e
        if world > 1:
            oks = [None] * world
            dist.all_gather_object(oks, err is None)                  # also the "header is zeroed" barrier
            if not all(oks) and err is None:
                err = RuntimeError("another rank could not set up the shared frame segment")
        if err is not None:
            self.lens = self.len_tag = self.done_tag = self.released = self.frames = None
            hdr = None
            self.close()
            raise RuntimeError("host-assembled frames unavailable: %r" % (err,))
        # header (int64): per slot lens[16], len_tag[16], done_tag[16]; then released[1]
        self.lens = [hdr[(3 * s) * 16:(3 * s) * 16 + world] for s in range(NS)]
        self.len_tag = [hdr[(3 * s + 1) * 16:(3 * s + 1) * 16 + world] for s in range(NS)]
        self.done_tag = [hdr[(3 * s + 2) * 16:(3 * s + 2) * 16 + world] for s in range(NS)]
        self.released = hdr[3 * NS * 16:3 * NS * 16 + 1]              # highest frame tag whose buffer may be overwritten
        self.frames = [torch.frombuffer(self.shm.buf, dtype=torch.uint8, count=self.frame_cap,
                                        offset=self.HDR + s * self.frame_cap) for s in range(NS)]
        self.k_sub = self.k_issue = self.k_col = self.k_step = 0
        self.t_wait_gpu = self.t_wait_len = self.t_copy = self.t_wait_done = self.t_submit = 0.0

    def step(self, params, flags=0, slot=None):
        """Device work of one frame (what `value` times): trace + shade + encode of this rank's band."""
        ctx, W, bpp = self.ctx, self.W, self.bpp
        if slot is None:                           # free-running (device-timed loop): any slot will do
            slot = self.k_step % self.NS
            self.k_step += 1
        rows = self.r1 - self.r0
        ctx.trace_band(params, self.mode, self.c0, self.r1, self.color.data_ptr(), self.glyph.data_ptr() if self.gl else 0, flags)
        skip = (self.r0 - self.c0) * W
        ctx.encode_band(self.color.data_ptr() + skip * bpp, (self.glyph.data_ptr() + skip) if self.gl else 0, self.x, rows,
                        self.mode, self.r0 > 0, self.out[slot].data_ptr(), self.band_cap, self.total[slot:].data_ptr())
        return slot

    def submit(self, params, flags=0):
        import time
        torch = self.torch
        if self.k_sub - self.k_col >= self.NS:
            raise RuntimeError("%d frames are already in flight: collect one first" % self.NS)
        t0 = time.perf_counter()
        slot = self.k_sub % self.NS                # frame j of the submit/collect sequence lives in slot j % NS
        self.k_sub += 1
        self.step(params, flags, slot)
        self.h_total[slot:slot + 1].copy_(self.total[slot:slot + 1], non_blocking=True)
        self.done_ev[slot].record(torch.cuda.current_stream())
        self.t_submit += time.perf_counter() - t0
        return slot

    def _spin(self, cond):
        import time
        t0 = time.perf_counter()
        while not cond():
            if time.perf_counter() - t0 > 60.0:
                raise RuntimeError("rank %d: timed out waiting for a peer in the shared frame header" % self.rank)

    def _issue(self, j):
        """Frame j: wait for its kernels, publish its length, find its offset, start its copy into the shared frame."""
        import time
        torch = self.torch
        slot, tag, g = j % self.NS, j + 1, self.rank
        t0 = time.perf_counter()
        self.done_ev[slot].synchronize()
        t1 = time.perf_counter()
        n = int(self.h_total[slot])
        self.lens[slot][g] = n
        self.len_tag[slot][g] = tag
        self._spin(lambda: all(self.len_tag[slot][h] >= tag for h in range(g)))
        off = int(sum(int(self.lens[slot][h]) for h in range(g)))
        if j >= self.NS and g != 0:                # frame j - NS used this slot: wait until rank 0 has released it
            self._spin(lambda: self.released[0] >= tag - self.NS)
        t2 = time.perf_counter()
        if n:
            with torch.cuda.stream(self.copy_stream):
                self.frames[slot][off:off + n].copy_(self.out[slot][:n], non_blocking=True)
        self.copy_ev[slot].record(self.copy_stream)
        self.t_wait_gpu += t1 - t0
        self.t_wait_len += t2 - t1
        self.k_issue = j + 1

    def collect(self):
        import time
        if self.k_col >= self.k_sub:
            raise RuntimeError("no frame in flight")
        j = self.k_col
        slot, tag, g = j % self.NS, j + 1, self.rank
        if self.k_issue <= j:
            self._issue(j)
        t0 = time.perf_counter()
        self.copy_ev[slot].synchronize()
        t1 = time.perf_counter()
        self.done_tag[slot][g] = tag
        view, total = None, 0
        if g == 0:
            self._spin(lambda: all(self.done_tag[slot][h] >= tag for h in range(self.world)))
            total = int(sum(int(self.lens[slot][h]) for h in range(self.world)))
            view = self.frames[slot][:total]
            self.released[0] = max(int(self.released[0]), tag - 2)    # frames up to j - 2 may be overwritten from now on
        self.t_copy += t1 - t0
        self.t_wait_done += time.perf_counter() - t1
        self.k_col = j + 1
        # Three frames in flight: start the next frame's copy now, so that it runs under the next submit and the
        # kernels queued behind it (with fewer frames in flight this would only stall the caller).
        if self.k_sub - self.k_issue >= 2:
            self._issue(self.k_issue)
        return view, total

    def close(self):
        if getattr(self, "_addr", None) is not None:
            try:
                self.torch.cuda.cudart().cudaHostUnregister(self._addr)
            except Exception:
                pass
            self._addr = None
        self.lens = self.len_tag = self.done_tag = self.released = self.frames = None
        if getattr(self, "shm", None) is not None:
            try:
                self.shm.close()
                if self.rank == 0:
                    self.shm.unlink()
            except Exception:
                pass
            self.shm = None
