// rtc_encode.cu -- kernel 3: warp-cooperative ANSI encoder (HBM-bound).
//
// Replaces the reference's host-side, serial, byte-at-a-time MinimizeRGB / Minimize8bit
// (RayTracingManager.cu:251-319 / :181-249) AND the 20/12-byte cell formatting at the end of
// every RayTrace_* kernel (RayTracing.cu:231-251, :312-331, :448-471, :585-608, :727-750).
//
// Input : quantised colour plane (3 B/cell RGB, or 1 B/cell xterm index), optional glyph plane,
//         cells in raster order, W = x-1 per row, no padding.
// Output: the minimised stream: a cell emits its full escape sequence (20 or 12 bytes, NUL
//         padded digits included) iff its colour key differs from the previous traced cell in
//         raster order (carried across rows; the very first cell always emits), else only its
//         character; one '\n' after each row.  (Proven byte-identical to the reference's scan
//         by tests)
//
// Two launches, no inter-CTA waiting (a fused single-pass version with decoupled look-back -- per slice, per
// tile, and two-level -- was measured at 0.32-0.44 ms against 0.18 ms: with the store stream saturating HBM the
// descriptor round trips are exposed; the same kernel with the look-back skipped ran in 0.152 ms.  Kept under
// scripts/experiments/; numbers in profiles/r01_encode_history.md):
//   1. count : a tile is 1280 cells = 8 warp slices of 160; writes the tile's emitted byte count and the
//              exclusive offset of each of its 8 slices (a lane counts 20 consecutive cells here);
//              and accumulates group / super-group totals from which any tile's offset is three loads away;
//   2. emit  : every WARP is autonomous (no CTA barrier after the LUT is staged): it re-derives its
//              lanes' lengths, prefix-sums them with shuffles, and each lane streams its cells through a
//              4-byte shift register into a shared-memory image of the warp's slice of the stream.  A full
//              cell is exactly 5 (3) words, so the byte phase only moves on 1-byte cells; every store is
//              a whole aligned STS.32 -- a lane's leading partial word is completed with the trailing
//              bytes of its left neighbour, passed by one shuffle (a lane always owns >= 5 bytes, so a
//              word never spans three lanes).  5 cells per lane makes the lane stride odd (25 / 15 words
//              when every cell is full, the worst case), hence bank-conflict free.  The slice then leaves
//              through ONE TMA bulk store (cp.async.bulk shared -> global; the image is phase-aligned with the
//              global offset) plus <= 15 head and tail bytes.
// Algorithmic traffic: BPP (+1) bytes read and the emitted bytes written per cell.
// (r01a design -- 4 lane-strided cells per thread, byte-granular predicated stores, 3 CTA barriers -- cost
// 198 thread instructions per cell and was issue-bound at 30 % of HBM peak: profiles/r01a_encode_3pass_ncu.md.)
#include <algorithm>

#include "rtc_device.cuh"
#include "rtc_kernels.h"

namespace rtc {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncC = 5;                                       // consecutive cells per lane
constexpr int kEncWarpCells = 32 * kEncC;                      // 160
constexpr int kEncTile = kEncWarps * kEncWarpCells;            // 1280 cells per CTA
// staging image of one warp slice: 160 cells x (20 + 1 newline) + 15 phase bytes, whole words, 16 B multiple
constexpr int kEncStageBytes = ((kEncWarpCells * 21 + 16 + 8) + 15) & ~15;

// NUL-padded 3 decimal digits of v (RayTracing.cu:526-543) packed as D2 | D1<<8 | D0<<16 | ';'<<24.
// The ';' rides along so that one PRMT assembles "D1 D0 ; D2'" words of the cell.
__host__ __device__ constexpr uint32_t digits_entry(uint32_t v)
{
    return (v >= 100u ? 48u + v / 100u : 0u) | ((v >= 10u ? 48u + (v / 10u) % 10u : 0u) << 8) | ((48u + v % 10u) << 16) | (59u << 24);
}
struct DigitLut { uint32_t v[256]; };
constexpr DigitLut make_digit_lut()
{
    DigitLut t{};
    for (uint32_t i = 0; i < 256u; ++i) t.v[i] = digits_entry(i);
    return t;
}
__device__ const DigitLut d_digit_lut = make_digit_lut();

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)   // PRMT without __byte_perm's selector mask
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// ---- per-lane cell analysis, shared by the count and emit passes -------------------------------
// Launch-invariant geometry, computed once on the host.
struct EncGeom {
    uint32_t W, magic, shift;        // exact n / W for n < 2^31 by one widening multiply (round-up method:
                                     // magic = ceil(2^shift / W), shift = 31 + ceil(log2 W); the error magic*W - 2^shift
                                     // is < W <= 2^(shift-31), so n * error < 2^shift for every n < 2^31)
    uint32_t n_cells;
    uint32_t fast_hi_emit, fast_hi_count;   // lanes with cell-1 < fast_hi may load their key window unclamped
    uint32_t force_first;            // 1: cell 0 starts the frame and always emits; 0: the plane continues a frame and
                                     //    the cell stored just before it is cell 0's predecessor (row-band encoding)
    unsigned long long first_w, last_w;     // first / last 32-bit word holding plane bytes
};
template <int BPP, int C> struct EncIn {
    static constexpr int NIN = ((C + 1) * BPP + 3) / 4 + 1;                // aligned words of key bytes + 1 for the phase
    static constexpr int MARGIN = (4 * NIN - BPP + BPP - 1) / BPP;         // cell + MARGIN <= n_cells: window inside the plane
};
__device__ __forceinline__ uint32_t row_of(const EncGeom& g, uint32_t n)
{
    return (uint32_t)(((unsigned long long)n * g.magic) >> g.shift);
}

// Colour keys of this lane's C consecutive cells (key[1..C]) and of the cell before them (key[0]) from aligned
// 32-bit loads around an arbitrarily aligned plane.  Words outside the plane are never touched.
// L2 residency hints: the count pass reads the plane with evict_last, the emit pass writes the stream with
// evict_first, so that the emit pass finds (most of) a <= 126 MB plane still in L2 instead of re-reading HBM.
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <bool KEEP>
__device__ __forceinline__ uint32_t ld_plane(const uint32_t* p, unsigned long long policy)
{
    uint32_t v;
    if (KEEP) asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    else v = __ldg(p);
    return v;
}

template <int BPP, int C, bool KEEP>
__device__ __forceinline__ void load_keys(const uint8_t* __restrict__ color, const EncGeom& g, uint32_t fast_hi, uint32_t cell,
                                          uint32_t (&key)[C + 1])
{
    const unsigned long long policy = KEEP ? l2_policy_evict_last() : 0ull;
    constexpr int NIN = EncIn<BPP, C>::NIN;
    const uintptr_t a = reinterpret_cast<uintptr_t>(color) + (size_t)cell * BPP - BPP;   // predecessor key (unused for cell 0)
    const uintptr_t wa = a & ~(uintptr_t)3;
    const uint32_t sh = 8u * (uint32_t)(a & 3u);
    uint32_t w[NIN];
    if (cell - 1u < fast_hi) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(wa);
#pragma unroll
        for (int j = 0; j < NIN; ++j) w[j] = ld_plane<KEEP>(p + j, policy);
    } else {                                                    // first / last lanes of the frame
#pragma unroll
        for (int j = 0; j < NIN; ++j) {
            uintptr_t q = wa + 4u * j;
            q = q < (uintptr_t)g.first_w ? (uintptr_t)g.first_w : q;
            q = q > (uintptr_t)g.last_w ? (uintptr_t)g.last_w : q;
            w[j] = __ldg(reinterpret_cast<const uint32_t*>(q));
        }
    }
    uint32_t A[NIN - 1];                                        // the (C+1)*BPP key bytes, byte-aligned
#pragma unroll
    for (int j = 0; j < NIN - 1; ++j) A[j] = __funnelshift_r(w[j], w[j + 1], sh);
#pragma unroll
    for (int i = 0; i <= C; ++i) {
        if (BPP == 3) {
            const int off = 3 * i, wd = off >> 2, s = off & 3;
            key[i] = s == 0 ? (A[wd] & 0xffffffu) : s == 1 ? (A[wd] >> 8) : (__funnelshift_r(A[wd], A[wd + 1], 8 * s) & 0xffffffu);
        } else {
            key[i] = (A[i >> 2] >> (8 * (i & 3))) & 0xffu;
        }
    }
}

// Bit i of the result: cell i emits its whole escape sequence (its colour key differs from its predecessor's).
template <int C>
__device__ __forceinline__ uint32_t full_cells(const uint32_t (&key)[C + 1], uint32_t cell, int n_valid, uint32_t force_first)
{
    uint32_t fm = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) fm |= (key[i + 1] != key[i]) ? (1u << i) : 0u;
    fm |= (cell == 0u && force_first) ? 1u : 0u;                // first cell of the frame always emits
    return fm & ((1u << n_valid) - 1u);
}

// ---- pass 1: per-tile byte counts + per-slice offsets inside the tile ---------------------------
// Counting does not need the emit pass's 5-cells-per-lane layout, only its slice boundaries: here a lane owns
// kCntC = 20 consecutive cells (8 lanes per 160-cell slice, 4 slices per warp, 4 tiles per CTA), which
// amortises the address arithmetic over 4x the cells.
constexpr int kGroupShift = 5, kSuperShift = 10;              // 32 tiles per group, 32 groups per super-group
constexpr int kCntC = 4 * kEncC;
constexpr int kCntTiles = kEncThreads * kCntC / kEncTile;      // tiles per count CTA
static_assert(kCntC * 8 == kEncWarpCells && kCntTiles * kEncTile == kEncThreads * kCntC && (32 % kCntTiles) == 0, "count layout");
template <int BPP>
__global__ void __launch_bounds__(kEncThreads)
count_kernel(const uint8_t* __restrict__ color, const EncGeom g, uint32_t* __restrict__ tile_len, uint32_t* __restrict__ warp_excl,
             uint32_t* __restrict__ group_len, uint32_t* __restrict__ super_len, uint32_t* __restrict__ zero_next, uint32_t zero_n)
{
    constexpr uint32_t CS = BPP == 3 ? 20u : 12u;               // SIZE_RGB / SIZE_8BIT (RayTracing.h:120-123)
    __shared__ uint32_t s_len[kEncThreads / 8];                 // one per slice
    const int tid = threadIdx.x;
    // the accumulators of the NEXT launch (other parity): nobody reads or writes them during this launch
    if (blockIdx.x == 0) for (uint32_t i = tid; i < zero_n; i += kEncThreads) zero_next[i] = 0u;
    const uint32_t cell = (blockIdx.x * (uint32_t)kEncThreads + (uint32_t)tid) * kCntC;
    const int n_valid = cell >= g.n_cells ? 0 : (int)min((uint32_t)kCntC, g.n_cells - cell);
    uint32_t len = 0;
    if (n_valid > 0) {
        uint32_t key[kCntC + 1];
        load_keys<BPP, kCntC, true>(color, g, g.fast_hi_count, cell, key);
        const uint32_t fm = full_cells<kCntC>(key, cell, n_valid, g.force_first);
        const uint32_t newlines = row_of(g, cell + (uint32_t)n_valid) - row_of(g, cell);   // row ends in [cell, cell + n_valid)
        len = (uint32_t)n_valid + (CS - 1u) * __popc(fm) + newlines;
    }
    len += __shfl_xor_sync(0xffffffffu, len, 1);
    len += __shfl_xor_sync(0xffffffffu, len, 2);
    len += __shfl_xor_sync(0xffffffffu, len, 4);
    if ((tid & 7) == 0) s_len[tid >> 3] = len;
    __syncthreads();
    if (tid < kEncThreads / 8) {                                // one thread per slice
        const uint32_t tile = blockIdx.x * (uint32_t)kCntTiles + ((uint32_t)tid >> 3);
        if ((uint64_t)tile * kEncTile < g.n_cells) {
            const int s0 = tid & ~7;
            uint32_t excl = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kEncWarps; ++w) {
                const uint32_t v = s_len[s0 + w];
                excl += w < (tid & 7) ? v : 0u;
                tot += v;
            }
            warp_excl[(size_t)tile * kEncWarps + (tid & 7)] = excl;
            if ((tid & 7) == 0) tile_len[tile] = tot;
        }
    }
    if (tid == 0) {                                             // the CTA's 4 tiles share one group and one super-group
        uint32_t cta = 0;
#pragma unroll
        for (int w = 0; w < kEncThreads / 8; ++w) cta += s_len[w];
        const uint32_t tile0 = blockIdx.x * (uint32_t)kCntTiles;
        if (cta) {
            atomicAdd(group_len + (tile0 >> kGroupShift), cta);
            atomicAdd(super_len + (tile0 >> kSuperShift), cta);
        }
    }
}

// ---- (no scan pass) ---------------------------------------------------------------------------
// The count pass also accumulates the byte counts of 32-tile groups and of 32-group super-groups (two atomics per
// count CTA).  An emit warp then obtains its tile's stream offset as
//     sum(super-groups before mine) + sum(earlier groups of my super-group) + sum(earlier tiles of my group)
// = three loads per lane + one warp reduction, with no inter-CTA dependency and no third launch.  (A one-CTA scan
// kernel between the two passes cost 13.5 us however it was written -- row-per-warp shuffles or 32 consecutive
// tiles per thread -- plus a launch gap: profiles/r01_encode_history.md.)

// ---- pass 2: emit -----------------------------------------------------------------------------
// One cell into the lane's byte stream.  `acc` holds the last 4 stream bytes, `sel` = 0x7654 - 0x1111*k encodes
// the k pending (not yet stored) bytes -- the top k bytes of acc -- as the PRMT selector that splices them in
// front of the next word.
__device__ __forceinline__ void put_byte(uint32_t*& wp, uint32_t& acc, uint32_t& sel, uint32_t ch)
{
    acc = prmt(acc, ch, 0x4321u);
    sel -= 0x1111u;
    if (sel == 0x3210u) { *wp++ = acc; sel = 0x7654u; }
}

// SLOW: lanes with fewer than kEncC valid cells or with a row end among their cells.
template <int BPP, bool GLYPH, bool SLOW>
__device__ __forceinline__ void emit_cells(const uint32_t* __restrict__ s_lut, const uint8_t* __restrict__ glyph, uint32_t cell,
                                           int n_valid, const uint32_t (&key)[kEncC + 1], uint32_t fm, uint32_t nm,
                                           uint32_t*& wp, uint32_t& acc, uint32_t& sel)
{
    constexpr int NWC = BPP == 3 ? 5 : 3;                       // words per full cell
#pragma unroll
    for (int i = 0; i < kEncC; ++i) {
        if (SLOW && i >= n_valid) break;
        const uint32_t g = GLYPH ? (uint32_t)__ldg(glyph + cell + i) : 32u;
        if ((fm >> i) & 1u) {
            const uint32_t fg = (GLYPH && g != 32u) ? (uint32_t)'3' : (uint32_t)'4';   // fg for an ASCII-mode hit
            const uint32_t mch = 'm' | (g << 8);
            const uint32_t k = key[i + 1];
            uint32_t c[NWC];
            c[0] = 0x1bu | ('[' << 8) | (fg << 16) | ('8' << 24);
            if (BPP == 3) {
                // ESC [ S 8 | ; 2 ; R2 | R1 R0 ; G2 | G1 G0 ; B2 | B1 B0 m CH   (RayTracing.cu:585-594)
                const uint32_t lr = s_lut[k & 255u], lg = s_lut[(k >> 8) & 255u], lb = s_lut[k >> 16];
                c[1] = prmt(';' | ('2' << 8) | (';' << 16), lr, 0x4210u);
                c[2 % NWC] = prmt(lr, lg, 0x4321u);
                c[3 % NWC] = prmt(lg, lb, 0x4321u);
                c[NWC - 1] = prmt(lb, mch, 0x5421u);
            } else {
                // ESC [ S 8 | ; 5 ; I2 | I1 I0 m CH                              (RayTracing.cu:231-237)
                const uint32_t li8 = s_lut[k];
                c[1] = prmt(';' | ('5' << 8) | (';' << 16), li8, 0x4210u);
                c[NWC - 1] = prmt(li8, mch, 0x5421u);
            }
            wp[0] = prmt(acc, c[0], sel);
#pragma unroll
            for (int j = 1; j < NWC; ++j) wp[j] = prmt(c[j - 1], c[j], sel);
            acc = c[NWC - 1];
            wp += NWC;
        } else {
            put_byte(wp, acc, sel, g);                          // same colour as the previous cell: character only
        }
        if (SLOW && ((nm >> i) & 1u)) put_byte(wp, acc, sel, (uint32_t)'\n');
    }
}

template <int BPP, bool GLYPH>
__global__ void __launch_bounds__(kEncThreads)
emit_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ glyph, const EncGeom g,
            char* __restrict__ out, unsigned long long cap, const uint32_t* __restrict__ tile_len,
            const uint32_t* __restrict__ warp_excl, const uint32_t* __restrict__ group_len,
            const uint32_t* __restrict__ super_len, unsigned long long* __restrict__ total)
{
    constexpr uint32_t CS = BPP == 3 ? 20u : 12u;
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_lut[tid] = d_digit_lut.v[tid];
    __syncthreads();                                            // the only CTA-wide barrier

    const uint32_t tile = blockIdx.x;
    const uint32_t cell_w0 = tile * (uint32_t)kEncTile + (uint32_t)warp * kEncWarpCells;
    if (cell_w0 >= g.n_cells) return;                           // warp-uniform
    const uint32_t n_here = min((uint32_t)kEncWarpCells, g.n_cells - cell_w0);
    const uint32_t t5 = (uint32_t)lane * kEncC;
    const uint32_t cell = cell_w0 + t5;
    const int n_valid = t5 >= n_here ? 0 : (int)min((uint32_t)kEncC, n_here - t5);
    // stream offset of this warp's slice (see "no scan pass")
    unsigned long long goff = warp_excl[(size_t)tile * kEncWarps + warp];
    {
        const uint32_t grp = tile >> kGroupShift, sup = tile >> kSuperShift;
        const uint32_t gi = (sup << (kSuperShift - kGroupShift)) + (uint32_t)lane, ti = (grp << kGroupShift) + (uint32_t)lane;
        uint32_t v = (gi < grp ? __ldg(group_len + gi) : 0u) + (ti < tile ? __ldg(tile_len + ti) : 0u);
        if ((uint32_t)lane < sup) v += __ldg(super_len + lane);
        goff += __reduce_add_sync(0xffffffffu, v);               // < 32 x 2^25 + 2 x 32 x 2^20: fits 32 bits
        for (uint32_t i = 32u + (uint32_t)lane; i < ((sup + 31u) & ~31u); i += 32u)   // frames beyond 32 super-groups (> 42 Mcells)
            goff += __reduce_add_sync(0xffffffffu, i < sup ? __ldg(super_len + i) : 0u);
    }
    unsigned char* stage = smem + 1024 + warp * kEncStageBytes;

    // ---- keys, lengths, warp prefix sum --------------------------------------------------------
    uint32_t key[kEncC + 1], fm = 0, nm = 0, len = 0;
    if (n_valid > 0) {
        load_keys<BPP, kEncC, false>(color, g, g.fast_hi_emit, cell, key);
        fm = full_cells<kEncC>(key, cell, n_valid, g.force_first);
        const uint32_t col = cell - row_of(g, cell) * g.W;
        if (g.W >= (uint32_t)kEncC) {                           // at most one row end among kEncC consecutive cells
            const uint32_t d = g.W - 1u - col;
            nm = d < (uint32_t)kEncC ? (1u << d) : 0u;
        } else {                                                // very narrow consoles
            uint32_t c = col;
#pragma unroll
            for (int i = 0; i < kEncC; ++i) {
                const bool nl = c == g.W - 1u;
                nm |= nl ? (1u << i) : 0u;
                c = nl ? 0u : c + 1u;
            }
        }
        nm &= (1u << n_valid) - 1u;
        len = (uint32_t)n_valid + (CS - 1u) * __popc(fm) + __popc(nm);
    }
    uint32_t inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t warp_len = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 0 && cell_w0 + n_here == g.n_cells) *total = goff + warp_len;   // the frame's last slice

    // ---- stream the cells into the staging image (phase-aligned with the output) ---------------
    const uint32_t out_phase = (uint32_t)(reinterpret_cast<uintptr_t>(out + goff) & 15u);
    const uint32_t pos = out_phase + (inc - len);
    const uint32_t k0 = pos & 3u;                               // bytes of my first word that belong to my left neighbour
    uint32_t* const wp0 = reinterpret_cast<uint32_t*>(stage) + (pos >> 2);
    uint32_t* wp = wp0;
    uint32_t acc = 0u, sel = 0x7654u - 0x1111u * k0;
    if (n_valid == kEncC && nm == 0u) emit_cells<BPP, GLYPH, false>(s_lut, glyph, cell, n_valid, key, fm, nm, wp, acc, sel);
    else if (n_valid > 0) emit_cells<BPP, GLYPH, true>(s_lut, glyph, cell, n_valid, key, fm, nm, wp, acc, sel);
    // The slice's last lane flushes its pending bytes itself (sel & 7 == 4 - pending) ...
    if (n_valid > 0 && t5 + (uint32_t)n_valid == n_here && sel != 0x7654u) *wp = acc >> (8u * (sel & 7u));
    // ... everyone else's complete the first word of the right neighbour (a lane owns >= 5 bytes, so that word
    // was written -- by the neighbour alone -- with zeros in its low k0 bytes).
    const uint32_t left = __shfl_up_sync(0xffffffffu, acc, 1);
    if (lane > 0 && n_valid > 0 && k0 != 0u) *wp0 |= left >> (8u * (4u - k0));

    // ---- copy the slice out: head bytes, one TMA bulk store of the 16-byte aligned body, tail bytes ---------
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy STS -> visible to the async proxy
    __syncwarp();
    if (goff >= cap) return;
    const uint32_t n_out = (uint32_t)min((unsigned long long)warp_len, cap - goff);
    char* dst = out + goff;
    const unsigned char* src = stage + out_phase;
    const uint32_t head = out_phase ? min(16u - out_phase, n_out) : 0u;
    const uint32_t body = (n_out - head) & ~15u;
    if (lane == 0 && body) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     :: "l"(dst + head), "r"((uint32_t)__cvta_generic_to_shared(src + head)), "r"(body), "l"(l2_policy_evict_first())
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if ((uint32_t)lane < head) dst[lane] = (char)src[lane];
    const uint32_t done = head + body;
    if ((uint32_t)lane < n_out - done) dst[done + lane] = (char)src[done + lane];
    if (lane == 0 && body) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the read
}

template <int BPP>
static EncGeom make_geom(const uint8_t* color, uint32_t W, uint32_t n_cells, bool continues)
{
    EncGeom g;
    uint32_t l = 0;
    while ((1ull << l) < W) ++l;
    g.W = W; g.shift = 31u + l;
    g.magic = (uint32_t)(((1ull << g.shift) + W - 1u) / W);
    g.n_cells = n_cells;
    const uint32_t me = (uint32_t)EncIn<BPP, kEncC>::MARGIN, mc = (uint32_t)EncIn<BPP, kCntC>::MARGIN;
    g.fast_hi_emit = n_cells >= me ? n_cells - me : 0u;
    g.fast_hi_count = n_cells >= mc ? n_cells - mc : 0u;
    const unsigned long long base = (unsigned long long)reinterpret_cast<uintptr_t>(color);
    g.force_first = continues ? 0u : 1u;
    g.first_w = (base - (continues ? (unsigned long long)BPP : 0ull)) & ~3ull;   // the predecessor cell is readable
    g.last_w = (base + (unsigned long long)n_cells * BPP - 1ull) & ~3ull;
    return g;
}

// SDL mode (reference RayTrace_SDL writes nothing, RayTracing.cu:755-795): y newlines.
__global__ void newline_kernel(char* __restrict__ out, uint32_t y, unsigned long long cap, unsigned long long* total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < y && i < cap) out[i] = '\n';
    if (i == 0) *total = y;
}

// ---- compatibility: the reference's RAW cell buffer --------------------------------------------------------------------
// What RayTracing::RayTrace leaves in `resultArray` (RayTracing.cu:585-608 and siblings): one SIZE-byte cell per console
// position at (row * x + col) * SIZE, SIZE = 20 (RGB modes) or 12 (8-bit modes), every traced cell written in full
// (NUL-padded digits), the newline column x-1 left zero.  The product path never materialises this buffer (the encoder
// goes from the planes straight to the minimised stream); it exists for callers that keep the reference's own host-side
// MinimizeRGB / Minimize8bit, and as a parity hook (tests compare it with the reference's raw buffer byte for byte).
template <int BPP, bool GLYPH>
__global__ void __launch_bounds__(256)
expand_raw_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ glyph, uint32_t x, uint32_t y, uint32_t* __restrict__ out)
{
    constexpr int NWC = BPP == 3 ? 5 : 3;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)x * y) return;
    const uint32_t row = (uint32_t)(i / x), col = (uint32_t)(i - (size_t)row * x);
    uint32_t c[NWC];
#pragma unroll
    for (int j = 0; j < NWC; ++j) c[j] = 0u;
    if (col + 1u < x) {
        const size_t cell = (size_t)row * (x - 1u) + col;
        const uint32_t g = GLYPH ? (uint32_t)glyph[cell] : 32u;
        const uint32_t fg = (GLYPH && g != 32u) ? (uint32_t)'3' : (uint32_t)'4';
        const uint32_t mch = 'm' | (g << 8);
        c[0] = 0x1bu | ('[' << 8) | (fg << 16) | ('8' << 24);
        if (BPP == 3) {
            const uint32_t lr = d_digit_lut.v[color[cell * 3]], lg = d_digit_lut.v[color[cell * 3 + 1]], lb = d_digit_lut.v[color[cell * 3 + 2]];
            c[1] = prmt(';' | ('2' << 8) | (';' << 16), lr, 0x4210u);
            c[2 % NWC] = prmt(lr, lg, 0x4321u);
            c[3 % NWC] = prmt(lg, lb, 0x4321u);
            c[NWC - 1] = prmt(lb, mch, 0x5421u);
        } else {
            const uint32_t li8 = d_digit_lut.v[color[cell]];
            c[1] = prmt(';' | ('5' << 8) | (';' << 16), li8, 0x4210u);
            c[NWC - 1] = prmt(li8, mch, 0x5421u);
        }
    }
#pragma unroll
    for (int j = 0; j < NWC; ++j) out[i * NWC + j] = c[j];
}

cudaError_t launch_expand_raw(cudaStream_t st, const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y, int mode, char* out)
{
    const size_t n = (size_t)x * y;
    if (n == 0) return cudaSuccess;
    if (reinterpret_cast<uintptr_t>(out) & 3u) return cudaErrorInvalidValue;
    const bool bit8 = mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL;
    const bool gl = (mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII) && glyph != nullptr;
    cudaError_t e;
    if (mode == RTC_SDL) return cudaMemsetAsync(out, 0, 20 * n, st);          // RayTrace_SDL writes nothing into the cleared buffer
    if (bit8 && (e = cudaMemsetAsync(out + 12 * n, 0, 8 * n, st)) != cudaSuccess) return e;   // the reference clears all 20*x*y bytes
    const unsigned grid = (unsigned)((n + 255) / 256);
    uint32_t* o = reinterpret_cast<uint32_t*>(out);
    if (bit8) { if (gl) expand_raw_kernel<1, true><<<grid, 256, 0, st>>>(color, glyph, x, y, o); else expand_raw_kernel<1, false><<<grid, 256, 0, st>>>(color, glyph, x, y, o); }
    else      { if (gl) expand_raw_kernel<3, true><<<grid, 256, 0, st>>>(color, glyph, x, y, o); else expand_raw_kernel<3, false><<<grid, 256, 0, st>>>(color, glyph, x, y, o); }
    return cudaGetLastError();
}

static size_t enc_smem() { return 1024 + (size_t)kEncWarps * kEncStageBytes; }

cudaError_t configure_encode()
{
    cudaError_t e;
    const int bytes = (int)enc_smem();
    if ((e = cudaFuncSetAttribute(emit_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(emit_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(emit_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(emit_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    return cudaSuccess;
}

// scratch layout (all u32): [parity 0: group_len, super_len][parity 1: same][tile_len][warp_excl x 8].
// The accumulator capacities follow from the ALLOCATION size, so that the layout (and what "zero everything of the
// other parity" means) does not change between launches of different frame sizes on the same scratch buffer.
static uint32_t scratch_tiles(size_t scratch_bytes)              // tiles a scratch allocation can describe
{
    return scratch_bytes < 1024 ? 0u : (uint32_t)std::min<size_t>((scratch_bytes - 1024) / (4 + 4 * kEncWarps + 1), 0x7fffffffu);
}
size_t encode_state_bytes(uint64_t n_cells)
{
    const uint64_t n_tiles = (n_cells + kEncTile - 1) / kEncTile + 1;
    return (size_t)(n_tiles * (4 + 4 * kEncWarps + 1) + 1024 + 64);
}

cudaError_t launch_encode(cudaStream_t st, const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y,
                          int mode, char* out, size_t cap, unsigned long long* total, void* scratch, size_t scratch_bytes,
                          uint32_t parity, bool continues)
{
    if (mode == RTC_SDL) {
        newline_kernel<<<(y + 255) / 256, 256, 0, st>>>(out, y, cap, total);
        return cudaGetLastError();
    }
    const uint32_t W = x - 1u;
    const uint64_t n_cells64 = (uint64_t)W * y;
    if (W == 0 || y == 0) {
        newline_kernel<<<(y + 255) / 256 + 1, 256, 0, st>>>(out, y, cap, total);   // x == 1: only the newline column exists
        return cudaGetLastError();
    }
    if (n_cells64 >= (1ull << 31)) return cudaErrorInvalidValue;
    const uint32_t n_cells = (uint32_t)n_cells64;
    const uint32_t n_tiles = (n_cells + kEncTile - 1) / kEncTile;
    const uint32_t t_max = scratch_tiles(scratch_bytes);
    if (n_tiles > t_max) return cudaErrorInvalidValue;
    const uint32_t n_grp = (t_max >> kGroupShift) + 1u, n_sup = (t_max >> kSuperShift) + 1u;
    const uint32_t acc_n = (n_grp + n_sup + 3u) & ~3u;           // accumulators per parity
    uint32_t* base = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* group_len = base + (parity & 1u) * acc_n;
    uint32_t* super_len = group_len + n_grp;
    uint32_t* zero_next = base + ((parity & 1u) ^ 1u) * acc_n;
    uint32_t* tile_len = base + 2u * acc_n;
    uint32_t* warp_excl = tile_len + ((t_max + 4u) & ~3u);
    const bool has_glyph = (mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII) && glyph != nullptr;
    const bool bit8 = (mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL);
    const EncGeom g = bit8 ? make_geom<1>(color, W, n_cells, continues) : make_geom<3>(color, W, n_cells, continues);
    const uint32_t n_cnt = (n_tiles + kCntTiles - 1) / kCntTiles;
    if (bit8) count_kernel<1><<<n_cnt, kEncThreads, 0, st>>>(color, g, tile_len, warp_excl, group_len, super_len, zero_next, acc_n);
    else count_kernel<3><<<n_cnt, kEncThreads, 0, st>>>(color, g, tile_len, warp_excl, group_len, super_len, zero_next, acc_n);
#define RTC_LAUNCH_ENC(BPP, GL)                                                                         \
    emit_kernel<BPP, GL><<<n_tiles, kEncThreads, enc_smem(), st>>>(                                     \
        color, glyph, g, out, (unsigned long long)cap, tile_len, warp_excl, group_len, super_len, total)
    if (bit8) { if (has_glyph) RTC_LAUNCH_ENC(1, true); else RTC_LAUNCH_ENC(1, false); }
    else      { if (has_glyph) RTC_LAUNCH_ENC(3, true); else RTC_LAUNCH_ENC(3, false); }
#undef RTC_LAUNCH_ENC
    return cudaGetLastError();
}

}  // namespace rtc
