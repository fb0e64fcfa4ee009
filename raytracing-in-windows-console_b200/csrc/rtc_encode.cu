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
// Three launches, no inter-CTA waiting:
//   1. count : a lane owns 5 CONSECUTIVE cells, a warp 160, a CTA (8 warps) one tile of 1280; writes the
//              tile's emitted byte count and the exclusive offset of each of its 8 warp slices;
//   2. scan  : one CTA turns the tile counts into exclusive 64-bit offsets + the stream length;
//   3. emit  : every WARP is autonomous (no CTA barrier after the LUT is staged): it re-derives its
//              lanes' lengths, prefix-sums them with shuffles, and each lane streams its cells through a
//              4-byte shift register into a shared-memory image of the warp's slice of the stream.  A full
//              cell is exactly 5 (3) words, so the byte phase only moves on 1-byte cells; every store is
//              a whole aligned STS.32 -- a lane's leading partial word is completed with the trailing
//              bytes of its left neighbour, passed by one shuffle (a lane always owns >= 5 bytes, so a
//              word never spans three lanes).  5 cells per lane makes the lane stride odd (25 / 15 words
//              when every cell is full, the worst case), hence bank-conflict free.  The slice is then
//              copied out with coalesced 128-bit stores, phase-aligned with the global offset.
// Algorithmic traffic: BPP (+1) bytes read and the emitted bytes written per cell.
// (r01a design -- 4 lane-strided cells per thread, byte-granular predicated stores, 3 CTA barriers -- cost
// 198 thread instructions per cell and was issue-bound at 30 % of HBM peak: profiles/r01a_encode_3pass_ncu.md.)
#include "rtc_device.cuh"
#include "rtc_kernels.h"

#ifndef RTC_ENC_BULK_STORE
#define RTC_ENC_BULK_STORE 1     // slice copy-out by TMA bulk store (cp.async.bulk) instead of a LDS/STG loop
#endif

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

// ---- per-lane cell analysis, shared by the count and emit passes -------------------------------
// Exact n / W and n % W for n < 2^31 by one widening multiply: magic = ceil(2^shift / W), shift = 31 + ceil(log2 W)
// (round-up method: the error magic*W - 2^shift is < W <= 2^(shift-31), so n * error < 2^shift for n < 2^31).
struct RowDiv { uint32_t W, magic, shift; };
static RowDiv make_rowdiv(uint32_t W)
{
    uint32_t l = 0;
    while ((1ull << l) < W) ++l;
    RowDiv d;
    d.W = W; d.shift = 31u + l;
    d.magic = (uint32_t)(((1ull << d.shift) + W - 1u) / W);
    return d;
}
__device__ __forceinline__ uint32_t row_col(const RowDiv& d, uint32_t n)
{
    const uint32_t q = (uint32_t)(((unsigned long long)n * d.magic) >> d.shift);
    return n - q * d.W;
}

// Colour keys of this lane's kEncC cells (key[1..C]) and of the cell before them (key[0]) from aligned
// 32-bit loads around an arbitrarily aligned plane.  Words outside the plane are never touched.
template <int BPP>
__device__ __forceinline__ void load_keys(const uint8_t* __restrict__ color, uintptr_t first_w, uintptr_t last_w, uintptr_t last_w5, uint32_t cell,
                                          uint32_t (&key)[kEncC + 1])
{
    constexpr int NIN = BPP == 3 ? 6 : 3;                       // words covering (C+1)*BPP bytes at any phase
    const uintptr_t a = reinterpret_cast<uintptr_t>(color) + (size_t)cell * BPP - BPP;   // predecessor key (unused for cell 0)
    const uintptr_t wa = a & ~(uintptr_t)3;
    const uint32_t sh = 8u * (uint32_t)(a & 3u);
    uint32_t w[NIN];
    if (wa >= first_w && wa <= last_w5) {                       // last_w5 = last word of the plane - 4*(NIN-1)
        const uint32_t* p = reinterpret_cast<const uint32_t*>(wa);
#pragma unroll
        for (int j = 0; j < NIN; ++j) w[j] = __ldg(p + j);
    } else {                                                    // first / last lanes of the frame
#pragma unroll
        for (int j = 0; j < NIN; ++j) {
            uintptr_t q = wa + 4u * j;
            q = q < first_w ? first_w : q;
            q = q > last_w ? last_w : q;
            w[j] = __ldg(reinterpret_cast<const uint32_t*>(q));
        }
    }
    uint32_t A[NIN - 1];                                        // the (C+1)*BPP key bytes, byte-aligned
#pragma unroll
    for (int j = 0; j < NIN - 1; ++j) A[j] = __funnelshift_r(w[j], w[j + 1], sh);
#pragma unroll
    for (int i = 0; i <= kEncC; ++i) {
        if (BPP == 3) {
            const int off = 3 * i, wd = off >> 2, s = off & 3;
            key[i] = s == 0 ? (A[wd] & 0xffffffu) : s == 1 ? (A[wd] >> 8) : (__funnelshift_r(A[wd], A[wd + 1], 8 * s) & 0xffffffu);
        } else {
            key[i] = (A[i >> 2] >> (8 * (i & 3))) & 0xffu;
        }
    }
}
// Plane bounds for load_keys (word addresses; the plane has at least one byte).
template <int BPP>
__device__ __forceinline__ void plane_words(const uint8_t* color, uint32_t n_cells, uintptr_t& first_w, uintptr_t& last_w, uintptr_t& last_w5)
{
    constexpr int NIN = BPP == 3 ? 6 : 3;
    const uintptr_t base = reinterpret_cast<uintptr_t>(color);
    first_w = base & ~(uintptr_t)3;
    last_w = (base + (size_t)n_cells * BPP - 1) & ~(uintptr_t)3;
    // frames smaller than one lane's window always take the clamped path (first_w > last_w5)
    last_w5 = last_w >= first_w + 4u * (NIN - 1) ? last_w - 4u * (NIN - 1) : first_w - 4u;
}

// Bit i of full_mask: cell i emits its whole escape sequence; bit i of nl_mask: cell i ends a row.
// Both restricted to the n_valid leading cells.  Returns the lane's emitted byte count.
template <int BPP>
__device__ __forceinline__ uint32_t lane_layout(const uint32_t (&key)[kEncC + 1], uint32_t cell, int n_valid, const RowDiv& rd,
                                                uint32_t& full_mask, uint32_t& nl_mask)
{
    constexpr uint32_t CS = BPP == 3 ? 20u : 12u;               // SIZE_RGB / SIZE_8BIT (RayTracing.h:120-123)
    uint32_t fm = 0;
#pragma unroll
    for (int i = 0; i < kEncC; ++i) fm |= (key[i + 1] != key[i]) ? (1u << i) : 0u;
    fm |= cell == 0u ? 1u : 0u;                                 // first cell of the frame always emits
    const uint32_t col = row_col(rd, cell);
    uint32_t nm;
    if (rd.W >= (uint32_t)kEncC) {                              // at most one row end among 5 consecutive cells
        const uint32_t d = rd.W - 1u - col;
        nm = d < (uint32_t)kEncC ? (1u << d) : 0u;
    } else {                                                    // very narrow consoles
        nm = 0;
        uint32_t c = col;
#pragma unroll
        for (int i = 0; i < kEncC; ++i) {
            const bool nl = c == rd.W - 1u;
            nm |= nl ? (1u << i) : 0u;
            c = nl ? 0u : c + 1u;
        }
    }
    const uint32_t vm = (1u << n_valid) - 1u;
    full_mask = fm & vm;
    nl_mask = nm & vm;
    return (uint32_t)n_valid + (CS - 1u) * __popc(full_mask) + __popc(nl_mask);
}

// ---- pass 1: per-tile byte counts + per-warp offsets inside the tile ---------------------------
template <int BPP>
__global__ void __launch_bounds__(kEncThreads)
count_kernel(const uint8_t* __restrict__ color, const RowDiv rd, uint32_t n_cells, uint32_t* __restrict__ tile_len,
             uint32_t* __restrict__ warp_excl)
{
    __shared__ uint32_t s_sum[kEncWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t cell = tile * (uint32_t)kEncTile + (uint32_t)warp * kEncWarpCells + (uint32_t)lane * kEncC;
    const int n_valid = cell >= n_cells ? 0 : (int)min((uint32_t)kEncC, n_cells - cell);
    uint32_t len = 0;
    if (n_valid > 0) {
        uint32_t key[kEncC + 1], fm, nm;
        uintptr_t first_w, last_w, last_w5;
        plane_words<BPP>(color, n_cells, first_w, last_w, last_w5);
        load_keys<BPP>(color, first_w, last_w, last_w5, cell, key);
        len = lane_layout<BPP>(key, cell, n_valid, rd, fm, nm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
    if (lane == 0) s_sum[warp] = len;
    __syncthreads();
    if (tid < kEncWarps) {
        uint32_t excl = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kEncWarps; ++w) {
            const uint32_t v = s_sum[w];
            excl += w < tid ? v : 0u;
            tot += v;
        }
        warp_excl[(size_t)tile * kEncWarps + tid] = excl;
        if (tid == 0) tile_len[tile] = tot;
    }
}

// ---- pass 2: exclusive scan of the tile counts (one CTA) ------------------------------------
// Chunks of 1024 x 16 tiles; a warp owns 512 consecutive tiles and walks them in 16 coalesced rows of 32
// (loads issued up front), so tile_len is read and tile_off written in whole 128/256-byte lines.
// A chunk's sum fits 32 bits (16384 tiles x <= 26880 bytes).
constexpr int kScanRows = 16;
__global__ void __launch_bounds__(1024)
scan_kernel(const uint32_t* __restrict__ tile_len, uint32_t n_tiles, unsigned long long* __restrict__ tile_off,
            unsigned long long* __restrict__ total)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long carry = 0ull;
    for (uint32_t base = 0; base < n_tiles; base += 1024u * kScanRows) {
        const uint32_t a = base + warp * (32u * kScanRows) + lane;
        uint32_t v[kScanRows];
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) v[k] = (a + 32u * k < n_tiles) ? tile_len[a + 32u * k] : 0u;
        uint32_t run = 0;                                          // exclusive offset inside the warp's 512 tiles
#pragma unroll
        for (int k = 0; k < kScanRows; ++k) {
            uint32_t inc = v[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += t;
            }
            const uint32_t row_total = __shfl_sync(0xffffffffu, inc, 31);
            v[k] = run + inc - v[k];
            run += row_total;
        }
        if (lane == 0u) s_warp[warp] = run;
        __syncthreads();
        uint32_t wsum = s_warp[lane];                              // every warp scans the 32 warp sums itself
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wsum, o);
            if (lane >= (uint32_t)o) wsum += t;
        }
        const uint32_t before = __shfl_sync(0xffffffffu, wsum, (warp + 31u) & 31u);   // inclusive sum of warps < warp
        const uint32_t chunk_total = __shfl_sync(0xffffffffu, wsum, 31);
        const unsigned long long wbase = carry + (warp ? before : 0u);
#pragma unroll
        for (int k = 0; k < kScanRows; ++k)
            if (a + 32u * k < n_tiles) tile_off[a + 32u * k] = wbase + v[k];
        carry += chunk_total;
        __syncthreads();                                           // s_warp is rewritten by the next chunk
    }
    if (tid == 0) *total = carry;
}

// ---- pass 3: emit -----------------------------------------------------------------------------
// One cell into the lane's byte stream.  `acc` holds the last 4 stream bytes, `sel` = 0x7654 - 0x1111*k encodes
// the k pending (not yet stored) bytes -- the top k bytes of acc -- as the PRMT selector that splices them in
// front of the next word.
__device__ __forceinline__ void put_byte(uint32_t*& wp, uint32_t& acc, uint32_t& sel, uint32_t ch)
{
    acc = __byte_perm(acc, ch, 0x4321);
    sel -= 0x1111u;
    if (sel == 0x3210u) { *wp++ = acc; sel = 0x7654u; }
}

template <int BPP, bool GLYPH, bool PARTIAL>
__device__ __forceinline__ void emit_cells(const uint32_t* __restrict__ s_lut, const uint8_t* __restrict__ glyph, uint32_t cell,
                                           int n_valid, const uint32_t (&key)[kEncC + 1], uint32_t fm, uint32_t nm,
                                           uint32_t*& wp, uint32_t& acc, uint32_t& sel)
{
    constexpr int NWC = BPP == 3 ? 5 : 3;                       // words per full cell
#pragma unroll
    for (int i = 0; i < kEncC; ++i) {
        if (PARTIAL && i >= n_valid) break;
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
                c[1] = __byte_perm(';' | ('2' << 8) | (';' << 16), lr, 0x4210);
                c[2 % NWC] = __byte_perm(lr, lg, 0x4321);
                c[3 % NWC] = __byte_perm(lg, lb, 0x4321);
                c[NWC - 1] = __byte_perm(lb, mch, 0x5421);
            } else {
                // ESC [ S 8 | ; 5 ; I2 | I1 I0 m CH                              (RayTracing.cu:231-237)
                const uint32_t li8 = s_lut[k];
                c[1] = __byte_perm(';' | ('5' << 8) | (';' << 16), li8, 0x4210);
                c[NWC - 1] = __byte_perm(li8, mch, 0x5421);
            }
            wp[0] = __byte_perm(acc, c[0], sel);
#pragma unroll
            for (int j = 1; j < NWC; ++j) wp[j] = __byte_perm(c[j - 1], c[j], sel);
            acc = c[NWC - 1];
            wp += NWC;
        } else {
            put_byte(wp, acc, sel, g);                          // same colour as the previous cell: character only
        }
        if ((nm >> i) & 1u) put_byte(wp, acc, sel, (uint32_t)'\n');
    }
}

template <int BPP, bool GLYPH>
__global__ void __launch_bounds__(kEncThreads)
emit_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ glyph, const RowDiv rd, uint32_t n_cells,
            char* __restrict__ out, unsigned long long cap, const unsigned long long* __restrict__ tile_off,
            const uint32_t* __restrict__ warp_excl)
{
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_lut[tid] = d_digit_lut.v[tid];
    __syncthreads();                                            // the only CTA-wide barrier

    const uint32_t tile = blockIdx.x;
    const uint32_t cell_w0 = tile * (uint32_t)kEncTile + (uint32_t)warp * kEncWarpCells;
    if (cell_w0 >= n_cells) return;                             // warp-uniform
    const uint32_t n_here = min((uint32_t)kEncWarpCells, n_cells - cell_w0);
    const uint32_t t5 = (uint32_t)lane * kEncC;
    const uint32_t cell = cell_w0 + t5;
    const int n_valid = t5 >= n_here ? 0 : (int)min((uint32_t)kEncC, n_here - t5);
    const unsigned long long goff = tile_off[tile] + warp_excl[(size_t)tile * kEncWarps + warp];
    unsigned char* stage = smem + 1024 + warp * kEncStageBytes;

    // ---- keys, lengths, warp prefix sum --------------------------------------------------------
    uint32_t key[kEncC + 1], fm = 0, nm = 0, len = 0;
    if (n_valid > 0) {
        uintptr_t first_w, last_w, last_w5;
        plane_words<BPP>(color, n_cells, first_w, last_w, last_w5);
        load_keys<BPP>(color, first_w, last_w, last_w5, cell, key);
        len = lane_layout<BPP>(key, cell, n_valid, rd, fm, nm);
    }
    uint32_t inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t warp_len = __shfl_sync(0xffffffffu, inc, 31);

    // ---- stream the cells into the staging image (phase-aligned with the output) ---------------
    const uint32_t out_phase = (uint32_t)(reinterpret_cast<uintptr_t>(out + goff) & 15u);
    const uint32_t pos = out_phase + (inc - len);
    const uint32_t k0 = pos & 3u;                               // bytes of my first word that belong to my left neighbour
    uint32_t* const wp0 = reinterpret_cast<uint32_t*>(stage) + (pos >> 2);
    uint32_t* wp = wp0;
    uint32_t acc = 0u, sel = 0x7654u - 0x1111u * k0;
    if (n_valid == kEncC) emit_cells<BPP, GLYPH, false>(s_lut, glyph, cell, n_valid, key, fm, nm, wp, acc, sel);
    else if (n_valid > 0) emit_cells<BPP, GLYPH, true>(s_lut, glyph, cell, n_valid, key, fm, nm, wp, acc, sel);
    // The slice's last lane flushes its pending bytes itself (sel & 7 == 4 - pending) ...
    if (n_valid > 0 && t5 + (uint32_t)n_valid == n_here && sel != 0x7654u) *wp = acc >> (8u * (sel & 7u));
    // ... everyone else's complete the first word of the right neighbour (a lane owns >= 5 bytes, so that word
    // was written -- by the neighbour alone -- with zeros in its low k0 bytes).
    const uint32_t left = __shfl_up_sync(0xffffffffu, acc, 1);
    if (lane > 0 && n_valid > 0 && k0 != 0u) *wp0 |= left >> (8u * (4u - k0));
    __syncwarp();

    // ---- copy the slice out ---------------------------------------------------------------------
    if (goff >= cap) return;
    const uint32_t n_out = (uint32_t)min((unsigned long long)warp_len, cap - goff);
    char* dst = out + goff;
    const unsigned char* src = stage + out_phase;
    const uint32_t head = out_phase ? min(16u - out_phase, n_out) : 0u;
    if ((uint32_t)lane < head) dst[lane] = (char)src[lane];
    const uint32_t nvec = (n_out - head) >> 4;
#if RTC_ENC_BULK_STORE
    // TMA bulk store of the 16-byte aligned body: one instruction instead of a LDS.128/STG.128 loop.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0 && nvec) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst + head), "r"((uint32_t)__cvta_generic_to_shared(src + head)), "r"(nvec << 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#else
    uint4* vdst = reinterpret_cast<uint4*>(dst + head);
    const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
#pragma unroll 1
    for (uint32_t i = lane; i < nvec; i += 32u) vdst[i] = vsrc[i];
#endif
    const uint32_t done = head + (nvec << 4);
    if ((uint32_t)lane < n_out - done) dst[done + lane] = (char)src[done + lane];
#if RTC_ENC_BULK_STORE
    if (lane == 0 && nvec) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the read
#endif
}

// SDL mode (reference RayTrace_SDL writes nothing, RayTracing.cu:755-795): y newlines.
__global__ void newline_kernel(char* __restrict__ out, uint32_t y, unsigned long long cap, unsigned long long* total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < y && i < cap) out[i] = '\n';
    if (i == 0) *total = y;
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

// scratch: per tile one u64 offset, one u32 count, 8 u32 warp offsets
size_t encode_state_bytes(uint64_t n_cells)
{
    const uint64_t n_tiles = (n_cells + kEncTile - 1) / kEncTile + 1;
    return (size_t)(n_tiles * (8 + 4 + 4 * kEncWarps) + 64);
}

cudaError_t launch_encode(cudaStream_t st, const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y,
                          int mode, char* out, size_t cap, unsigned long long* total, void* scratch)
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
    unsigned long long* tile_off = reinterpret_cast<unsigned long long*>(scratch);
    uint32_t* tile_len = reinterpret_cast<uint32_t*>(tile_off + n_tiles + 1);
    uint32_t* warp_excl = tile_len + n_tiles + 1;
    const bool has_glyph = (mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII) && glyph != nullptr;
    const bool bit8 = (mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL);
    const RowDiv rd = make_rowdiv(W);
    if (bit8) count_kernel<1><<<n_tiles, kEncThreads, 0, st>>>(color, rd, n_cells, tile_len, warp_excl);
    else count_kernel<3><<<n_tiles, kEncThreads, 0, st>>>(color, rd, n_cells, tile_len, warp_excl);
    scan_kernel<<<1, 1024, 0, st>>>(tile_len, n_tiles, tile_off, total);
#define RTC_LAUNCH_ENC(BPP, GL)                                                                         \
    emit_kernel<BPP, GL><<<n_tiles, kEncThreads, enc_smem(), st>>>(                                     \
        color, glyph, rd, n_cells, out, (unsigned long long)cap, tile_off, warp_excl)
    if (bit8) { if (has_glyph) RTC_LAUNCH_ENC(1, true); else RTC_LAUNCH_ENC(1, false); }
    else      { if (has_glyph) RTC_LAUNCH_ENC(3, true); else RTC_LAUNCH_ENC(3, false); }
#undef RTC_LAUNCH_ENC
    return cudaGetLastError();
}

}  // namespace rtc
