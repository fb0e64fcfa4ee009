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
// Colour keys of this lane's kEncC cells (key[1..C]) and of the cell before them (key[0]) from aligned
// 32-bit loads around an arbitrarily aligned plane.  Words outside the plane are never touched.
template <int BPP>
__device__ __forceinline__ void load_keys(const uint8_t* __restrict__ color, size_t plane_bytes, uint32_t cell,
                                          uint32_t (&key)[kEncC + 1])
{
    constexpr int NIN = BPP == 3 ? 6 : 3;                       // words covering (C+1)*BPP bytes at any phase
    const uintptr_t base = reinterpret_cast<uintptr_t>(color);
    const uintptr_t first_w = base & ~(uintptr_t)3, last_w = (base + plane_bytes - 1) & ~(uintptr_t)3;
    const uintptr_t a = base + (size_t)cell * BPP - BPP;        // predecessor key (garbage for cell 0: unused)
    const uintptr_t wa = a & ~(uintptr_t)3;
    const uint32_t sh = 8u * (uint32_t)(a & 3u);
    uint32_t w[NIN];
    if (wa >= first_w && wa + 4u * (NIN - 1) <= last_w) {
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

// Bit i of full_mask: cell i emits its whole escape sequence; bit i of nl_mask: cell i ends a row.
// Both restricted to the n_valid leading cells.  Returns the lane's emitted byte count.
template <int BPP>
__device__ __forceinline__ uint32_t lane_layout(const uint32_t (&key)[kEncC + 1], uint32_t cell, int n_valid, uint32_t W,
                                                uint32_t& full_mask, uint32_t& nl_mask)
{
    constexpr uint32_t CS = BPP == 3 ? 20u : 12u;               // SIZE_RGB / SIZE_8BIT (RayTracing.h:120-123)
    uint32_t fm = 0, nm = 0;
    uint32_t col = cell % W;
#pragma unroll
    for (int i = 0; i < kEncC; ++i) {
        const bool differs = key[i + 1] != key[i] || (cell + i == 0u);   // first cell of the frame always emits
        fm |= differs ? (1u << i) : 0u;
        const bool nl = col == W - 1u;
        nm |= nl ? (1u << i) : 0u;
        col = nl ? 0u : col + 1u;
    }
    const uint32_t vm = (1u << n_valid) - 1u;
    full_mask = fm & vm;
    nl_mask = nm & vm;
    return (uint32_t)n_valid + (CS - 1u) * __popc(full_mask) + __popc(nl_mask);
}

// ---- pass 1: per-tile byte counts + per-warp offsets inside the tile ---------------------------
template <int BPP>
__global__ void __launch_bounds__(kEncThreads)
count_kernel(const uint8_t* __restrict__ color, uint32_t W, uint32_t n_cells, uint32_t* __restrict__ tile_len,
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
        load_keys<BPP>(color, (size_t)n_cells * BPP, cell, key);
        len = lane_layout<BPP>(key, cell, n_valid, W, fm, nm);
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
constexpr int kScanPerThread = 16;
__global__ void __launch_bounds__(1024)
scan_kernel(const uint32_t* __restrict__ tile_len, uint32_t n_tiles, unsigned long long* __restrict__ tile_off,
            unsigned long long* __restrict__ total)
{
    __shared__ uint32_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long carry = 0ull;
    // chunks of 1024 x 16 tiles; a chunk's sum fits 32 bits (16384 tiles x <= 26880 bytes)
    for (uint32_t base = 0; base < n_tiles; base += 1024u * kScanPerThread) {
        uint32_t v[kScanPerThread];
        const uint32_t a = base + kScanPerThread * tid;
#pragma unroll
        for (int k = 0; k < kScanPerThread; ++k) v[k] = (a + k < n_tiles) ? tile_len[a + k] : 0u;
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < kScanPerThread; ++k) sum += v[k];
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31u) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wsum = s_warp[lane];                              // every warp scans the 32 warp sums itself
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wsum, o);
            if (lane >= (uint32_t)o) wsum += t;
        }
        const uint32_t before = __shfl_sync(0xffffffffu, wsum, (warp + 31u) & 31u);   // inclusive sum of warps < warp
        const uint32_t chunk_total = __shfl_sync(0xffffffffu, wsum, 31);
        unsigned long long run = carry + (warp ? before : 0u) + (inc - sum);
#pragma unroll
        for (int k = 0; k < kScanPerThread; ++k) {
            if (a + k < n_tiles) tile_off[a + k] = run;
            run += v[k];
        }
        carry += chunk_total;
        __syncthreads();                                           // s_warp is rewritten by the next chunk
    }
    if (tid == 0) *total = carry;
}

// ---- pass 3: emit -----------------------------------------------------------------------------
template <int BPP, bool GLYPH>
__global__ void __launch_bounds__(kEncThreads)
emit_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ glyph, uint32_t W, uint32_t n_cells,
            char* __restrict__ out, unsigned long long cap, const unsigned long long* __restrict__ tile_off,
            const uint32_t* __restrict__ warp_excl)
{
    constexpr int CS = BPP == 3 ? 20 : 12;
    constexpr int NWC = CS / 4;                                 // words per full cell
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

    // ---- phase A: keys, lengths, formatted full cells, trailing-bytes shift register ----------
    uint32_t key[kEncC + 1], fm = 0, nm = 0, len = 0;
    if (n_valid > 0) {
        load_keys<BPP>(color, (size_t)n_cells * BPP, cell, key);
        len = lane_layout<BPP>(key, cell, n_valid, W, fm, nm);
    }
    uint32_t cw[kEncC][NWC];
    uint32_t gch[kEncC];
    uint32_t sr = 0;                                            // the last 4 bytes this lane emits
#pragma unroll
    for (int i = 0; i < kEncC; ++i) {
        gch[i] = 32u;
        if (i < n_valid) {
            if (GLYPH) gch[i] = (uint32_t)__ldg(glyph + cell + i);
            if ((fm >> i) & 1u) {
                const uint32_t g = gch[i];
                const uint32_t sel = (GLYPH && g != 32u) ? (uint32_t)'3' : (uint32_t)'4';   // fg for an ASCII-mode hit
                const uint32_t mch = 'm' | (g << 8);
                const uint32_t k = key[i + 1];
                cw[i][0] = 0x1bu | ('[' << 8) | (sel << 16) | ('8' << 24);
                if (BPP == 3) {
                    // ESC [ S 8 | ; 2 ; R2 | R1 R0 ; G2 | G1 G0 ; B2 | B1 B0 m CH   (RayTracing.cu:585-594)
                    const uint32_t lr = s_lut[k & 255u], lg = s_lut[(k >> 8) & 255u], lb = s_lut[k >> 16];
                    cw[i][1] = __byte_perm(';' | ('2' << 8) | (';' << 16), lr, 0x4210);
                    cw[i][2 % NWC] = __byte_perm(lr, lg, 0x4321);
                    cw[i][3 % NWC] = __byte_perm(lg, lb, 0x4321);
                    cw[i][NWC - 1] = __byte_perm(lb, mch, 0x5421);
                } else {
                    // ESC [ S 8 | ; 5 ; I2 | I1 I0 m CH                              (RayTracing.cu:231-237)
                    const uint32_t li8 = s_lut[k];
                    cw[i][1] = __byte_perm(';' | ('5' << 8) | (';' << 16), li8, 0x4210);
                    cw[i][NWC - 1] = __byte_perm(li8, mch, 0x5421);
                }
                sr = cw[i][NWC - 1];
            } else {
                sr = __byte_perm(sr, gch[i], 0x4321);           // same colour as the previous cell: character only
            }
            if ((nm >> i) & 1u) sr = __byte_perm(sr, (uint32_t)'\n', 0x4321);
        }
    }

    // ---- warp prefix sum of the lane lengths ---------------------------------------------------
    uint32_t inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t warp_len = __shfl_sync(0xffffffffu, inc, 31);
    const uint32_t left_sr = __shfl_up_sync(0xffffffffu, sr, 1);   // lane 0: its own sr, bytes never copied out

    // ---- phase B: stream the cells into the staging image (phase-aligned with the output) ------
    const uint32_t out_phase = (uint32_t)(reinterpret_cast<uintptr_t>(out + goff) & 15u);
    if (n_valid > 0) {
        const uint32_t pos = out_phase + (inc - len);
        uint32_t k = pos & 3u;                                  // pending bytes (the top k bytes of acc)
        uint32_t* wp = reinterpret_cast<uint32_t*>(stage) + (pos >> 2);
        uint32_t acc = left_sr;
#pragma unroll
        for (int i = 0; i < kEncC; ++i) {
            if (i < n_valid) {
                if ((fm >> i) & 1u) {
                    const uint32_t sel = 0x7654u - 0x1111u * k;   // bytes [4-k .. 7-k] of {previous word, this word}
                    wp[0] = __byte_perm(acc, cw[i][0], sel);
#pragma unroll
                    for (int j = 1; j < NWC; ++j) wp[j] = __byte_perm(cw[i][j - 1], cw[i][j], sel);
                    acc = cw[i][NWC - 1];
                    wp += NWC;
                } else {
                    acc = __byte_perm(acc, gch[i], 0x4321);
                    if (++k == 4u) { *wp++ = acc; k = 0u; }
                }
                if ((nm >> i) & 1u) {
                    acc = __byte_perm(acc, (uint32_t)'\n', 0x4321);
                    if (++k == 4u) { *wp++ = acc; k = 0u; }
                }
            }
        }
        // the slice's last lane flushes its pending bytes itself (everyone else's go to the right neighbour)
        if (k != 0u && t5 + (uint32_t)n_valid == n_here) *wp = acc >> (8u * (4u - k));
    }
    __syncwarp();

    // ---- copy the slice out: coalesced 128-bit stores -------------------------------------
    if (goff >= cap) return;
    const uint32_t n_out = (uint32_t)min((unsigned long long)warp_len, cap - goff);
    char* dst = out + goff;
    const unsigned char* src = stage + out_phase;
    const uint32_t head = out_phase ? min(16u - out_phase, n_out) : 0u;
    if ((uint32_t)lane < head) dst[lane] = (char)src[lane];
    const uint32_t nvec = (n_out - head) >> 4;
    uint4* vdst = reinterpret_cast<uint4*>(dst + head);
    const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
    for (uint32_t i = lane; i < nvec; i += 32u) vdst[i] = vsrc[i];
    const uint32_t done = head + (nvec << 4);
    if ((uint32_t)lane < n_out - done) dst[done + lane] = (char)src[done + lane];
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
    if (bit8) count_kernel<1><<<n_tiles, kEncThreads, 0, st>>>(color, W, n_cells, tile_len, warp_excl);
    else count_kernel<3><<<n_tiles, kEncThreads, 0, st>>>(color, W, n_cells, tile_len, warp_excl);
    scan_kernel<<<1, 1024, 0, st>>>(tile_len, n_tiles, tile_off, total);
#define RTC_LAUNCH_ENC(BPP, GL)                                                                         \
    emit_kernel<BPP, GL><<<n_tiles, kEncThreads, enc_smem(), st>>>(                                     \
        color, glyph, W, n_cells, out, (unsigned long long)cap, tile_off, warp_excl)
    if (bit8) { if (has_glyph) RTC_LAUNCH_ENC(1, true); else RTC_LAUNCH_ENC(1, false); }
    else      { if (has_glyph) RTC_LAUNCH_ENC(3, true); else RTC_LAUNCH_ENC(3, false); }
#undef RTC_LAUNCH_ENC
    return cudaGetLastError();
}

}  // namespace rtc
