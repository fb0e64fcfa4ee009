// rtc_encode.cu -- kernel 3: warp-cooperative ANSI encoder (single pass, HBM-bound).
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
// One CTA = one tile of 2048 cells.  Per tile: stage the colour bytes in shared memory with
// 128-bit loads; every warp derives the per-cell lengths of its 8 x 32 cells from two ballots
// per round (no shuffles: the in-round exclusive offset is popc arithmetic); tile totals are
// chained across CTAs with a decoupled look-back (one 64-bit descriptor per tile, epoch-tagged
// so it never needs clearing); cells are then formatted straight into a shared-memory image of
// the tile's slice of the stream, phase-aligned with the global offset, and the slice is
// copied out with coalesced 128-bit stores.  Algorithmic traffic: BPP (+1) bytes read and the
// emitted bytes written per cell; nothing else touches HBM.
#include "rtc_device.cuh"
#include "rtc_kernels.h"

namespace rtc {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncRounds = 8;                                  // rounds of 32 cells per warp
constexpr int kEncTile = kEncWarps * kEncRounds * 32;          // 2048 cells per CTA

// descriptor: [63:62] status (1 = aggregate, 2 = inclusive prefix) [61:40] epoch [39:0] value
constexpr unsigned long long kValMask = (1ull << 40) - 1ull;
__device__ __forceinline__ unsigned long long make_desc(unsigned status, unsigned epoch, unsigned long long v)
{
    return ((unsigned long long)status << 62) | ((unsigned long long)(epoch & 0x3fffffu) << 40) | (v & kValMask);
}

__device__ __forceinline__ void digits3(uint32_t v, uint32_t& d2, uint32_t& d1, uint32_t& d0)   // RayTracing.cu:526-543
{
    const uint32_t h = (v * 41u) >> 12;             // v / 100 for v < 256
    const uint32_t rem = v - h * 100u;
    const uint32_t t = (rem * 205u) >> 11;          // rem / 10 for rem < 100
    d2 = v >= 100u ? 48u + h : 0u;                  // NUL padding, not '0' or ' '
    d1 = v >= 10u ? 48u + t : 0u;
    d0 = 48u + (rem - t * 10u);
}

// Store `n_words` little-endian words of cell bytes at byte position `pos` of the staging
// image (arbitrary alignment): whole words with STS.32, the ragged head/tail with byte stores.
template <int NW>
__device__ __forceinline__ void put_words(unsigned char* stage, uint32_t pos, const uint32_t (&w)[NW])
{
    const uint32_t k = pos & 3u;
    uint32_t* wp = reinterpret_cast<uint32_t*>(stage + (pos - k));
    if (k == 0u) {
#pragma unroll
        for (int j = 0; j < NW; ++j) wp[j] = w[j];
    } else {
        const uint32_t sh = 8u * k;
#pragma unroll
        for (int j = 1; j < NW; ++j) wp[j] = __funnelshift_l(w[j - 1], w[j], sh);
        unsigned char* hp = stage + pos;           // head: bytes 0 .. 3-k of w[0]
        for (uint32_t b = 0; b < 4u - k; ++b) hp[b] = (unsigned char)(w[0] >> (8u * b));
        unsigned char* tp = stage + pos + 4u * NW - k;   // tail: top k bytes of w[NW-1]
        for (uint32_t b = 0; b < k; ++b) tp[b] = (unsigned char)(w[NW - 1] >> (8u * (4u - k + b)));
    }
}

template <int BPP, bool GLYPH>
__global__ void __launch_bounds__(kEncThreads)
encode_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ glyph, uint32_t W, uint32_t n_cells,
              char* __restrict__ out, unsigned long long cap, unsigned long long* __restrict__ total,
              unsigned long long* __restrict__ desc, unsigned int* __restrict__ ticket, unsigned int ticket_base,
              unsigned int epoch, unsigned int n_tiles)
{
    constexpr int CS = BPP == 3 ? 20 : 12;          // SIZE_RGB / SIZE_8BIT (RayTracing.h:120-123)
    constexpr int IN_BYTES = kEncTile * BPP + BPP + 32;
    constexpr int STAGE_BYTES = kEncTile * (CS + 1) + 48;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* s_in = smem;                                 // colour bytes, phase-aligned with global
    unsigned char* s_gl = s_in + ((IN_BYTES + 15) & ~15);       // glyph bytes
    unsigned char* s_stage = s_gl + (GLYPH ? kEncTile + 32 : 0);
    __shared__ unsigned long long s_warp_tot[kEncWarps];
    __shared__ unsigned long long s_tile_base;
    __shared__ unsigned int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;   // dynamic tile id: look-back never waits on an unscheduled CTA
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t cell0 = tile * (uint32_t)kEncTile;
    const uint32_t n_here = min((uint32_t)kEncTile, n_cells - cell0);

    // ---- stage input -------------------------------------------------------------------
    // colour bytes [b0, b1) with b0 one cell before the tile (the predecessor key)
    const size_t b0 = cell0 == 0 ? 0 : (size_t)cell0 * BPP - BPP;
    const size_t b1 = ((size_t)cell0 + n_here) * BPP;
    const uint32_t in_phase = (uint32_t)(reinterpret_cast<uintptr_t>(color + b0) & 15u);
    {
        const uint8_t* src = color + b0;
        const uint32_t nbytes = (uint32_t)(b1 - b0);
        const uint32_t head = in_phase ? min(16u - in_phase, nbytes) : 0u;
        const uint32_t nvec = (nbytes - head) >> 4;
        if ((uint32_t)tid < head) s_in[in_phase + tid] = src[tid];
        const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
        uint4* vdst = reinterpret_cast<uint4*>(s_in + in_phase + head);
        for (uint32_t i = tid; i < nvec; i += kEncThreads) vdst[i] = __ldg(vsrc + i);
        const uint32_t done = head + (nvec << 4);
        if ((uint32_t)tid < nbytes - done) s_in[in_phase + done + tid] = src[done + tid];
    }
    uint32_t gl_phase = 0;
    if (GLYPH) {
        const uint8_t* src = glyph + cell0;
        gl_phase = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint32_t head = gl_phase ? min(16u - gl_phase, n_here) : 0u;
        const uint32_t nvec = (n_here - head) >> 4;
        if ((uint32_t)tid < head) s_gl[gl_phase + tid] = src[tid];
        const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
        uint4* vdst = reinterpret_cast<uint4*>(s_gl + gl_phase + head);
        for (uint32_t i = tid; i < nvec; i += kEncThreads) vdst[i] = __ldg(vsrc + i);
        const uint32_t done = head + (nvec << 4);
        if ((uint32_t)tid < n_here - done) s_gl[gl_phase + done + tid] = src[done + tid];
    }
    __syncthreads();
    // key of local cell i lives at s_key + i*BPP; its predecessor at s_key + (i-1)*BPP
    const unsigned char* s_key = s_in + in_phase + (cell0 == 0 ? 0 : BPP);

    // ---- phase A: lengths -> warp totals ---------------------------------------------
    uint32_t full_mask[kEncRounds], nl_mask[kEncRounds], valid_mask[kEncRounds];
    uint32_t key[kEncRounds];
    uint32_t warp_total = 0;
    const uint32_t wcell0 = warp * (kEncRounds * 32);
    uint32_t col = 0;
    {
        const uint32_t g = cell0 + wcell0 + lane;
        col = g % W;
    }
#pragma unroll
    for (int r = 0; r < kEncRounds; ++r) {
        const uint32_t li = wcell0 + r * 32 + lane;             // local cell index
        const bool valid = li < n_here;
        uint32_t k = 0, kp = 0xffffffffu;
        if (valid) {
            const unsigned char* p = s_key + li * BPP;
            if (BPP == 3) {
                k = p[0] | (p[1] << 8) | (p[2] << 16);
                if (cell0 + li != 0u) kp = p[-3] | (p[-2] << 8) | (p[-1] << 16);
            } else {
                k = p[0];
                if (cell0 + li != 0u) kp = p[-1];
            }
        }
        key[r] = k;
        const bool full = valid && (k != kp);
        const bool nl = valid && (col == W - 1u);
        full_mask[r] = __ballot_sync(0xffffffffu, full);
        nl_mask[r] = __ballot_sync(0xffffffffu, nl);
        valid_mask[r] = __ballot_sync(0xffffffffu, valid);
        warp_total += (uint32_t)CS * __popc(full_mask[r]) + __popc(valid_mask[r] & ~full_mask[r]) + __popc(nl_mask[r]);
        col += 32u;
        if (col >= W) col %= W;
    }
    if (lane == 0) s_warp_tot[warp] = warp_total;
    __syncthreads();

    // ---- tile prefix: decoupled look-back (warp 0) ------------------------------------
    if (warp == 0) {
        unsigned long long tile_total = 0;
#pragma unroll
        for (int w = 0; w < kEncWarps; ++w) tile_total += s_warp_tot[w];
        unsigned long long exclusive = 0;
        if (tile == 0) {
            if (lane == 0) {
                __threadfence();
                atomicExch(&desc[0], make_desc(2u, epoch, tile_total));
            }
        } else {
            if (lane == 0) atomicExch(&desc[tile], make_desc(1u, epoch, tile_total));
            int look = (int)tile - 1;
            for (;;) {
                const int j = look - lane;                      // 32 predecessors per probe
                unsigned long long d = 0;
                bool ok = true;
                if (j >= 0) {
                    d = *reinterpret_cast<volatile unsigned long long*>(&desc[j]);
                    ok = ((unsigned)(d >> 40) & 0x3fffffu) == (epoch & 0x3fffffu) && (d >> 62) != 0ull;
                }
                const uint32_t ready = __ballot_sync(0xffffffffu, ok);
                if (ready != 0xffffffffu) continue;             // some descriptor not published yet: spin
                const bool is_prefix = j >= 0 && (d >> 62) == 2ull;
                const uint32_t pm = __ballot_sync(0xffffffffu, is_prefix);
                // sum aggregates of lanes before (and including) the first inclusive prefix
                const int first = pm ? __ffs(pm) - 1 : 32;
                unsigned long long v = (j >= 0 && lane <= first) ? (d & kValMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                exclusive += v;
                if (pm || look - 32 < 0) break;
                look -= 32;
            }
            if (lane == 0) {
                __threadfence();
                atomicExch(&desc[tile], make_desc(2u, epoch, exclusive + tile_total));
            }
        }
        if (lane == 0) {
            s_tile_base = exclusive;
            if (tile == n_tiles - 1u) *total = exclusive + tile_total;
        }
    }
    __syncthreads();
    const unsigned long long gbase = s_tile_base;
    const uint32_t out_phase = (uint32_t)(reinterpret_cast<uintptr_t>(out + gbase) & 15u);

    // ---- phase B: format cells into the staging image ------------------------------------
    uint32_t wofs = 0;                                          // warp's first byte within the tile slice
#pragma unroll
    for (int w = 0; w < kEncWarps; ++w) wofs += (w < warp) ? (uint32_t)s_warp_tot[w] : 0u;
    uint32_t tile_len = 0;
#pragma unroll
    for (int w = 0; w < kEncWarps; ++w) tile_len += (uint32_t)s_warp_tot[w];

    const uint32_t lt = (1u << lane) - 1u;
    uint32_t pos_round = out_phase + wofs;
#pragma unroll
    for (int r = 0; r < kEncRounds; ++r) {
        const uint32_t fm = full_mask[r], nm = nl_mask[r], vm = valid_mask[r];
        const uint32_t pos = pos_round + (uint32_t)CS * __popc(fm & lt) + __popc(vm & ~fm & lt) + __popc(nm & lt);
        const bool valid = (vm >> lane) & 1u, full = (fm >> lane) & 1u, nl = (nm >> lane) & 1u;
        if (valid) {
            const uint32_t li = wcell0 + r * 32 + lane;
            const uint32_t g = GLYPH ? (uint32_t)s_gl[gl_phase + li] : 32u;
            uint32_t len = 1;
            if (full) {
                const uint32_t sel = (GLYPH && g != 32u) ? (uint32_t)'3' : (uint32_t)'4';   // fg for an ASCII-mode hit
                if (BPP == 3) {
                    uint32_t r2, r1, r0, g2, g1, g0, b2, b1, b0_;
                    digits3(key[r] & 255u, r2, r1, r0);
                    digits3((key[r] >> 8) & 255u, g2, g1, g0);
                    digits3((key[r] >> 16) & 255u, b2, b1, b0_);
                    // ESC [ S 8 | ; 2 ; R2 | R1 R0 ; G2 | G1 G0 ; B2 | B1 B0 m CH   (RayTracing.cu:585-594)
                    const uint32_t w[5] = {0x1bu | ('[' << 8) | (sel << 16) | ('8' << 24),
                                           ';' | ('2' << 8) | (';' << 16) | (r2 << 24),
                                           r1 | (r0 << 8) | (';' << 16) | (g2 << 24),
                                           g1 | (g0 << 8) | (';' << 16) | (b2 << 24),
                                           b1 | (b0_ << 8) | ('m' << 16) | (g << 24)};
                    put_words<5>(s_stage, pos, w);
                } else {
                    uint32_t i2, i1, i0;
                    digits3(key[r] & 255u, i2, i1, i0);
                    // ESC [ S 8 | ; 5 ; I2 | I1 I0 m CH                              (RayTracing.cu:231-237)
                    const uint32_t w[3] = {0x1bu | ('[' << 8) | (sel << 16) | ('8' << 24),
                                           ';' | ('5' << 8) | (';' << 16) | (i2 << 24),
                                           i1 | (i0 << 8) | ('m' << 16) | (g << 24)};
                    put_words<3>(s_stage, pos, w);
                }
                len = CS;
            } else {
                s_stage[pos] = (unsigned char)g;                // same colour as the previous cell: character only
            }
            if (nl) s_stage[pos + len] = '\n';
        }
        pos_round += (uint32_t)CS * __popc(fm) + __popc(vm & ~fm) + __popc(nm);
    }
    __syncthreads();

    // ---- copy the slice out: coalesced 128-bit stores -------------------------------------
    if (gbase >= cap) return;
    const uint32_t len = (uint32_t)min((unsigned long long)tile_len, cap - gbase);
    char* dst = out + gbase;
    const unsigned char* src = s_stage + out_phase;
    const uint32_t head = out_phase ? min(16u - out_phase, len) : 0u;
    if ((uint32_t)tid < head) dst[tid] = (char)src[tid];
    const uint32_t nvec = (len - head) >> 4;
    uint4* vdst = reinterpret_cast<uint4*>(dst + head);
    const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
    for (uint32_t i = tid; i < nvec; i += kEncThreads) vdst[i] = vsrc[i];
    const uint32_t done = head + (nvec << 4);
    if ((uint32_t)tid < len - done) dst[done + tid] = (char)src[done + tid];
}

// SDL mode (reference RayTrace_SDL writes nothing, RayTracing.cu:755-795): y newlines.
__global__ void newline_kernel(char* __restrict__ out, uint32_t y, unsigned long long cap, unsigned long long* total)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < y && i < cap) out[i] = '\n';
    if (i == 0) *total = y;
}

template <int BPP, bool GLYPH>
static size_t enc_smem()
{
    constexpr int CS = BPP == 3 ? 20 : 12;
    constexpr int IN_BYTES = kEncTile * BPP + BPP + 32;
    return (size_t)((IN_BYTES + 15) & ~15) + (GLYPH ? kEncTile + 32 : 0) + kEncTile * (CS + 1) + 48;
}

cudaError_t configure_encode()
{
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(encode_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem<3, false>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(encode_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem<3, true>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(encode_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem<1, false>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(encode_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc_smem<1, true>())) != cudaSuccess) return e;
    return cudaSuccess;
}

size_t encode_state_bytes(uint64_t n_cells) { return ((n_cells + kEncTile - 1) / kEncTile + 1) * sizeof(unsigned long long); }

cudaError_t launch_encode(cudaStream_t st, const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y,
                          int mode, char* out, size_t cap, unsigned long long* total, unsigned long long* desc,
                          unsigned int* ticket, unsigned int* ticket_base /* host, in/out */,
                          unsigned int* epoch /* host, in/out */)
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
    *epoch += 1u;
    const unsigned int base = *ticket_base;
    *ticket_base += n_tiles;
    const bool has_glyph = (mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII) && glyph != nullptr;
    const bool bit8 = (mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL);
#define RTC_LAUNCH_ENC(BPP, GL)                                                                         \
    encode_kernel<BPP, GL><<<n_tiles, kEncThreads, enc_smem<BPP, GL>(), st>>>(                          \
        color, glyph, W, n_cells, out, (unsigned long long)cap, total, desc, ticket, base, *epoch, n_tiles)
    if (bit8) { if (has_glyph) RTC_LAUNCH_ENC(1, true); else RTC_LAUNCH_ENC(1, false); }
    else      { if (has_glyph) RTC_LAUNCH_ENC(3, true); else RTC_LAUNCH_ENC(3, false); }
#undef RTC_LAUNCH_ENC
    return cudaGetLastError();
}

}  // namespace rtc
