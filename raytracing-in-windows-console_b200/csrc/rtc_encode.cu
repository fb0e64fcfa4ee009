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
// Three launches, no inter-CTA waiting (a single-pass decoupled look-back was measured first: its
// per-tile dependency latency, not bandwidth, bounded it at ~2.2 TB/s -- profiles/r01_encode.md):
//   1. count : per tile of 1024 cells, the emitted byte count (reads the colour plane once;
//              for frames up to 8K the plane stays in the 126 MB L2 for pass 3);
//   2. scan  : one CTA turns the tile counts into exclusive 64-bit offsets + the stream length;
//   3. emit  : one CTA per tile: stage the colour bytes in shared memory with 128-bit loads;
//              every warp derives the per-cell lengths of its 4 x 32 cells from two ballots per
//              round (no shuffles: the in-round exclusive offset is popc arithmetic); cells are
//              formatted with a 256-entry digit LUT and byte permutes straight into a
//              shared-memory image of the tile's slice of the stream, phase-aligned with the
//              global offset, and the slice is copied out with coalesced 128-bit stores.
// Algorithmic traffic: BPP (+1) bytes read and the emitted bytes written per cell.
#include "rtc_device.cuh"
#include "rtc_kernels.h"

namespace rtc {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kEncRounds = 4;                                  // rounds of 32 cells per warp
constexpr int kEncTile = kEncWarps * kEncRounds * 32;          // 1024 cells per CTA

// NUL-padded 3 decimal digits of v (RayTracing.cu:526-543) packed as D2 | D1<<8 | D0<<16 | ';'<<24.
// The ';' rides along so that one PRMT assembles "D1 D0 ; D2'" words of the cell.
__host__ __device__ constexpr uint32_t digits_entry(uint32_t v)
{
    return (v >= 100u ? 48u + v / 100u : 0u) | ((v >= 10u ? 48u + (v / 10u) % 10u : 0u) << 8) | ((48u + v % 10u) << 16) | (59u << 24);
}

// little-endian 32-bit load at an arbitrary byte offset of a 4-byte aligned shared buffer
__device__ __forceinline__ uint32_t lds_unaligned(const unsigned char* base, uint32_t byte_off)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(base) + (byte_off >> 2);
    return __funnelshift_r(w[0], w[1], 8u * (byte_off & 3u));
}

// Store NW little-endian words of cell bytes at byte position `pos` of the staging image
// (arbitrary alignment).  One byte-funnel (PRMT with a run-time selector) per output word;
// the NW-1 inner words are whole STS.32, the ragged head and tail are predicated byte stores.
template <int NW>
__device__ __forceinline__ void put_words(unsigned char* stage, uint32_t pos, const uint32_t (&w)[NW])
{
    const uint32_t k = pos & 3u;
    const uint32_t sel = 0x7654u - 0x1111u * k;             // bytes [4-k .. 7-k] of {prev, cur}
    uint32_t* wp = reinterpret_cast<uint32_t*>(stage + (pos - k));
    uint32_t o[NW + 1];
    o[0] = __byte_perm(0u, w[0], sel);
#pragma unroll
    for (int j = 1; j < NW; ++j) o[j] = __byte_perm(w[j - 1], w[j], sel);
    o[NW] = __byte_perm(w[NW - 1], 0u, sel);
#pragma unroll
    for (int j = 1; j < NW; ++j) wp[j] = o[j];
    unsigned char* hb = reinterpret_cast<unsigned char*>(wp);
    if (k == 0u) wp[0] = o[0];
    if (k == 1u) hb[1] = (unsigned char)(o[0] >> 8);
    if (k == 1u || k == 2u) hb[2] = (unsigned char)(o[0] >> 16);
    if (k != 0u) hb[3] = (unsigned char)(o[0] >> 24);
    unsigned char* tb = reinterpret_cast<unsigned char*>(wp + NW);
    if (k != 0u) tb[0] = (unsigned char)o[NW];
    if (k >= 2u) tb[1] = (unsigned char)(o[NW] >> 8);
    if (k == 3u) tb[2] = (unsigned char)(o[NW] >> 16);
}

// ---- pass 1: per-tile emitted byte counts ---------------------------------------------------
// A thread takes 4 consecutive cells = 12 (or 4) colour bytes as aligned 32-bit words; the row-end
// newlines of a tile are counted arithmetically (no per-cell modulo).
template <int BPP>
__global__ void __launch_bounds__(kEncThreads)
count_kernel(const uint8_t* __restrict__ color, uint32_t W, uint32_t n_cells, uint32_t* __restrict__ tile_len)
{
    constexpr uint32_t CS = BPP == 3 ? 20u : 12u;
    __shared__ uint32_t s_sum[kEncWarps];
    const uint32_t tile = blockIdx.x, cell0 = tile * (uint32_t)kEncTile;
    const uint32_t n_here = min((uint32_t)kEncTile, n_cells - cell0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t n_full = 0;
    const uint32_t c = cell0 + 4u * tid;                         // first of this thread's 4 cells
    const uint8_t* p = color + (size_t)c * BPP;
    if (c + 4u <= n_cells && (reinterpret_cast<uintptr_t>(p) & 3u) == 0 && c != 0u) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
        if (BPP == 3) {
            const uint32_t wp = __ldg(w - 1), w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            const uint32_t kp = wp >> 8, k0 = w0 & 0xffffffu, k1 = __funnelshift_r(w0, w1, 24) & 0xffffffu,
                           k2 = __funnelshift_r(w1, w2, 16) & 0xffffffu, k3 = w2 >> 8;
            n_full = (k0 != kp) + (k1 != k0) + (k2 != k1) + (k3 != k2);
        } else {
            const uint32_t wp = __ldg(w - 1), w0 = __ldg(w);
            const uint32_t sh = __funnelshift_r(wp, w0, 24);      // each byte's predecessor
            const uint32_t d = w0 ^ sh;
            n_full = ((d & 0xffu) != 0) + ((d & 0xff00u) != 0) + ((d & 0xff0000u) != 0) + ((d & 0xff000000u) != 0);
        }
    } else {
        for (uint32_t i = c; i < min(c + 4u, n_cells); ++i) {    // ragged tail / unaligned plane / very first cell
            const uint8_t* q = color + (size_t)i * BPP;
            bool differs = (i == 0u);
            if (!differs) {
#pragma unroll
                for (int b = 0; b < BPP; ++b) differs |= q[b] != q[b - BPP];
            }
            n_full += differs ? 1u : 0u;
        }
    }
    uint32_t sum = n_full;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_sum[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w2 = 0; w2 < kEncWarps; ++w2) t += s_sum[w2];
        const uint32_t newlines = (cell0 + n_here) / W - cell0 / W;     // row ends inside [cell0, cell0 + n_here)
        tile_len[tile] = (CS - 1u) * t + n_here + newlines;
    }
}

// ---- pass 2: exclusive scan of the tile counts (one CTA) ------------------------------------
__global__ void __launch_bounds__(1024)
scan_kernel(const uint32_t* __restrict__ tile_len, uint32_t n_tiles, unsigned long long* __restrict__ tile_off,
            unsigned long long* __restrict__ total)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0ull;
    __syncthreads();
    // chunks of 1024 x 8 tiles; within a chunk a thread owns 8 consecutive tiles (two 16-byte loads)
    for (uint32_t base = 0; base < n_tiles; base += 8192u) {
        uint32_t v[8];
        const uint32_t a = base + 8u * tid;
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (a + k < n_tiles) ? tile_len[a + k] : 0u;
        unsigned long long sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += v[k];
        unsigned long long inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31u) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += t;
            }
            s_warp[lane] = w;                                       // inclusive over warps
        }
        __syncthreads();
        unsigned long long run = s_carry + (warp ? s_warp[warp - 1] : 0ull) + (inc - sum);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (a + k < n_tiles) tile_off[a + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (tid == 1023u) s_carry = run;
        __syncthreads();
    }
    if (tid == 0) *total = s_carry;
}

// ---- pass 3: emit -----------------------------------------------------------------------------
template <int BPP, bool GLYPH>
__global__ void __launch_bounds__(kEncThreads)
encode_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ glyph, uint32_t W, uint32_t n_cells,
              char* __restrict__ out, unsigned long long cap, const unsigned long long* __restrict__ tile_off)
{
    constexpr int CS = BPP == 3 ? 20 : 12;          // SIZE_RGB / SIZE_8BIT (RayTracing.h:120-123)
    constexpr int IN_BYTES = kEncTile * BPP + BPP + 32;
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* s_in = smem;                                 // colour bytes, phase-aligned with global
    unsigned char* s_gl = s_in + ((IN_BYTES + 15) & ~15);       // glyph bytes
    unsigned char* s_stage = s_gl + (GLYPH ? kEncTile + 32 : 0);
    __shared__ uint32_t s_lut[256];
    __shared__ uint32_t s_warp_tot[kEncWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    s_lut[tid] = digits_entry((uint32_t)tid);
    const uint32_t tile = blockIdx.x;
    const uint32_t cell0 = tile * (uint32_t)kEncTile;
    const uint32_t n_here = min((uint32_t)kEncTile, n_cells - cell0);
    const unsigned long long gbase = tile_off[tile];

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
    // key of local cell i lives at byte key_off + i*BPP of s_in; its predecessor BPP bytes before
    const uint32_t key_off = in_phase + (cell0 == 0 ? 0u : (uint32_t)BPP);
    constexpr uint32_t KEYMASK = BPP == 3 ? 0xffffffu : 0xffu;

    // ---- phase A: lengths -> warp totals ---------------------------------------------
    uint32_t full_mask[kEncRounds], nl_mask[kEncRounds], valid_mask[kEncRounds];
    uint32_t key[kEncRounds];
    uint32_t warp_total = 0;
    const uint32_t wcell0 = warp * (kEncRounds * 32);
    uint32_t col = (cell0 + wcell0 + lane) % W;
#pragma unroll
    for (int r = 0; r < kEncRounds; ++r) {
        const uint32_t li = wcell0 + r * 32 + lane;             // local cell index
        const bool valid = li < n_here;
        uint32_t k = 0, kp = 0xffffffffu;
        if (valid) {
            const uint32_t bo = key_off + li * BPP;
            k = lds_unaligned(s_in, bo) & KEYMASK;
            if (cell0 + li != 0u) kp = lds_unaligned(s_in, bo - BPP) & KEYMASK;
        }
        key[r] = k;
        const bool full = valid && (k != kp);
        const bool nl = valid && (col == W - 1u);
        full_mask[r] = __ballot_sync(0xffffffffu, full);
        nl_mask[r] = __ballot_sync(0xffffffffu, nl);
        valid_mask[r] = __ballot_sync(0xffffffffu, valid);
        warp_total += (uint32_t)(CS - 1) * __popc(full_mask[r]) + __popc(valid_mask[r]) + __popc(nl_mask[r]);
        col += 32u;
        if (col >= W) col %= W;
    }
    if (lane == 0) s_warp_tot[warp] = warp_total;
    __syncthreads();

    // ---- phase B1: format full cells into registers ---------------------------------------------
    uint32_t cw[kEncRounds][BPP == 3 ? 5 : 3];
    uint32_t gch[kEncRounds];
#pragma unroll
    for (int r = 0; r < kEncRounds; ++r) {
        const uint32_t li = wcell0 + r * 32 + lane;
        const bool valid = (valid_mask[r] >> lane) & 1u;
        const uint32_t g = (GLYPH && valid) ? (uint32_t)s_gl[gl_phase + li] : 32u;
        gch[r] = g;
        const uint32_t sel = (GLYPH && g != 32u) ? (uint32_t)'3' : (uint32_t)'4';   // fg for an ASCII-mode hit
        const uint32_t w0 = 0x1bu | ('[' << 8) | (sel << 16) | ('8' << 24);
        const uint32_t mch = 'm' | (g << 8);
        if (BPP == 3) {
            // ESC [ S 8 | ; 2 ; R2 | R1 R0 ; G2 | G1 G0 ; B2 | B1 B0 m CH   (RayTracing.cu:585-594)
            const uint32_t lr = s_lut[key[r] & 255u], lg = s_lut[(key[r] >> 8) & 255u], lb = s_lut[(key[r] >> 16) & 255u];
            cw[r][0] = w0;
            cw[r][1] = __byte_perm(';' | ('2' << 8) | (';' << 16), lr, 0x4210);
            cw[r][2] = __byte_perm(lr, lg, 0x4321);
            cw[r][3] = __byte_perm(lg, lb, 0x4321);
            cw[r][BPP == 3 ? 4 : 2] = __byte_perm(lb, mch, 0x5421);
        } else {
            // ESC [ S 8 | ; 5 ; I2 | I1 I0 m CH                              (RayTracing.cu:231-237)
            const uint32_t li8 = s_lut[key[r] & 255u];
            cw[r][0] = w0;
            cw[r][1] = __byte_perm(';' | ('5' << 8) | (';' << 16), li8, 0x4210);
            cw[r][2] = __byte_perm(li8, mch, 0x5421);
        }
    }
    const uint32_t out_phase = (uint32_t)(reinterpret_cast<uintptr_t>(out + gbase) & 15u);

    // ---- phase B2: place the cells into the staging image (phase-aligned with the output) ----
    uint32_t wofs = 0, tile_len = 0;
#pragma unroll
    for (int w = 0; w < kEncWarps; ++w) {
        const uint32_t t = s_warp_tot[w];
        wofs += (w < warp) ? t : 0u;
        tile_len += t;
    }
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t pos_round = out_phase + wofs;
#pragma unroll
    for (int r = 0; r < kEncRounds; ++r) {
        const uint32_t fm = full_mask[r], nm = nl_mask[r], vm = valid_mask[r];
        const uint32_t pos = pos_round + (uint32_t)(CS - 1) * __popc(fm & lt) + __popc(vm & lt) + __popc(nm & lt);
        const bool valid = (vm >> lane) & 1u, full = (fm >> lane) & 1u, nl = (nm >> lane) & 1u;
        if (valid) {
            if (full) put_words(s_stage, pos, cw[r]);
            else s_stage[pos] = (unsigned char)gch[r];          // same colour as the previous cell: character only
            if (nl) s_stage[pos + (full ? CS : 1)] = '\n';
        }
        pos_round += (uint32_t)(CS - 1) * __popc(fm) + __popc(vm) + __popc(nm);
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

// scratch: per tile one u32 count + one u64 offset
size_t encode_state_bytes(uint64_t n_cells)
{
    const uint64_t n_tiles = (n_cells + kEncTile - 1) / kEncTile + 1;
    return (size_t)(n_tiles * 12 + 64);
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
    const bool has_glyph = (mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII) && glyph != nullptr;
    const bool bit8 = (mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL);
    if (bit8) count_kernel<1><<<n_tiles, kEncThreads, 0, st>>>(color, W, n_cells, tile_len);
    else count_kernel<3><<<n_tiles, kEncThreads, 0, st>>>(color, W, n_cells, tile_len);
    scan_kernel<<<1, 1024, 0, st>>>(tile_len, n_tiles, tile_off, total);
#define RTC_LAUNCH_ENC(BPP, GL)                                                                         \
    encode_kernel<BPP, GL><<<n_tiles, kEncThreads, enc_smem<BPP, GL>(), st>>>(                          \
        color, glyph, W, n_cells, out, (unsigned long long)cap, tile_off)
    if (bit8) { if (has_glyph) RTC_LAUNCH_ENC(1, true); else RTC_LAUNCH_ENC(1, false); }
    else      { if (has_glyph) RTC_LAUNCH_ENC(3, true); else RTC_LAUNCH_ENC(3, false); }
#undef RTC_LAUNCH_ENC
    return cudaGetLastError();
}

}  // namespace rtc
