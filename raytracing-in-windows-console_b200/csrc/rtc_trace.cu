// rtc_trace.cu -- kernel 0 (per-frame scene hoist) and kernel 1 (ray generation + nearest hit).
//
// Replaces the hot loop of the reference's RayTrace (RayTracing.cu:81-136) and the
// Sphere::Trace / Plane::Trace it calls per object (Sphere.cu:30-68, Plane.cu:38-73).
//
// Design (B200 / sm_100a):
//   * All primary rays share the origin (RayTracing.cu:195), so oc = origin - centre and
//     c = |oc|^2 - r^2 are per-sphere, per-frame constants: kernel 0 hoists them (bit-identical
//     to what the reference recomputes per ray).
//   * Kernel 1 is persistent: one 512-thread CTA per SM keeps the whole sphere list in shared
//     memory (32 B per sphere PAIR) and walks 16x16-pixel screen tiles, one tile per warp,
//     8 rays per thread.  The inner loop tests TWO spheres against one ray per packed
//     instruction (FMUL2 + 3x FFMA2 = 7 FLOP per test, the algorithmic minimum) and folds the
//     two discriminants into a running maximum with one FMNMX3; one LDS.128 pair (warp
//     broadcast) feeds 16 tests.  The packed test is only a CONSERVATIVE filter (each sphere's
//     c is deflated by a rounding-error bound); survivors are re-evaluated with the reference's
//     exact operation order, so hit decisions and distances are bit-identical to the reference.
//   * The running best (distance, object index) per ray lives in shared memory: it is touched
//     only on the (rare) exact path and would otherwise cost 16 registers in the hot loop.
#include "rtc_device.cuh"
#include "rtc_kernels.h"

namespace rtc {

// Relative deflation of c for the conservative filter.  The packed discriminant
// s'^2 - c (s' fused) differs from the reference's fl(fl(s^2) - fl(a*c)) (s un-fused, a = d.d)
// by at most ~22 * 2^-24 * |oc|^2 (DESIGN.md "filter bound"); 6e-6 leaves > 4x headroom.
#define RTC_FILTER_EPS 6.0e-6f

// ---- kernel 0: hoist --------------------------------------------------------------------
// One thread per sphere slot.  sph_pairs: per pair p two float4: (ocx0,ocx1,ocy0,ocy1),
// (ocz0,ocz1,nc0,nc1) with nc = -(c - eps*|oc|^2).  sph_c: exact c per sphere.  Slots past
// n_spheres are never-hit sentinels (nc = -1).
__global__ void __launch_bounds__(256)
hoist_kernel(const rtc_object* __restrict__ objs, const int32_t* __restrict__ sphere_obj, int n_spheres,
             int n_slots, float camx, float camy, float camz, float4* __restrict__ sph_pairs,
             float* __restrict__ sph_c, unsigned int* __restrict__ counters, int n_counters)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_counters) counters[j] = 0u;     // tile tickets / scan state heads for this frame
    if (j >= n_slots) return;
    float ocx = 0.f, ocy = 0.f, ocz = 0.f, c = 1.f, nc = -1.f;
    if (j < n_spheres) {
        const rtc_object& s = objs[sphere_obj[j]];
        ocx = sub(camx, s.center[0]);                                   // Sphere.cu:34
        ocy = sub(camy, s.center[1]);
        ocz = sub(camz, s.center[2]);
        const float oc2 = vdot(v3(ocx, ocy, ocz), v3(ocx, ocy, ocz));
        c = sub(oc2, mul(s.radius, s.radius));                          // Sphere.cu:37
        nc = fmaf(RTC_FILTER_EPS, oc2, -c);
        // NaN/inf geometry: force "always a candidate" so the exact path decides.
        if (!(nc == nc) || fabsf(nc) > 3.0e38f) nc = 3.0e38f;
    }
    float* base = reinterpret_cast<float*>(sph_pairs + 2 * (j >> 1));
    const int h = j & 1;
    base[0 + h] = ocx; base[2 + h] = ocy; base[4 + h] = ocz; base[6 + h] = nc;
    sph_c[j] = c;
}

// ---- kernel 1: trace --------------------------------------------------------------------
constexpr int kRays = 8;            // rays per thread
constexpr int kThreads = 512;       // 16 warps, one CTA per SM
constexpr int kTile = 16;           // warp tile = 16 x 16 pixels

// Shared-memory layout (dynamic): [pairs float4 x 2*n_pairs][c float x n_slots][state]
// state: best_t, best_idx, fourA, divTwoA -- each [kRays][kThreads].
struct Smem {
    float4* pairs;
    float* c;
    float* best_t;
    int* best_idx;
    float* fourA;
    float* div2A;
};
__device__ __forceinline__ Smem carve(unsigned char* raw, int n_slots)
{
    Smem s;
    s.pairs = reinterpret_cast<float4*>(raw);
    s.c = reinterpret_cast<float*>(raw + (size_t)n_slots * 16);
    float* st = s.c + n_slots;
    s.best_t = st;
    s.best_idx = reinterpret_cast<int*>(st + kRays * kThreads);
    s.fourA = st + 2 * kRays * kThreads;
    s.div2A = st + 3 * kRays * kThreads;
    return s;
}

// Exact re-evaluation of the two spheres of pair `p` for one ray (rare path).
// Accept rule == reference RayTracing.cu:123 (strict '<', lowest index wins ties), written
// order-independently as the lexicographic minimum of (distance, object index).
__device__ __noinline__ void exact_pair(const int32_t* __restrict__ sphere_obj, int n_spheres, int n_slots,
                                        int p, int slot, float dx, float dy, float dz, float qlo, float qhi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, n_slots);
    const float4 A = s.pairs[2 * p], B = s.pairs[2 * p + 1];
    const float fourA = s.fourA[slot], div2A = s.div2A[slot];
    float best = s.best_t[slot];
    int bidx = s.best_idx[slot];
    bool changed = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float q = h ? qhi : qlo;
        const int j = 2 * p + h;
        if (!(q >= 0.0f) || j >= n_spheres) continue;
        const float ocx = h ? A.y : A.x, ocy = h ? A.w : A.z, ocz = h ? B.y : B.x;
        float t;
        if (!sphere_trace_hoisted(ocx, ocy, ocz, s.c[j], v3(dx, dy, dz), fourA, div2A, t)) continue;
        const int oi = sphere_obj[j];
        if (t < best || (t == best && oi < bidx)) { best = t; bidx = oi; changed = true; }
    }
    if (changed) { s.best_t[slot] = best; s.best_idx[slot] = bidx; }
}

__global__ void __launch_bounds__(kThreads, 1)
trace_kernel(const FrameParams fp, const float4* __restrict__ g_pairs, const float* __restrict__ g_c,
             const int32_t* __restrict__ sphere_obj, int n_spheres, int n_slots,
             const rtc_object* __restrict__ objs, const int32_t* __restrict__ plane_obj, int n_planes,
             float* __restrict__ hit_t, int32_t* __restrict__ hit_idx, unsigned int* __restrict__ tile_counter,
             int carry_in /* 1: continue from hit_t/hit_idx (sphere list chunking) */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, n_slots);
    const int tid = threadIdx.x, lane = tid & 31;

    // Stage the hoisted sphere list once per CTA (persistent kernel).
    for (int i = tid; i < n_slots; i += kThreads) {   // n_slots floats4 == 2 * n_pairs
        s.pairs[i] = g_pairs[i];
        s.c[i] = g_c[i];
    }
    __syncthreads();

    const uint32_t W = fp.x - 1u;
    const uint32_t rows = fp.row1 - fp.row0;
    const uint32_t tiles_x = (W + kTile - 1) / kTile, tiles_y = (rows + kTile - 1) / kTile;
    const uint32_t n_tiles = tiles_x * tiles_y;
    const int n_pairs = n_slots >> 1;
    const V3 o = v3(fp.cam[0], fp.cam[1], fp.cam[2]);
    const uint32_t lx = lane & 15u, ly = lane >> 4;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const uint32_t col = tx * kTile + lx;
        const uint32_t colc = col < W ? col : W - 1u;     // clamp: out-of-frame lanes trace a duplicate ray

        float dx[kRays], dy[kRays], dz[kRays];
#pragma unroll
        for (int r = 0; r < kRays; ++r) {
            uint32_t row = fp.row0 + ty * kTile + ly + 2u * r;
            if (row >= fp.row1) row = fp.row1 - 1u;
            const V3 d = initial_direction(fp, row, colc);
            dx[r] = d.x; dy[r] = d.y; dz[r] = d.z;
            const float a = vdot(d, d);                                   // RayTracing.cu:91
            const int slot = r * kThreads + tid;
            s.fourA[slot] = mul(4.0f, a);                                 // :92
            s.div2A[slot] = dvd(1.0f, mul(2.0f, a));                      // :93
            float bt = 99999999.f;                                        // RayTracing.h:21
            int bi = -1;
            if (carry_in) {
                const size_t pix = (size_t)(row - fp.row0) * W + colc;
                bt = hit_t[pix]; bi = hit_idx[pix];
            }
            s.best_t[slot] = bt;
            s.best_idx[slot] = bi;
        }

        // ---- hot loop: 2 spheres x 8 rays per iteration -----------------------------------
#pragma unroll 2
        for (int p = 0; p < n_pairs; ++p) {
            const float4 A = s.pairs[2 * p], B = s.pairs[2 * p + 1];     // LDS.128 x2, warp broadcast
            const f32x2 OX = pack2(A.x, A.y), OY = pack2(A.z, A.w), OZ = pack2(B.x, B.y), NC = pack2(B.z, B.w);
            f32x2 q[kRays];
            float m = -1.0f;
#pragma unroll
            for (int r = 0; r < kRays; ++r) {
                f32x2 t = mul2(pack2(dx[r], dx[r]), OX);                 // FMUL2  (scalar-broadcast operand)
                t = fma2(pack2(dy[r], dy[r]), OY, t);                    // FFMA2
                t = fma2(pack2(dz[r], dz[r]), OZ, t);                    // FFMA2  s' = d . oc   (two spheres)
                q[r] = fma2(t, t, NC);                                   // FFMA2  s'^2 - c'
            }
#pragma unroll
            for (int r = 0; r < kRays; ++r) {
                float lo, hi;
                unpack2(q[r], lo, hi);
                m = max3(m, lo, hi);                                     // FMNMX3
            }
            if (m >= 0.0f) {                                             // some ray of this lane may hit
#pragma unroll
                for (int r = 0; r < kRays; ++r) {
                    float lo, hi;
                    unpack2(q[r], lo, hi);
                    if (fmaxf(lo, hi) >= 0.0f)
                        exact_pair(sphere_obj, n_spheres, n_slots, p, r * kThreads + tid, dx[r], dy[r], dz[r], lo, hi);
                }
            }
        }

        // ---- planes (few): exact, every ray --------------------------------------------
        for (int k = 0; k < n_planes; ++k) {
            const int oi = plane_obj[k];
            const rtc_object pl = objs[oi];
#pragma unroll
            for (int r = 0; r < kRays; ++r) {
                float t;
                if (plane_trace(pl, o, v3(dx[r], dy[r], dz[r]), t)) {
                    const int slot = r * kThreads + tid;
                    const float best = s.best_t[slot];
                    if (t < best || (t == best && oi < s.best_idx[slot])) { s.best_t[slot] = t; s.best_idx[slot] = oi; }
                }
            }
        }

        // ---- hit records ------------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < kRays; ++r) {
            const uint32_t row = fp.row0 + ty * kTile + ly + 2u * r;
            if (row < fp.row1 && col < W) {
                const size_t pix = (size_t)(row - fp.row0) * W + col;
                const int slot = r * kThreads + tid;
                hit_t[pix] = s.best_t[slot];
                hit_idx[pix] = s.best_idx[slot];
            }
        }
    }
}

cudaError_t configure_trace()   // per device, once per context
{
    return cudaFuncSetAttribute(trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

size_t trace_smem_bytes(int n_slots)
{
    return (size_t)n_slots * 16 + (size_t)n_slots * 4 + (size_t)4 * kRays * kThreads * 4;
}

cudaError_t launch_hoist(cudaStream_t st, const rtc_object* objs, const int32_t* sphere_obj, int n_spheres,
                         int n_slots, const float cam[3], float4* sph_pairs, float* sph_c,
                         unsigned int* counters, int n_counters)
{
    const int n = n_slots > n_counters ? n_slots : n_counters;
    hoist_kernel<<<(n + 255) / 256, 256, 0, st>>>(objs, sphere_obj, n_spheres, n_slots, cam[0], cam[1], cam[2],
                                                  sph_pairs, sph_c, counters, n_counters);
    return cudaGetLastError();
}

cudaError_t launch_trace(cudaStream_t st, int n_ctas, const FrameParams& fp, const float4* g_pairs, const float* g_c,
                         const int32_t* sphere_obj, int n_spheres, int n_slots, const rtc_object* objs,
                         const int32_t* plane_obj, int n_planes, float* hit_t, int32_t* hit_idx,
                         unsigned int* tile_counter, int carry_in)
{
    trace_kernel<<<n_ctas, kThreads, trace_smem_bytes(n_slots), st>>>(fp, g_pairs, g_c, sphere_obj, n_spheres, n_slots,
                                                                      objs, plane_obj, n_planes, hit_t, hit_idx,
                                                                      tile_counter, carry_in);
    return cudaGetLastError();
}

}  // namespace rtc
