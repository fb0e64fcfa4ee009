// rtc_trace.cu -- kernel 0 (per-frame scene hoist) and kernel 1 (ray generation + nearest hit).
//
// Replaces the hot loop of the reference's RayTrace (RayTracing.cu:81-136) and the
// Sphere::Trace / Plane::Trace it calls per object (Sphere.cu:30-68, Plane.cu:38-73).
//
// Design (B200 / sm_100a), driven by on-box microbenchmarks (profiles/r01_microbench.md):
//   * All primary rays share the origin (RayTracing.cu:195), so oc = origin - centre and
//     c = |oc|^2 - r^2 are per-sphere, per-frame constants: kernel 0 hoists them (bit-identical
//     to what the reference recomputes per ray).
//   * Kernel 1 is persistent: one CTA of 24 or 28 warps per SM (plan_trace picks from the number of tiles)
//     keeps the sphere list (2496 / 1772 spheres per launch; longer lists are chunked) in shared memory and
//     walks 16x16-pixel screen tiles, one tile per warp, 8 rays per thread.
//   * The inner loop tests TWO spheres against one ray per packed instruction (FMUL2/FFMA2).
//     Measured on B200: an FFMA2 only sustains 1 per 2 cycles when at most one operand pair is
//     fresh (register-bank limit), and every ALU-pipe instruction (FMNMX3, FSETP, ...) costs
//     ~2.8 FMA-pipe cycles.  So (a) the loop is operand-major -- 8 consecutive packed ops share
//     one shared-memory operand (register reuse cache), and (b) hit detection uses NO per-test
//     ALU work: the sphere vector is pre-scaled by 2^64/sqrt(c') and the ray by 2^64, so
//     u = d.oc overflows to +-inf exactly when |d.oc| >= sqrt(c') (the discriminant is >= 0);
//     a NaN-sticky accumulator A = fma(u, 0, A) collects the 16 tests of an iteration on the FMA
//     pipe and is examined once (one FSETP) per iteration.
//   * The packed test is only a CONSERVATIVE filter (c' is c deflated by a rounding-error
//     bound); survivors are re-evaluated with the reference's exact operation order, so hit
//     decisions and distances are bit-identical to the reference.
//   * The running best (distance, object index) per ray lives in shared memory: it is touched
//     only on the (rare) exact path and would otherwise cost 16 registers in the hot loop.
//   * Sphere slots are in Morton order (host), 4 neighbours per packed group; per group the hoist also
//     produces a lower bound of any hit distance (rays that already have a nearer hit skip the group's
//     exact path) and a bounding cone (RTC_FLAG_CULL: warp tiles skip groups their ray cone misses).
//   * The shadow-ray extension is the same kernel with the light as the common origin (SHADOW = true).
#include <cstdlib>

#include "rtc_device.cuh"
#include "rtc_kernels.h"
#include "rtc_shade.cuh"

namespace rtc {

// Relative deflation of c for the conservative filter.  A reference hit needs
// fl(fl(s^2)) >= fl(a*c) with s the un-fused dot product and a = d.d; against the fused,
// scaled u this is implied by |u| >= 2^128 once c is deflated by >= 30.4 * 2^-24 * |oc|^2
// (DESIGN.md "filter bound"); the screen-affine form of the filter (below) needs 44.6 * 2^-24 = 2.7e-6; 1e-5 leaves
// 3.7x headroom.
#define RTC_FILTER_EPS 1.0e-5f
#define RTC_TWO64 18446744073709551616.0f

// Which packed filter the primary pass will run, and the columns of the inverse view matrix it needs.
struct HoistBasis {
    int affine;          // 1: store (A, B, C) = g . (col0, col1, col2) per sphere; 0: store g itself
    float eps;           // deflation of c: RTC_FILTER_EPS, or the packet filter's wider one (packet_eps)
    float c0[3], c1[3], c2[3];
};

// ---- kernel 1: trace --------------------------------------------------------------------
// Rays per thread are a template parameter (kRays = 8 or 4): a thread owns one column and kRays consecutive rows (the
// upper or the lower half) of its warp's 16 x 2*kRays tile.  8 is the efficient shape (the per-column operand F of the screen-affine filter is
// amortised over 8 tests); 4 halves the work quantum for launches of only one or two tile waves (an 8-GPU band of a 4K
// frame), where the phases of a warp's single tile -- ray set-up, sphere loop, shading epilogue -- would otherwise run in
// lock step on all warps and not overlap (plan_trace).
// Threads per CTA are a template parameter (compile-time state stride; a run-time stride cost 3 %): 768 (24 warps,
// 80 registers) and 896 (28 warps, 72 registers) are instantiated and plan_trace picks per launch.
constexpr int kTile = 16;           // warp tile = 16 columns x 2*kRays rows

// Shared-memory layout (dynamic): [exact float4 x n_slots][fast 12 B x n_slots][group dmin 4 B x n_slots/4][state]
// state, each [kRays][blockDim.x]: best_t, best_idx, divTwoA, and the exact ray direction
// (x, y, z) -- the rare exact path indexes rays dynamically, which registers cannot do.
struct ShadeCtx {
    ShadeParams sp;
    float cam[3];
    float far_dist;
    const rtc_object* objs;
    const float4* kd;
    int mode;
    int pad_;
};
static_assert(sizeof(ShadeCtx) <= 96, "ShadeCtx must fit the 96 bytes set aside in the shared-memory budget");
struct Smem {
    float4* exact;
    float4* fast;      // 3 float4 per group of 4 spheres
    float* gdmin;      // per group of 4 spheres: lower bound of any hit distance
    float4* gcone;     // per group: bounding cone seen from the ray origin (axis, cos half-angle) ...
    float* gsin;       // ... and the sine of its half-angle
    float* best_t;
    int* best_idx;
    float* div2A;
    float* dirx;
    float* diry;
    float* dirz;
    struct ShadeCtx* shade;   // tile epilogue: light / material block, camera, mode (one per CTA)
};
__device__ __forceinline__ Smem carve(unsigned char* raw, int n_slots, int n_threads, int n_rays)
{
    Smem s;
    s.exact = reinterpret_cast<float4*>(raw);
    s.fast = s.exact + n_slots;
    const int ng4 = ((n_slots >> 2) + 3) & ~3;
    s.gcone = s.fast + (n_slots >> 2) * 3;
    s.gdmin = reinterpret_cast<float*>(s.gcone + ng4);
    s.gsin = s.gdmin + ng4;
    float* st = s.gsin + ng4;
    const int n = n_rays * n_threads;
    s.best_t = st;
    s.best_idx = reinterpret_cast<int*>(st + n);
    s.div2A = st + 2 * n;
    s.dirx = st + 3 * n;
    s.diry = st + 4 * n;
    s.dirz = st + 5 * n;
    s.shade = reinterpret_cast<ShadeCtx*>(st + 6 * n);
    return s;
}

__device__ __forceinline__ float sqrt_approx(float x)   // MUFU.SQRT, rel. error < 2^-22
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ bool not_finite(float x) { return !(fabsf(x) <= 3.402823466e+38f); }
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ---- per-frame scene hoist, done by every CTA of the ray kernel for its own shared-memory copy -----------------------
// (Round 1 and the first half of round 2 ran this as a launch of its own that wrote global arrays every CTA then staged;
// a 1000-sphere list is one slot per thread of a CTA, so recomputing it 148 times costs nothing and saves a launch, its
// gap, and a global round trip per frame -- 5 % of an 8-GPU band, 20 % of a console-sized frame.)
// Per sphere slot j (slots are padded to a multiple of 4; a warp's 4-lane groups see 4 consecutive slots):
//   fast : per GROUP of 4 spheres 12 floats: gx[4], gy[4], gz[4] with g = oc * 2^64/sqrt(c') -- or, screen-affine form,
//          (A, B, C)[4] = g . (col0, col1, col2) -- (three LDS.128 feed two packed sphere pairs)
//   exact: per sphere float4 (ocx, ocy, ocz, c), exact.
// Slots past n_spheres are never-hit sentinels (g = 0).  c' <= 0 (origin inside / on the sphere) or non-finite geometry
// => g = inf: always a candidate, the exact path decides.
__device__ __noinline__ void hoist_into_smem(const Smem sm, const rtc_object* __restrict__ objs, const int32_t* __restrict__ sphere_obj,
                                             int n_spheres, int n_slots, float camx, float camy, float camz, const HoistBasis hb,
                                             bool want_cone, int tid, int n_threads)
{
    for (int j0 = 0; j0 < n_slots; j0 += n_threads) {          // n_threads is a multiple of 32: warp-uniform trip count
    const int j = j0 + tid;
    // (no early return: the group minimum below is a warp shuffle; n_slots is a multiple of 4, blockDim of 32)
    float ocx = 0.f, ocy = 0.f, ocz = 0.f, c = 1.f, gx = 0.f, gy = 0.f, gz = 0.f;
    float dmin = 3.0e38f;                     // lower bound of any reference hit distance on this sphere
    float wx = 0.f, wy = 0.f, wz = 0.f, wr = 0.f, wn = 0.f;   // world centre, radius, 1 if this slot holds a sphere
    if (j < n_spheres) {
        const rtc_object& s = objs[sphere_obj[j]];
        wx = s.center[0]; wy = s.center[1]; wz = s.center[2]; wr = fabsf(s.radius); wn = 1.0f;
        ocx = sub(camx, s.center[0]);                                   // Sphere.cu:34
        ocy = sub(camy, s.center[1]);
        ocz = sub(camz, s.center[2]);
        const float oc2 = vdot(v3(ocx, ocy, ocz), v3(ocx, ocy, ocz));
        c = sub(oc2, mul(s.radius, s.radius));                          // Sphere.cu:37
        const float cd = fmaf(-hb.eps, oc2, c);                         // deflated c'
        const float inf = __int_as_float(0x7f800000);
        if (cd > 0.0f && cd < 3.0e38f) {
            if (hb.affine) {
                // screen-affine filter: (A, B, C) = g . (col0, col1, col2) of the inverse view matrix, g = oc 2^64 / sqrt(c'),
                // in binary64 and rounded once (stored where the dot-product form keeps gx, gy, gz)
                const double gs = 18446744073709551616.0 / sqrt((double)cd);
                const double dx = (double)ocx * gs, dy = (double)ocy * gs, dz = (double)ocz * gs;
                gx = (float)(dx * (double)hb.c0[0] + dy * (double)hb.c0[1] + dz * (double)hb.c0[2]);
                gy = (float)(dx * (double)hb.c1[0] + dy * (double)hb.c1[1] + dz * (double)hb.c1[2]);
                gz = (float)(dx * (double)hb.c2[0] + dy * (double)hb.c2[1] + dz * (double)hb.c2[2]);
            } else {
                const float g = RTC_TWO64 / sqrtf(cd);
                gx = ocx * g; gy = ocy * g; gz = ocz * g;
            }
            // Lower bound of the reference's ROUNDED hit distance over every ray: the minimum over s1 of exact_group's
            // per-candidate bound t_lb(s1) (same discriminant slack: the rounding error of b*b - 4ac, ~23u|oc|^2, is
            // amplified by the sqrt, so near-tangent hits of small or distant spheres come out up to ~3e-3|oc| NEARER
            // than the geometric |oc| - r).  With p = -s1 and K = 3e-6(|oc|^2 + |c|) - c < 0 (implied by c' > 0),
            // t_lb(p) = p - sqrt(p^2 + K)(1 + 1e-6) - 2e-6(p + |oc|) decreases in p, so its minimum sits at the largest
            // p a unit direction (to 1e-6) allows, p = |oc|(1 + 1e-6).  Evaluated in binary64, then rounded down.
            const double L = sqrt((double)oc2) * 1.000001;
            const double K = 3.0e-6 * ((double)oc2 + fabs((double)c)) - (double)c;
            const double q = fmax(L * L + K, 0.0);
            const double lb = (L - sqrt(q) * 1.000001) - 2.0e-6 * (2.0 * L);
            dmin = fmaxf((float)(lb * (lb >= 0.0 ? 0.999999 : 1.000001)) - 1.0e-30f, 0.0f);
#ifdef RTC_TEST_R01_GROUP_BOUND   // the round-1 bound (geometric |oc| - r): kept only to show that the regression test bites
            dmin = fmaxf((sqrtf(oc2) - fabsf(s.radius)) - 1.0e-5f * (sqrtf(oc2) + fabsf(s.radius)), 0.0f);
#endif
        } else {
            gx = gy = gz = inf;
            dmin = 0.0f;
        }
    }
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, 1));
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, 2));
    // Bounding sphere of the group of 4 (centre = mean of the centres, radius = max(|c_i - centre| + r_i)) as a cone
    // seen from the ray origin: unit axis u, sin and cos of its half-angle, both rounded towards "wider".
    // No sphere in the group: never kept.  Origin inside the bounding sphere or odd numbers: always kept.
    float sx = wx, sy = wy, sz = wz, sn = wn;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o); sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    const float inv_n = sn > 0.0f ? 1.0f / sn : 0.0f;
    const float mx = sx * inv_n, my = sy * inv_n, mz = sz * inv_n;
    float R = wn > 0.0f ? sqrtf((wx - mx) * (wx - mx) + (wy - my) * (wy - my) + (wz - mz) * (wz - mz)) + wr : 0.0f;
    R = fmaxf(R, __shfl_xor_sync(0xffffffffu, R, 1));
    R = fmaxf(R, __shfl_xor_sync(0xffffffffu, R, 2));
    if (j < n_slots) {
    if (want_cone && (j & 3) == 0) {
        float4 cone = make_float4(0.f, 0.f, 0.f, 3.0f);          // empty group: threshold 3 cos(theta) > 1 >= any dot
        float sn_a = 0.0f;
        if (sn > 0.0f) {
            const float vx = mx - camx, vy = my - camy, vz = mz - camz;
            const float L = sqrtf(vx * vx + vy * vy + vz * vz);
            const float Ri = R * 1.00002f + 1.0e-5f * (L + R);   // inflated: covers the rounding of everything above
            if (L > Ri && L < 3.0e37f && Ri == Ri) {
                const float sa = fminf(Ri / L * 1.000001f + 1.0e-7f, 1.0f);
                const float ca = sqrtf(fmaxf(1.0f - sa * sa, 0.0f)) * 0.999999f;
                cone = make_float4(vx / L, vy / L, vz / L, ca);
                sn_a = sa;
            } else {
                cone = make_float4(0.f, 0.f, 0.f, -3.0f);        // always kept: threshold < 0 <= dot = 0
            }
        }
        sm.gcone[j >> 2] = cone;
        sm.gsin[j >> 2] = sn_a;
    }
    float* base = reinterpret_cast<float*>(sm.fast) + 12 * (j >> 2);
    const int k = j & 3;
    base[k] = gx; base[4 + k] = gy; base[8 + k] = gz;
    sm.exact[j] = make_float4(ocx, ocy, ocz, c);
    if (k == 0) sm.gdmin[j >> 2] = dmin;
    }
    }
}

// The rare path: exact re-evaluation of the candidates of one loop iteration (4 spheres x 8 rays).
// `mask` bit (q*16 + r*2 + h) <=> ray r may hit sphere 4g + 2q + h.
// Accept rule == reference RayTracing.cu:123 (strict '<', lowest index wins ties), written
// order-independently as the lexicographic minimum of (distance, object index).
// A cheap, rigorous lower bound of the hit distance skips candidates that cannot beat the
// running best (most of them: a ray pierces ~N/100 spheres but only the nearest matters).
template <int kThreads, int kRays>
__device__ __noinline__ void exact_group(const int32_t* __restrict__ sphere_obj, int n_slots, int g, uint32_t mask, int tid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, n_slots, kThreads, kRays);
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const int r = (b >> 1) & 7;
        const int j = 4 * g + 2 * (b >> 4) + (b & 1);
        const int slot = r * kThreads + tid;
        const float4 e = s.exact[j];
        const float dx = s.dirx[slot], dy = s.diry[slot], dz = s.dirz[slot];
        const float best = s.best_t[slot];
        if (best < 0.0f) continue;                              // inactive ray (shadow pass: no primary hit to shade)
        {
            // t_ref ~ (-s - sqrt(D))/a with a = 1 +- 8u, |s1 - s| <= 6u|oc|, and
            // D_ref <= s1^2 - c + 23u(|oc|^2 + |c|), so t_lb <= t_ref (DESIGN.md "reject bound").
            const float s1 = fmaf(dz, e.z, fmaf(dy, e.y, dx * e.x));
            const float oc2 = fmaf(e.z, e.z, fmaf(e.y, e.y, e.x * e.x));
            const float q1 = fmaxf(fmaf(s1, s1, fmaf(3.0e-6f, oc2 + fabsf(e.w), -e.w)), 0.0f);
            const float t_lb = (-s1 - sqrt_approx(q1) * 1.000001f) - 2.0e-6f * (fabsf(s1) + sqrt_approx(oc2) * 1.000001f);
            if (t_lb * (t_lb >= 0.0f ? 0.999999f : 1.000001f) > best) continue;
        }
        const V3 d = v3(dx, dy, dz);
        const float fourA = mul(4.0f, vdot(d, d));                       // RayTracing.cu:91-92
        float t;
        if (!sphere_trace_hoisted(e.x, e.y, e.z, e.w, d, fourA, s.div2A[slot], t)) continue;
        if (t < best) { s.best_t[slot] = t; s.best_idx[slot] = sphere_obj[j]; continue; }
        if (t == best) {
            const int oi = sphere_obj[j];
            if (oi < s.best_idx[slot]) s.best_idx[slot] = oi;
        }
    }
}

// Tile epilogue, per ray: shade + quantise (rtc_shade.cuh).  Out of line so that its registers (binary64 pow, IEEE
// division / square-root sequences) do not weigh on the hot loop; executed only by warps with a hit among 32 rays.
__device__ __noinline__ uint32_t shade_call(const ShadeCtx* __restrict__ sc, float dx, float dy, float dz, float t, int idx)
{
    const int mode = sc->mode;
    return shade_pixel(mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL, mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII, mode,
                       sc->sp, sc->objs, sc->kd, v3(sc->cam[0], sc->cam[1], sc->cam[2]), sc->far_dist, v3(dx, dy, dz), t, idx, false);
}

// The filter flagged something in group g: per-ray candidate mask from the non-finite u, group reject, exact path.
template <int kThreads, int kRays>
__device__ __forceinline__ void resolve_candidates(const Smem& s, const f32x2 (&u)[2][kRays], int g, const int32_t* __restrict__ sphere_obj,
                                                   int n_slots, int tid)
{
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int r = 0; r < kRays; ++r) {
            float lo, hi;
            unpack2(u[q][r], lo, hi);
            mask |= not_finite(lo) ? (1u << (q * 16 + r * 2)) : 0u;
            mask |= not_finite(hi) ? (1u << (q * 16 + r * 2 + 1)) : 0u;
        }
    }
    if (mask) {
        // Rays whose running best is already nearer than anything in this group of 4 spheres can offer
        // drop out here (most candidates of a ray lie behind its nearest hit).
        const float gd = s.gdmin[g];
#pragma unroll
        for (int r = 0; r < kRays; ++r)
            if (gd > s.best_t[r * kThreads + tid]) mask &= ~(0x00030003u << (2 * r));
        if (mask) exact_group<kThreads, kRays>(sphere_obj, n_slots, g, mask, tid);
    }
}

// One group of 4 spheres (2 packed pairs) against the thread's 8 rays, operand-major, + 3 LDS.128 + one NaN check.
//
// AFFINE = false (shadow rays; primary rays under an ill-conditioned view matrix): u = 2^64 d . g, 3 packed ops per test
//   pair + the sticky accumulate: 64 packed ops per group; measured 4.37 cycles/test against the 4.0 of a pure FFMA2
//   stream (profiles/r01_microbench.md).
// AFFINE = true (primary rays): the SCREEN-AFFINE form of the same filter.  A primary ray's un-normalised direction is
//   affine in its view-space coordinates, w = col2 + vx col0 + vy col1 (RayTracing.cu:20-22), so
//       2^64 d . g = (2^64 / |w|) (C + vx A + vy B),   (A, B, C) = g . (col0, col1, col2)  per sphere and frame (hoisted).
//   A thread's 8 rays share their column, i.e. vx: F = C + vx A is ONE packed op per sphere pair and thread, and a test
//   is t = F + vy_r B, u = t (2^64 / |w_r|) -- 2 packed ops per pair + the sticky accumulate: 50 packed ops per group
//   instead of 64.  The operands of the dot-product form (ex, ey, ez) become (vx, vy_r, 2^64 / |w_r|) here.
template <int kThreads, bool AFFINE, int kRays>
__device__ __forceinline__ void test_group(const Smem& s, uint32_t fa, int g, const float (&ex)[kRays], const float (&ey)[kRays],
                                           const float (&ez)[kRays], const int32_t* __restrict__ sphere_obj, int n_slots, int tid)
{
    const f32x2 ZERO2 = pack2(0.0f, 0.0f);
    const float4 FX = lds128(fa), FY = lds128(fa + 16u), FZ = lds128(fa + 32u);   // warp-broadcast
    f32x2 u[2][kRays];
    f32x2 acc0 = ZERO2, acc1 = ZERO2;                           // NaN-sticky "something overflowed"
    if (AFFINE) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const f32x2 A = q ? pack2(FX.z, FX.w) : pack2(FX.x, FX.y);
            const f32x2 B = q ? pack2(FY.z, FY.w) : pack2(FY.x, FY.y);
            const f32x2 C = q ? pack2(FZ.z, FZ.w) : pack2(FZ.x, FZ.y);
            const f32x2 F = fma2(pack2(ex[0], ex[0]), A, C);                                      // FFMA2 x1 per pair
#pragma unroll
            for (int r = 0; r < kRays; ++r) u[q][r] = fma2(pack2(ey[r], ey[r]), B, F);            // FFMA2 x8, B and F reused
        }
#pragma unroll
        for (int r = 0; r < kRays; ++r) {                                                         // FMUL2 x16: +-inf <=> candidate
            const f32x2 S = pack2(ez[r], ez[r]);                                                  // (S reused by the pair of ops)
            u[0][r] = mul2(u[0][r], S);
            u[1][r] = mul2(u[1][r], S);
        }
#pragma unroll
        for (int r = 0; r < kRays; ++r) {                                                         // FFMA2 x16: inf*0 -> NaN, sticky
            acc0 = fma2(u[0][r], ZERO2, acc0);
            acc1 = fma2(u[1][r], ZERO2, acc1);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const f32x2 GX = q ? pack2(FX.z, FX.w) : pack2(FX.x, FX.y);
            const f32x2 GY = q ? pack2(FY.z, FY.w) : pack2(FY.x, FY.y);
            const f32x2 GZ = q ? pack2(FZ.z, FZ.w) : pack2(FZ.x, FZ.y);
#pragma unroll
            for (int r = 0; r < kRays; ++r) u[q][r] = mul2(pack2(ex[r], ex[r]), GX);             // FMUL2 x8, GX reused
#pragma unroll
            for (int r = 0; r < kRays; ++r) u[q][r] = fma2(pack2(ey[r], ey[r]), GY, u[q][r]);    // FFMA2 x8, GY reused
#pragma unroll
            for (int r = 0; r < kRays; ++r) u[q][r] = fma2(pack2(ez[r], ez[r]), GZ, u[q][r]);    // FFMA2 x8: +-inf <=> candidate
#pragma unroll
            for (int r = 0; r < kRays; ++r) {                                                     // FFMA2 x8: inf*0 -> NaN, sticky
                if (r & 1) acc1 = fma2(u[q][r], ZERO2, acc1);
                else acc0 = fma2(u[q][r], ZERO2, acc0);
            }
        }
    }
    float alo, ahi;
    unpack2(add2(acc0, acc1), alo, ahi);
    if (__any_sync(0xffffffffu, !(alo == ahi)))                  // NaN in either half (both are 0 otherwise)
        resolve_candidates<kThreads, kRays>(s, u, g, sphere_obj, n_slots, tid);
}

// ---- the PACKET form of the screen-affine filter (RTC_FLAG_PACKET) ------------------------------------------------------
// A thread's kRays rays share their column and sit on consecutive rows, so along the packet only vy moves and
//     h(vy) = d(vy) . oc / |oc|,   d = w / |w|,   w = (col2 + vx col0) + vy col1
// is a smooth function of ONE variable with |h''| <= 5 |col1|^2 / |w|^2 <= 5.12 (3x3 orthonormal to 1e-3: |w| >= 0.999).
// On an interval of length D a function stays within M D^2 / 8 of its chord, and a chord is largest at an end point:
//     max_r |h(vy_r)|  <=  max(|h(vy_first)|, |h(vy_last)|) + 0.64 D^2 ,      D = |vy_first - vy_last|.
// A reference hit on ray r needs |h(vy_r)| >= sqrt(kappa - eps0), kappa = c / |oc|^2, eps0 = 44.6 u (DESIGN.md "filter
// bound"); sqrt(kappa - eps0) - sqrt(kappa - E) >= (E - eps0) / 2 for kappa <= 1, so with the spheres hoisted under the
// deflation E = 1e-5 + 1.5 D^2 (packet_eps; >= eps0 + 2 * 0.64 D^2 with the per-ray filter's own head-room kept) the
// per-ray filter evaluated on the FIRST and the LAST ray of the packet alone overflows whenever any ray of the packet
// can hit: 7 packed ops per sphere pair and 8 rays (F; t and u at both ends; two sticky accumulates) instead of 25.
// Flagged groups (about one warp iteration in a hundred) run the per-ray filter of test_group out of line, on the same
// operands, and go on to the exact path as before: the accepted hits are the reference's, bit for bit.
// (vy_r is monotone in r -- the roundings of cy and vy are monotone, clamped rows repeat the last row -- so every ray of
// the packet lies between its ends.)
template <int kRays>
struct RayOps {
    float vx;
    float vy[kRays];
    float sc[kRays];       // 2^64 / |w_r|
};

template <int kThreads, int kRays>
__device__ __noinline__ void packet_slow_group(uint32_t fa, int g, const RayOps<kRays> ro, const int32_t* __restrict__ sphere_obj, int n_slots, int tid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, n_slots, kThreads, kRays);
    const float4 FX = lds128(fa), FY = lds128(fa + 16u), FZ = lds128(fa + 32u);
    f32x2 u[2][kRays];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const f32x2 A = q ? pack2(FX.z, FX.w) : pack2(FX.x, FX.y);
        const f32x2 B = q ? pack2(FY.z, FY.w) : pack2(FY.x, FY.y);
        const f32x2 C = q ? pack2(FZ.z, FZ.w) : pack2(FZ.x, FZ.y);
        const f32x2 F = fma2(pack2(ro.vx, ro.vx), A, C);
#pragma unroll
        for (int r = 0; r < kRays; ++r) u[q][r] = mul2(fma2(pack2(ro.vy[r], ro.vy[r]), B, F), pack2(ro.sc[r], ro.sc[r]));
    }
    resolve_candidates<kThreads, kRays>(s, u, g, sphere_obj, n_slots, tid);
}

// NaN-sticky flag of one group of 4 spheres against the two end rays of the thread's packet: 14 packed ops + 3 LDS.128.
struct PacketEnds { f32x2 vx, vy0, vy1, s0, s1; };
__device__ __forceinline__ f32x2 packet_flag(uint32_t fa, const PacketEnds& e)
{
    const f32x2 ZERO2 = pack2(0.0f, 0.0f);
    const float4 FX = lds128(fa), FY = lds128(fa + 16u), FZ = lds128(fa + 32u);   // warp-broadcast
    f32x2 acc0 = ZERO2, acc1 = ZERO2;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const f32x2 A = q ? pack2(FX.z, FX.w) : pack2(FX.x, FX.y);
        const f32x2 B = q ? pack2(FY.z, FY.w) : pack2(FY.x, FY.y);
        const f32x2 C = q ? pack2(FZ.z, FZ.w) : pack2(FZ.x, FZ.y);
        const f32x2 F = fma2(e.vx, A, C);
        const f32x2 u0 = mul2(fma2(e.vy0, B, F), e.s0);
        const f32x2 u1 = mul2(fma2(e.vy1, B, F), e.s1);
        acc0 = fma2(u0, ZERO2, acc0);                          // inf * 0 -> NaN, sticky
        acc1 = fma2(u1, ZERO2, acc1);
    }
    return add2(acc0, acc1);
}

// A flagged group goes to the per-ray filter only if some lane that flagged it still has a ray the group could beat (the
// group reject of resolve_candidates, applied before the 50 packed operations instead of after: three flagged groups in
// four lie behind what their rays have already hit).
template <int kThreads, int kRays>
__device__ __forceinline__ void packet_resolve(const Smem& s, uint32_t fa, int g, f32x2 f, const RayOps<kRays>& ro,
                                               const int32_t* __restrict__ sphere_obj, int n_slots, int tid)
{
    float lo, hi;
    unpack2(f, lo, hi);
    bool mine = !(lo == hi);                                     // NaN in either half (both are 0 otherwise)
    if (mine) {
        const float gd = s.gdmin[g];
        float bmax = s.best_t[tid];
#pragma unroll
        for (int r = 1; r < kRays; ++r) bmax = fmaxf(bmax, s.best_t[r * kThreads + tid]);
        mine = !(gd > bmax);
    }
    if (__any_sync(0xffffffffu, mine)) packet_slow_group<kThreads, kRays>(fa, g, ro, sphere_obj, n_slots, tid);
}

// N groups (gs[0 .. N)) per NaN check.
template <int kThreads, int kRays, int N>
__device__ __forceinline__ void test_packet_groups(const Smem& s, uint32_t fast_base, const int (&gs)[N], const PacketEnds& e,
                                                   const RayOps<kRays>& ro, const int32_t* __restrict__ sphere_obj, int n_slots, int tid)
{
    f32x2 f[N];
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = packet_flag(fast_base + 48u * (uint32_t)gs[i], e);
    f32x2 sum = f[0];
#pragma unroll
    for (int i = 1; i < N; ++i) sum = add2(sum, f[i]);
    float alo, ahi;
    unpack2(sum, alo, ahi);
    if (__any_sync(0xffffffffu, !(alo == ahi))) {                // NaN somewhere (both halves are 0 otherwise)
#pragma unroll
        for (int i = 0; i < N; ++i) packet_resolve<kThreads, kRays>(s, fast_base + 48u * (uint32_t)gs[i], gs[i], f[i], ro, sphere_obj, n_slots, tid);
    }
}

// SHADOW = false: primary rays from the camera (hit_t / hit_idx are the outputs).
// SHADOW = true : the shadow-ray EXTENSION (the reference casts none).  One ray per shaded pixel, cast FROM
//   THE LIGHT (1,50,0) toward the shaded point P' = P + n*1e-3: all shadow rays then share their origin exactly like
//   the primary rays share the camera, so the same hoist, the same packed filter and the same exact path apply
//   unchanged (the "camera" of this launch is the light).  A pixel is in shadow iff some object is hit at a distance
//   strictly below |P' - light| -- the nearest-hit machinery with the running best initialised to that length and
//   no object.  hit_t / hit_idx are INPUTS here; the output is one byte per pixel in `shadow`.
// CULL = true (RTC_FLAG_CULL): per warp tile, the groups of 4 spheres whose bounding cone (hoisted) misses the tile's
//   ray cone are skipped -- results are identical, far fewer tests are executed (the count is reported).
// AFFINE = true: the screen-affine packed filter (see test_group); primary rays only.
// PACKET = true (RTC_FLAG_PACKET; needs AFFINE): the filter runs on the two end rays of every thread's packet (test_packet_groups).
template <bool SHADOW, int kThreads, bool CULL, bool AFFINE, int kRays, bool PACKET>
__global__ void __launch_bounds__(kThreads, 1)
trace_kernel(const FrameParams fp, const HoistBasis hb,
             const int32_t* __restrict__ sphere_obj, int n_spheres, int n_slots,
             const rtc_object* __restrict__ objs, const int32_t* __restrict__ plane_obj, int n_planes,
             float* __restrict__ hit_t, int32_t* __restrict__ hit_idx, unsigned long long* __restrict__ tile_counter,
             unsigned long long ticket_base /* tickets this counter has handed out before this launch (never reset) */,
             int carry_in /* 1: continue from hit_t/hit_idx (sphere list chunking) */,
             float lx, float ly, float lz, uint8_t* __restrict__ shadow, unsigned long long* __restrict__ groups_tested,
             unsigned long long* __restrict__ stats_zero /* the other frame parity's two counts: zeroed here for the next frame */,
             const ShadeParams sp, int shade_mode /* >= 0: shade + quantise into color / glyph in the tile epilogue */,
             uint8_t* __restrict__ color, uint8_t* __restrict__ glyph, int write_hits, const float4* __restrict__ obj_kd)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, n_slots, kThreads, kRays);
    constexpr uint32_t kTileH = 2u * kRays;                      // tile height in rows
    const int tid = threadIdx.x, lane = tid & 31;
    if (!SHADOW && tid == 0) {
        ShadeCtx* sc = s.shade;
        sc->sp = sp;
        sc->cam[0] = fp.cam[0]; sc->cam[1] = fp.cam[1]; sc->cam[2] = fp.cam[2];
        sc->far_dist = fp.far_dist; sc->objs = objs; sc->kd = obj_kd; sc->mode = shade_mode;
    }
    unsigned int my_groups = 0;                                  // groups of 4 spheres this warp ran the packed test on

    if (blockIdx.x == 0 && tid == 0 && stats_zero != nullptr) { stats_zero[0] = 0ull; stats_zero[1] = 0ull; }
    // Hoist this launch's sphere slots into shared memory, once per CTA (persistent kernel).
    hoist_into_smem(s, objs, sphere_obj, n_spheres, n_slots, SHADOW ? lx : fp.cam[0], SHADOW ? ly : fp.cam[1], SHADOW ? lz : fp.cam[2],
                    hb, CULL, tid, kThreads);
    __syncthreads();

    const uint32_t W = fp.x - 1u;
    const uint32_t rows = fp.row1 - fp.row0;
    const uint32_t tiles_x = (W + kTile - 1) / kTile, tiles_y = (rows + kTileH - 1) / kTileH;
    const uint32_t n_tiles = tiles_x * tiles_y;
    const int n_groups = n_slots >> 2;
    const V3 cam = v3(fp.cam[0], fp.cam[1], fp.cam[2]);
    const V3 o = SHADOW ? v3(lx, ly, lz) : cam;                 // common origin of this launch's rays
    const uint32_t px = lane & 15u, py = lane >> 4;
    const uint32_t fast_base = (uint32_t)__cvta_generic_to_shared(s.fast);

    for (;;) {
        unsigned long long ticket = 0ull;
        if (lane == 0) ticket = atomicAdd(tile_counter, 1ull) - ticket_base;     // (every warp draws one ticket past the end)
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= (unsigned long long)n_tiles) break;
        const uint32_t tile = (uint32_t)ticket;
        const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const uint32_t col = tx * kTile + px;
        const uint32_t colc = col < W ? col : W - 1u;     // clamp: out-of-frame lanes trace a duplicate ray

        float ex[kRays], ey[kRays], ez[kRays];            // ray direction scaled by 2^64 (exact)
#pragma unroll
        for (int r = 0; r < kRays; ++r) {
            uint32_t row = fp.row0 + ty * kTileH + py * kRays + r;
            if (row >= fp.row1) row = fp.row1 - 1u;
            float vx, vy, inv;
            V3 d = initial_direction_ex(fp, row, colc, vx, vy, inv);
            float bt = 99999999.f;                                        // RayTracing.h:21
            int bi = -1;
            const size_t pix = (size_t)(row - fp.row0) * W + colc;
            if (!SHADOW) {
                if (carry_in) { bt = hit_t[pix]; bi = hit_idx[pix]; }
            } else {
                const float t = hit_t[pix];
                const int idx = hit_idx[pix];
                bt = -1.0f;                                               // inactive unless there is a point to shade
                if (t <= fp.far_dist && idx >= 0 && !(carry_in && shadow[pix])) {
                    const rtc_object ob = objs[idx];
                    const V3 point = vadd(cam, vscale(d, t));             // RayTracing.cu:149
                    V3 n;
                    if (ob.type == RTC_OBJ_SPHERE)                        // Sphere.cu:67
                        n = vnormalize(vsub(point, v3(ob.center[0], ob.center[1], ob.center[2])));
                    else
                        n = v3(ob.normal[0], ob.normal[1], ob.normal[2]); // Plane.cu:72
                    n = vnormalize_unit(n);                               // RayTracing.cu:129
                    const V3 lp = vsub(vadd(point, vscale(n, 1.0e-3f)), o);
                    const float len = vlength(lp);
                    d = vscale(lp, rcp(len));
                    bt = len;
                } else {
                    d = v3(0.0f, 0.0f, 0.0f);
                }
            }
            if (AFFINE) { ex[r] = vx; ey[r] = vy; ez[r] = inv * RTC_TWO64; }                  // (vx is the same for all 8 rays)
            else { ex[r] = d.x * RTC_TWO64; ey[r] = d.y * RTC_TWO64; ez[r] = d.z * RTC_TWO64; }
            const int slot = r * kThreads + tid;
            s.dirx[slot] = d.x; s.diry[slot] = d.y; s.dirz[slot] = d.z;
            s.div2A[slot] = rcp(mul(2.0f, vdot(d, d)));                   // RayTracing.cu:91,93
            s.best_t[slot] = bt;
            s.best_idx[slot] = bi;
        }

        if (SHADOW) {                                           // tiles without a shaded pixel need no shadow rays
            bool any = false;
#pragma unroll
            for (int r = 0; r < kRays; ++r) any |= s.best_t[r * kThreads + tid] >= 0.0f;
            if (!__any_sync(0xffffffffu, any)) {
#pragma unroll
                for (int r = 0; r < kRays; ++r) {
                    const uint32_t row = fp.row0 + ty * kTileH + py * kRays + r;
                    if (row < fp.row1 && col < W && !carry_in) shadow[(size_t)(row - fp.row0) * W + col] = 0;
                }
                continue;
            }
        }

        // ---- hot loop over groups of 4 spheres ---------------------------------------------------------
        PacketEnds pe;
        RayOps<kRays> ro;
        if (PACKET) {
            pe.vx = pack2(ex[0], ex[0]);
            pe.vy0 = pack2(ey[0], ey[0]); pe.vy1 = pack2(ey[kRays - 1], ey[kRays - 1]);
            pe.s0 = pack2(ez[0], ez[0]);  pe.s1 = pack2(ez[kRays - 1], ez[kRays - 1]);
            ro.vx = ex[0];
#pragma unroll
            for (int r = 0; r < kRays; ++r) { ro.vy[r] = ey[r]; ro.sc[r] = ez[r]; }
        }
        if (!CULL) {
            if (PACKET) {
                int g = 0;
#pragma unroll 1
                for (; g + 4 <= n_groups; g += 4) {
                    const int gs[4] = {g, g + 1, g + 2, g + 3};
                    test_packet_groups<kThreads, kRays, 4>(s, fast_base, gs, pe, ro, sphere_obj, n_slots, tid);
                }
#pragma unroll 1
                for (; g < n_groups; ++g) {
                    const int gs[1] = {g};
                    test_packet_groups<kThreads, kRays, 1>(s, fast_base, gs, pe, ro, sphere_obj, n_slots, tid);
                }
            } else {
                uint32_t fa = fast_base;
#pragma unroll 2
                for (int g = 0; g < n_groups; ++g, fa += 48u) test_group<kThreads, AFFINE, kRays>(s, fa, g, ex, ey, ez, sphere_obj, n_slots, tid);
            }
            my_groups += (unsigned int)n_groups;
        } else {
            // Bounding cone of this warp's (active) rays: axis = normalised sum of the directions, cos(theta) = the
            // smallest axis.direction, both widened by the rounding slack.  A group is kept iff the angle between the
            // axes is at most theta + alpha: dot >= cos(theta) cos(alpha) - sin(theta) sin(alpha).  Wide cones
            // (tiny consoles, scattered shadow rays) keep everything.
            float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
            for (int r = 0; r < kRays; ++r) {
                const int slot = r * kThreads + tid;
                ax += s.dirx[slot]; ay += s.diry[slot]; az += s.dirz[slot];      // inactive shadow rays hold 0
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o); ay += __shfl_xor_sync(0xffffffffu, ay, o);
                az += __shfl_xor_sync(0xffffffffu, az, o);
            }
            const float al = sqrtf(ax * ax + ay * ay + az * az);
            const bool axis_ok = al > 1.0e-3f && al < 1.0e30f;
            const float il = axis_ok ? 1.0f / al : 0.0f;
            ax *= il; ay *= il; az *= il;
            float cmin = 1.0f;
#pragma unroll
            for (int r = 0; r < kRays; ++r) {
                const int slot = r * kThreads + tid;
                const float dx = s.dirx[slot], dy = s.diry[slot], dz = s.dirz[slot];
                const float dd = ax * dx + ay * dy + az * dz;
                const bool active = !SHADOW || s.best_t[slot] >= 0.0f;
                cmin = fminf(cmin, active ? dd : 1.0f);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            const float ct = cmin - 4.0e-6f;                                     // wider: unit vectors are unit to ~2e-7
            const bool keep_all = !axis_ok || !(ct > 0.5f);
            const float st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f)) + 4.0e-6f;
            for (int gb = 0; gb < n_groups; gb += 32) {
                const int gi = gb + lane;
                bool keep = false;
                if (gi < n_groups) {
                    const float4 cn = s.gcone[gi];
                    const float thr = ct * cn.w - st * s.gsin[gi] - 4.0e-6f;
                    keep = keep_all ? (cn.w < 2.0f) : (ax * cn.x + ay * cn.y + az * cn.z >= thr);
                }
                uint32_t m = __ballot_sync(0xffffffffu, keep);
                my_groups += (unsigned int)__popc(m);
                while (m) {
                    const int g0 = gb + __ffs(m) - 1;
                    m &= m - 1;
                    if (PACKET) {
                        if (m) {
                            const int gs[2] = {g0, gb + __ffs(m) - 1};
                            m &= m - 1;
                            test_packet_groups<kThreads, kRays, 2>(s, fast_base, gs, pe, ro, sphere_obj, n_slots, tid);
                        } else {
                            const int gs[1] = {g0};
                            test_packet_groups<kThreads, kRays, 1>(s, fast_base, gs, pe, ro, sphere_obj, n_slots, tid);
                        }
                        continue;
                    }
                    test_group<kThreads, AFFINE, kRays>(s, fast_base + 48u * (uint32_t)g0, g0, ex, ey, ez, sphere_obj, n_slots, tid);
                    if (m) {                                     // a second group back to back: the unroll-by-2 of the brute-force loop
                        const int g1 = gb + __ffs(m) - 1;
                        m &= m - 1;
                        test_group<kThreads, AFFINE, kRays>(s, fast_base + 48u * (uint32_t)g1, g1, ex, ey, ez, sphere_obj, n_slots, tid);
                    }
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
                const int slot = r * kThreads + tid;
                const V3 d = v3(s.dirx[slot], s.diry[slot], s.dirz[slot]);
                if (SHADOW && s.best_t[slot] < 0.0f) continue;
                if (plane_trace(pl, o, d, t)) {
                    const float best = s.best_t[slot];
                    if (t < best || (t == best && oi < s.best_idx[slot])) { s.best_t[slot] = t; s.best_idx[slot] = oi; }
                }
            }
        }

        // ---- hit records (primary pass: only when somebody reads them -- the next sphere chunk, the shadow pass, or
        //      rtc_frame_hits; shadow pass: the occlusion byte) ---------------------------------------------------
        if (SHADOW || write_hits) {
#pragma unroll
            for (int r = 0; r < kRays; ++r) {
                const uint32_t row = fp.row0 + ty * kTileH + py * kRays + r;
                if (row < fp.row1 && col < W) {
                    const size_t pix = (size_t)(row - fp.row0) * W + col;
                    const int slot = r * kThreads + tid;
                    if (!SHADOW) {
                        hit_t[pix] = s.best_t[slot];
                        hit_idx[pix] = s.best_idx[slot];
                    } else {
                        const bool occluded = s.best_idx[slot] != -1;
                        if (!carry_in || occluded) shadow[pix] = occluded ? 1 : 0;
                    }
                }
            }
        }

        // ---- tile epilogue: shade + quantise (replaces the separate shade launch and the 8 B/pixel hit-record round
        //      trip; direction, distance and object of every ray are still in shared memory) ------------------------
        if (!SHADOW && shade_mode >= 0) {
            const bool bit8 = shade_mode == RTC_BIT_ASCII || shade_mode == RTC_BIT_PIXEL;
            const bool has_gl = shade_mode == RTC_BIT_ASCII || shade_mode == RTC_RGB_ASCII;
            const uint32_t bpp = bit8 ? 1u : 3u;
            // The tile's planes are staged in the warp's own div2A runs (dead after the sphere loop; one 128-byte run per
            // r holds tile rows r and kRays + r: colour at py * 16 * bpp, glyphs at 96 + py * 16) and leave as 16-byte
            // stores -- which is what a peer GPU's memory wants when the band is written over NVLink.
            const bool fast = (W & 15u) == 0u && ((reinterpret_cast<uintptr_t>(color) | reinterpret_cast<uintptr_t>(glyph)) & 15u) == 0u;
            __syncwarp();
#pragma unroll 1
            for (int r = 0; r < kRays; ++r) {
                const int slot = r * kThreads + tid;
                const float t = s.best_t[slot];
                const int idx = s.best_idx[slot];
                uint32_t v = (bit8 ? 16u : 0u) | ((uint32_t)' ' << 24);
                if (t <= fp.far_dist) v = shade_call(s.shade, s.dirx[slot], s.diry[slot], s.dirz[slot], t, idx);
                const uint32_t row = fp.row0 + ty * kTileH + py * kRays + r;
                if (fast) {
                    unsigned char* run = reinterpret_cast<unsigned char*>(s.div2A + r * kThreads + (tid & ~31));
                    unsigned char* pc = run + (py * 16u + px) * bpp;
                    pc[0] = (unsigned char)v;
                    if (!bit8) { pc[1] = (unsigned char)(v >> 8); pc[2] = (unsigned char)(v >> 16); }
                    if (has_gl) run[96u + py * 16u + px] = (unsigned char)(v >> 24);
                } else if (row < fp.row1 && col < W) {
                    const size_t pix = (size_t)(row - fp.row0) * W + col;
                    color[pix * bpp] = (uint8_t)v;
                    if (!bit8) { color[pix * bpp + 1] = (uint8_t)(v >> 8); color[pix * bpp + 2] = (uint8_t)(v >> 16); }
                    if (has_gl) glyph[pix] = (uint8_t)(v >> 24);
                }
            }
            if (fast) {                                          // W % 16 == 0: every tile is 16 columns wide
                __syncwarp();
                const uint32_t row_t = fp.row0 + ty * kTileH;    // first row of the tile
                const uint32_t parts = bpp;                      // 16-byte pieces per tile row: 3 (RGB) or 1 (index)
                for (uint32_t i = (uint32_t)lane; i < kTileH * parts; i += 32u) {
                    const uint32_t rho = i / parts, part = i - rho * parts;      // tile row, piece
                    const uint32_t row = row_t + rho;
                    if (row < fp.row1) {
                        const uint32_t hy = rho / (uint32_t)kRays, hr = rho - hy * (uint32_t)kRays;   // half of the tile, ray index
                        const unsigned char* run = reinterpret_cast<const unsigned char*>(s.div2A + hr * kThreads + (tid & ~31));
                        const uint4 q = *reinterpret_cast<const uint4*>(run + hy * 16u * bpp + part * 16u);
                        *reinterpret_cast<uint4*>(color + ((size_t)(row - fp.row0) * W + tx * kTile) * bpp + part * 16u) = q;
                    }
                }
                if (has_gl && lane < (int)kTileH) {
                    const uint32_t row = row_t + (uint32_t)lane;
                    if (row < fp.row1) {
                        const uint32_t hy = (uint32_t)lane / (uint32_t)kRays, hr = (uint32_t)lane - hy * (uint32_t)kRays;
                        const unsigned char* run = reinterpret_cast<const unsigned char*>(s.div2A + hr * kThreads + (tid & ~31));
                        const uint4 q = *reinterpret_cast<const uint4*>(run + 96u + hy * 16u);
                        *reinterpret_cast<uint4*>(glyph + (size_t)(row - fp.row0) * W + tx * kTile) = q;
                    }
                }
                __syncwarp();                                    // the runs are div2A again in the next tile
            }
        }
    }
    if (lane == 0 && groups_tested != nullptr && my_groups) atomicAdd(groups_tested, (unsigned long long)my_groups * kRays);   // x 4 spheres x 32 lanes = tests
}

cudaError_t configure_trace()   // per device, once per context
{
    cudaError_t e;
#define RTC_TRACE_ATTR1(SH, T, C, A, R, P)                                                                                   \
    if ((e = cudaFuncSetAttribute(trace_kernel<SH, T, C, A, R, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e
#define RTC_TRACE_ATTR(SH, T, C, A, P) RTC_TRACE_ATTR1(SH, T, C, A, 8, P); RTC_TRACE_ATTR1(SH, T, C, A, 4, P)
    RTC_TRACE_ATTR(false, 768, false, false, false); RTC_TRACE_ATTR(true, 768, false, false, false); RTC_TRACE_ATTR(false, 896, false, false, false); RTC_TRACE_ATTR(true, 896, false, false, false);
    RTC_TRACE_ATTR(false, 768, true, false, false);  RTC_TRACE_ATTR(true, 768, true, false, false);  RTC_TRACE_ATTR(false, 896, true, false, false);  RTC_TRACE_ATTR(true, 896, true, false, false);
    RTC_TRACE_ATTR(false, 768, false, true, false);  RTC_TRACE_ATTR(false, 896, false, true, false); RTC_TRACE_ATTR(false, 768, true, true, false);   RTC_TRACE_ATTR(false, 896, true, true, false);
    RTC_TRACE_ATTR(false, 768, false, true, true);   RTC_TRACE_ATTR(false, 896, false, true, true);  RTC_TRACE_ATTR(false, 768, true, true, true);    RTC_TRACE_ATTR(false, 896, true, true, true);
#undef RTC_TRACE_ATTR
#undef RTC_TRACE_ATTR1
    return cudaSuccess;
}

size_t trace_smem_bytes(int n_slots, int threads, int rays)
{
    return (size_t)n_slots * 28 + (size_t)(((n_slots >> 2) + 3) & ~3) * 24 +     // spheres, per-group bounds (4 + 16 + 4 B),
           (size_t)threads * (6 * rays * 4) + 96;                                 // best_t, best_idx, div2A, dir x/y/z; ShadeCtx
}

// Threads per CTA, rays per thread and sphere slots per launch for a band of `rows` rows.
// A warp's work quantum is one tile (16 x 2*rays pixels), so T tiles take ceil(T / (CTAs * warps)) tile times, and a tile
// time is warps / throughput(warps).  With many waves 24 and 28 warps are equally fast, but an 8-GPU band of a 4K frame
// is 4080 tiles of 16 x 16 = 27.6 per SM: 28 warps finish it in ONE wave (profiles/r01_trace_kernel_ncu.md).
// 16 x 8 tiles (4 rays per thread) halve the quantum but cost 12-14 % more per test (measured, profiles/r02_band_rays.md:
// 271 rows of a 4K frame 0.151 ms with 8 rays, 0.168 with 4; 136 rows 0.133 against 0.095), so they only pay when the
// launch is less than about half a wave of 16 x 16 tiles -- console-sized frames, the reference's own use case.  More warps
// or rays leave less shared memory for spheres, i.e. more launches over a long sphere list -- priced in per chunk.
TracePlan plan_trace(uint32_t x, uint32_t rows, int n_slots, int n_ctas, bool packet)
{
    static const char* force = getenv("RTC_TRACE_THREADS_FORCE");      // experiments only
    static const char* force_rays = getenv("RTC_TRACE_RAYS_FORCE");
    const long long W = (long long)x - 1;
    TracePlan best{896, 0, 8};
    double best_cost = -1.0;
    for (int rays = 8; rays >= 4; rays -= 4) {                         // ties go to 8 rays, then to 28 warps
        if (force_rays && atoi(force_rays) != rays) continue;
        const long long tiles = ((W + kTile - 1) / kTile) * (((long long)rows + 2 * rays - 1) / (2 * rays));
        for (int w = 28; w >= 24; w -= 4) {
            if (force && atoi(force) != w * 32) continue;
            const int max_slots = (int)((227 * 1024 - 192 - (long long)w * 32 * (6 * rays * 4)) / 34) & ~3;
            const int chunks = n_slots <= max_slots ? 1 : (n_slots + max_slots - 1) / max_slots;
            if (chunks > kMaxChunks && w > 24) continue;               // (the API refuses more chunks than it has tickets for)
            const long long per_wave = (long long)n_ctas * w;
            const double waves = (double)((tiles + per_wave - 1) / per_wave);
            // per chunk: the sphere loop over its share of the list + a fixed ray set-up / write-back worth ~40 sphere tests
            // (packet filter: 14 packed ops per group and packet whatever its length, against 50 per 8 rays)
            const double per_test = packet ? (rays == 4 ? 0.62 : 0.30) : (rays == 4 ? 1.14 : 1.0);
            const double tile_time = w * rays * ((double)(n_slots > 0 ? n_slots : 1) * per_test + 40.0 * chunks);
            const double cost = waves * tile_time;
            if (best_cost < 0.0 || cost < best_cost * 0.999) { best_cost = cost; best.threads = w * 32; best.max_slots = max_slots; best.rays = rays; }
        }
    }
    if (best.max_slots == 0) best.max_slots = (int)((227 * 1024 - 192 - (long long)best.threads * (6 * best.rays * 4)) / 34) & ~3;
    return best;
}

// Tickets one launch draws from its counter: one per tile plus one per warp (the draw that tells a warp it is done).
unsigned long long trace_tickets(uint32_t x, uint32_t rows, int n_ctas, int threads, int rays)
{
    const unsigned long long W = x - 1u;
    return ((W + kTile - 1) / kTile) * (((unsigned long long)rows + 2 * rays - 1) / (2 * rays)) + (unsigned long long)n_ctas * (threads / 32);
}

// Deflation of c for the packet filter (see test_packet_groups): E = 1e-5 + 1.5 D^2, D = the vy span of one packet.
// vy = cy e2, cy = (y - 2 row) / y (RayTracing.cu:16,20): consecutive rows are 2 |e2| / y apart (+ the rounding of cy and vy).
static float packet_eps(const FrameParams& fp, int rays)
{
    const double e2 = fabs((double)fp.e2);
    const double D = (double)(rays - 1) * (2.0 * e2 / (double)fp.fy) * 1.0001 + 1.0e-6 * (1.0 + e2);
    const double E = 1.0e-5 + 1.5 * D * D;
    return (float)(E < 4.0 ? E : 4.0);                           // (E >= kappa: the sphere is an unconditional candidate)
}

cudaError_t launch_trace(cudaStream_t st, int n_ctas, const FrameParams& fp, const int32_t* sphere_obj, int n_spheres,
                         int n_slots, const rtc_object* objs, const int32_t* plane_obj, int n_planes, float* hit_t,
                         int32_t* hit_idx, unsigned long long* tile_counter, unsigned long long ticket_base, int carry_in,
                         const float* light, uint8_t* shadow, int threads, bool cull, unsigned long long* groups_tested,
                         unsigned long long* stats_zero, const ShadeParams& sp, int shade_mode,
                         uint8_t* color, uint8_t* glyph, bool write_hits, const float4* obj_kd, bool affine, int rays, bool packet)
{
    const size_t smem = trace_smem_bytes(n_slots, threads, rays);
    const float l0 = light ? light[0] : 0.f, l1 = light ? light[1] : 0.f, l2 = light ? light[2] : 0.f;
    if (light && affine) return cudaErrorInvalidValue;           // shadow rays do not come from a pixel grid
    if (packet && !affine) return cudaErrorInvalidValue;         // the packet filter is a form of the screen-affine one
    if (rays != 8 && rays != 4) return cudaErrorInvalidValue;
    HoistBasis hb;
    hb.affine = affine ? 1 : 0;
    hb.eps = packet ? packet_eps(fp, rays) : RTC_FILTER_EPS;
    for (int i = 0; i < 3; ++i) { hb.c0[i] = fp.m[4 * i + 0]; hb.c1[i] = fp.m[4 * i + 1]; hb.c2[i] = fp.m[4 * i + 2]; }
#define RTC_TRACE_LAUNCH1(SH, T, C, A, R, P)                                                                             \
    trace_kernel<SH, T, C, A, R, P><<<n_ctas, T, smem, st>>>(fp, hb, sphere_obj, n_spheres,                              \
                                                             n_slots, objs, plane_obj, n_planes, hit_t, hit_idx, tile_counter, \
                                                             ticket_base, carry_in, l0, l1, l2, shadow, groups_tested, stats_zero, \
                                                             sp, shade_mode, color, glyph, write_hits ? 1 : 0, obj_kd)
#define RTC_TRACE_LAUNCH(SH, T, C, A, P)                                                                                 \
    do { if (rays == 8) RTC_TRACE_LAUNCH1(SH, T, C, A, 8, P); else RTC_TRACE_LAUNCH1(SH, T, C, A, 4, P); } while (0)
#define RTC_TRACE_PICK(T)                                                                                                \
    do {                                                                                                                 \
        if (light) { if (cull) RTC_TRACE_LAUNCH(true, T, true, false, false); else RTC_TRACE_LAUNCH(true, T, false, false, false); } \
        else if (packet) { if (cull) RTC_TRACE_LAUNCH(false, T, true, true, true); else RTC_TRACE_LAUNCH(false, T, false, true, true); } \
        else if (affine) { if (cull) RTC_TRACE_LAUNCH(false, T, true, true, false); else RTC_TRACE_LAUNCH(false, T, false, true, false); } \
        else       { if (cull) RTC_TRACE_LAUNCH(false, T, true, false, false); else RTC_TRACE_LAUNCH(false, T, false, false, false); } \
    } while (0)
    if (threads == 896) RTC_TRACE_PICK(896);
    else if (threads == 768) RTC_TRACE_PICK(768);
    else return cudaErrorInvalidValue;
#undef RTC_TRACE_PICK
#undef RTC_TRACE_LAUNCH
#undef RTC_TRACE_LAUNCH1
    return cudaGetLastError();
}

}  // namespace rtc
