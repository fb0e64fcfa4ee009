// rtc_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Two kinds of arithmetic live here:
//   * "exact": IEEE binary32, one rounding per operation, evaluated in the order the
//     reference writes it (no FMA contraction) -- built from __fmul_rn/__fadd_rn/... so the
//     result does not depend on compiler flags.  Used wherever a rounding decides a byte of
//     the output (ray generation, accepted hits, shading, quantisation).
//   * "fast": fused / packed (FFMA2) arithmetic used only as a conservative filter in the
//     ray kernel's inner loop; survivors are re-evaluated with the exact path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtc.h"

namespace rtc {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqt(float a) { return __fsqrt_rn(a); }
// 1.0f / a.  Both __frcp_rn and __fdiv_rn(1.0f, a) are correctly rounded, hence identical; the reciprocal is the
// shorter sequence (MUFU.RCP + 3 FMA-pipe ops against 5 + FCHK).
__device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }

// MyMath (reference MyMath.h:60-106, MyMath.cu:4-34), left-to-right, un-fused.
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return v3(sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)); }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return v3(add(a.x, b.x), add(a.y, b.y), add(a.z, b.z)); }
__device__ __forceinline__ V3 vscale(V3 a, float s) { return v3(mul(a.x, s), mul(a.y, s), mul(a.z, s)); }
__device__ __forceinline__ V3 vdiv(V3 a, float s) { return v3(dvd(a.x, s), dvd(a.y, s), dvd(a.z, s)); }
__device__ __forceinline__ V3 vcmul(V3 a, V3 b) { return v3(mul(a.x, b.x), mul(a.y, b.y), mul(a.z, b.z)); }
__device__ __forceinline__ float vdot(V3 a, V3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
// Vector3::Normalize_GPU (MyMath.h:139-146): reciprocal length, no zero check.
__device__ __forceinline__ V3 vnormalize(V3 a)
{
    const float inv = rcp(sqt(vdot(a, a)));
    return v3(mul(a.x, inv), mul(a.y, inv), mul(a.z, inv));
}
// fl(1 / fl(sqrt(s))) for s within 128 ulp of 1, by integer arithmetic on the bit pattern -- EXACTLY the value the two
// correctly rounded operations give (checked for every such s in tests/test_oracle_golden.py::test_unit_rsqrt_formula):
//   s = 1 + m 2^-23 (m >= 0): sqrt = 1 + (m/2) 2^-23 - tiny -> rounds to 1 + floor(m/2) 2^-23 =: 1 + j 2^-23 (an exact tie
//       minus tiny rounds down); 1/(1 + j 2^-23) = 1 - 2j 2^-24 + tiny -> rounds to 1 - 2j 2^-24;
//   s = 1 - k 2^-24 (k > 0):  sqrt = 1 - (k/2) 2^-24 - tiny -> rounds to 1 - ceil(k/2) 2^-24 =: 1 - j 2^-24;
//       1/(1 - j 2^-24) = 1 + (j/2) 2^-23 + tiny -> rounds to 1 + ceil(j/2) 2^-23.
// The reference re-normalises vectors that are already normalised (the sphere normal three times, the view vector
// twice: Sphere.cu:67, RayTracing.cu:129, :56, :150, :57); each pass may move an ulp, so none can be dropped -- but its
// square root and reciprocal (~17 FMA-pipe instructions) reduce to this.
__device__ __forceinline__ float rsqrt_near_one_exact(float s)
{
    const int m = __float_as_int(s) - 0x3F800000;
    if (m >= -128 && m <= 128) {
        const int k = -m;
        const int bits = m >= 0 ? 0x3F800000 - (m & ~1) : 0x3F800000 + ((((k + 1) >> 1) + 1) >> 1);
        return __int_as_float(bits);
    }
    return rcp(sqt(s));                                          // not a unit vector (degenerate geometry): the general sequence
}
// Vector3::Normalize_GPU of a vector that is already of unit length to a few ulp (falls back to the general sequence
// when it is not): bit-identical to vnormalize.
__device__ __forceinline__ V3 vnormalize_unit(V3 a)
{
    const float inv = rsqrt_near_one_exact(vdot(a, a));
    return v3(mul(a.x, inv), mul(a.y, inv), mul(a.z, inv));
}
__device__ __forceinline__ float vlength(V3 a) { return sqt(vdot(a, a)); }
__device__ __forceinline__ float clampf_ref(float v, float lo, float hi)   // MyMath.cu:29-34 (NaN passes through)
{
    const float r = v < lo ? lo : v;
    return r > hi ? hi : r;
}
__device__ __forceinline__ float minf_ref(float a, float b) { return a < b ? a : b; }   // MyMath.cu:59-62

// Per-frame uniforms, passed by value as a kernel parameter (lands in the constant bank,
// so every use is a c[0][..] operand and costs no load).  == rtc_params + band info.
struct FrameParams {
    float m[12];        // rows 1..3 of inverseVMatrix (row 4 is unused: RayTracing.cu:22 takes .xyz())
    float cam[3];
    float e1, e2, far_dist;
    float fx, fy;       // (float)x, (float)y
    uint32_t x, y;      // console size; traced width = x-1
    uint32_t row0, row1;// band of rows this launch traces
};

// CalculateInitialDirection (reference RayTracing.cu:9-24), exact.  Also hands out the view-space coordinates
// vx = cx*e1, vy = cy*e2 (:20) and the reciprocal length `inv` of the un-normalised direction w = M (vx, vy, 1, 0) (:22-23):
// the ray kernel's packed filter works on w = col2 + vx col0 + vy col1 directly (rtc_trace.cu, "screen-affine filter").
__device__ __forceinline__ V3 initial_direction_ex(const FrameParams& p, uint32_t row, uint32_t col, float& vx, float& vy, float& inv)
{
    const float cy = dvd(sub(p.fy, (float)(2u * row)), p.fy);          // :16
    const float cx = dvd(sub((float)(2u * col), p.fx), p.fx);          // :17
    vx = mul(cx, p.e1); vy = mul(cy, p.e2);                            // :20  (vz = 1, vw = 0)
    V3 w;                                                              // Matrix::Mult, MyMath.h:303-311
    w.x = add(add(add(mul(p.m[0], vx), mul(p.m[1], vy)), mul(p.m[2], 1.0f)), mul(p.m[3], 0.0f));
    w.y = add(add(add(mul(p.m[4], vx), mul(p.m[5], vy)), mul(p.m[6], 1.0f)), mul(p.m[7], 0.0f));
    w.z = add(add(add(mul(p.m[8], vx), mul(p.m[9], vy)), mul(p.m[10], 1.0f)), mul(p.m[11], 0.0f));
    inv = rcp(sqt(vdot(w, w)));                                        // :23  Normalize_GPU, MyMath.h:139-146
    return v3(mul(w.x, inv), mul(w.y, inv), mul(w.z, inv));
}
__device__ __forceinline__ V3 initial_direction(const FrameParams& p, uint32_t row, uint32_t col)
{
    float vx, vy, inv;
    return initial_direction_ex(p, row, col, vx, vy, inv);
}

// Plane::Trace (reference Plane.cu:38-73), exact.  Returns true on hit and sets t.
__device__ __forceinline__ bool plane_trace(const rtc_object& pl, V3 o, V3 d, float& t)
{
    const V3 n = v3(pl.normal[0], pl.normal[1], pl.normal[2]);
    const V3 pos = v3(pl.center[0], pl.center[1], pl.center[2]);
    const float dn = vdot(d, n);                                              // :43
    if (dn > 0.0f || fabsf(sub(dn, 0.0f)) < 1.1920928955078125e-7f) return false;   // :47 (FloatEquals, MyMath.cu:43-47)
    const float t1 = dvd(vdot(vsub(pos, o), n), dn);                          // :52
    if (t1 <= 0.0f) return false;                                             // :54
    const V3 h = vadd(o, vscale(d, t1));                                      // :59
    const float hw = mul(pl.width, 0.5f), hh = mul(pl.height, 0.5f);          // :60-61
    if ((h.x <= sub(pos.x, hw) || h.x >= add(pos.x, hw)) || (h.z <= sub(pos.z, hh) || h.z >= add(pos.z, hh))) // :64-65
        return false;
    t = t1;
    return true;
}

// Sphere::Trace (reference Sphere.cu:30-68) from the hoisted per-sphere terms
// oc = origin - centre and c = oc.oc - r*r (both bit-identical to what the reference
// recomputes per ray, since every primary ray shares the origin).  Exact.
__device__ __forceinline__ bool sphere_trace_hoisted(float ocx, float ocy, float ocz, float c, V3 d,
                                                     float fourA, float divTwoA, float& t)
{
    const float b = mul(2.0f, vdot(d, v3(ocx, ocy, ocz)));           // :36
    const float disc = sub(mul(b, b), mul(fourA, c));                // :39
    if (disc < 0.0f) return false;                                   // :42
    const float sq = sqt(disc);
    const float mb = -b;
    const float t1 = mul(add(mb, sq), divTwoA);                      // :52
    const float t2 = mul(sub(mb, sq), divTwoA);                      // :53
    if (t1 < 0.0f || t2 < 0.0f) return false;                        // :57
    t = minf_ref(t1, t2);                                            // :63
    return true;
}

// ---- packed FP32x2 helpers (sm_100a FFMA2 / FMUL2; one issue slot, two FMAs per lane) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// 3-input max (FMNMX3 on sm_100a): folds two discriminants into the running maximum.
__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// float -> uint8_t the way the reference's host compiler does it (cvttss2si, low byte):
// [0,255] truncates, NaN -> 0, negatives wrap (only RGB_NORMALS sees those; RayTracing.cu:669).
__device__ __forceinline__ uint32_t to_u8(float f)
{
    if (!(f == f)) return 0u;
    return (uint32_t)__float2int_rz(f) & 0xffu;
}

// The same cast on the reference's CUDA platform: cvt.rzi.u8.f32 saturates (negative -> 0, > 255 -> 255, NaN -> 0).
__device__ __forceinline__ uint32_t to_u8_sat(float f)
{
    if (!(f == f)) return 0u;
    return (uint32_t)__float2int_rz(fminf(fmaxf(f, 0.0f), 255.0f));
}

}  // namespace rtc
