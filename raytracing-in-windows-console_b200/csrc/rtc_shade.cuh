// rtc_shade.cuh -- shade + quantise ONE hit pixel (device code shared by the ray kernel's tile epilogue, rtc_trace.cu, and
// the stand-alone shade kernel used after a shadow pass, rtc_shade.cu).
//
// Replaces, per traced pixel, the tail of the reference's RayTrace (normal, Blinn-Phong, RayTracing.cu:123-157 + :41-79),
// the (uint8_t) truncation and xterm-256 quantisation in the RayTrace_* kernels (RayTracing.cu:210, :527-567, :669-709;
// ANSIRGB.h:141-189) and GetASCIICharacter (RayTracing.cu:26-39).
//
// All arithmetic that can move an output byte is "exact" (rtc_device.cuh): same operations, same order, one rounding
// each, as the reference source.
#pragma once
#include "rtc_device.cuh"

namespace rtc {

// Light and material constants of the reference's call site (RayTracing.cu:143-152, :69, :77) as a kernel parameter
// (rtc_set_light).  The defaults ARE the reference's constants; with them every operation below sees the same operands
// as the reference's literals.  Shininess stays 32 (pow32 below is exact only for that exponent).
struct ShadeParams {
    float light[3];       // (1, 50, 0)                                   :146
    float diff_color;     // 1      diffuseColor                          :147
    float diff_power;     // 2000   diffusePower                          :147
    float spec_color;     // 1      specColor                             :148
    float spec_power;     // 3000   specPower                             :148
    float ambient[3];     // (0.2, 0.2, 0.2)                              :77
    float obj_specular;   // 1      the object's specular colour          :78
};

// ---- xterm-256 tables, restated from the palette definition (not copied from ANSIRGB.h) ----
struct GreyLut { uint8_t v[256]; };
constexpr int kCubeLevel[6] = {0, 95, 135, 175, 215, 255};
constexpr GreyLut make_grey_lut()
{
    // Nearest of the 30 greys the palette offers: the 24-step ramp 232..255 (8+10k) and the six
    // cube greys 16,59,102,145,188,231.  Exact ties go to the darker entry below 120 and to the
    // brighter one above (the reference table's behaviour; exhaustively tested).
    GreyLut t{};
    for (int v = 0; v < 256; ++v) {
        int best = 0, bestd = 1 << 30, bestval = 0;
        for (int j = 0; j < 30; ++j) {
            const int idx = j < 6 ? 16 + 43 * j : 232 + (j - 6);
            const int val = j < 6 ? kCubeLevel[j] : 8 + 10 * (j - 6);
            const int dd = val > v ? val - v : v - val;
            const bool tie_wins = (v < 120) ? (val < bestval) : (val > bestval);
            if (dd < bestd || (dd == bestd && tie_wins)) { bestd = dd; best = idx; bestval = val; }
        }
        t.v[v] = (uint8_t)best;
    }
    return t;
}
static __device__ const GreyLut d_grey_lut = make_grey_lut();

__device__ __forceinline__ uint32_t pal_rgb(uint32_t idx)   // xterm palette entry for idx >= 16
{
    if (idx >= 232u) { const uint32_t v = 8u + 10u * (idx - 232u); return (v << 16) | (v << 8) | v; }
    const uint32_t k = idx - 16u, r = k / 36u, g = (k / 6u) % 6u, b = k % 6u;
    // cube level l -> 0, 95, 135, 175, 215, 255
    const uint32_t lr = r ? 55u + 40u * r : 0u, lg = g ? 55u + 40u * g : 0u, lb = b ? 55u + 40u * b : 0u;
    return (lr << 16) | (lg << 8) | lb;
}
__device__ __forceinline__ uint32_t pal_distance(uint32_t x, uint32_t y)   // ANSIRGB.h:118-124
{
    const int rs = (int)((x >> 16) & 255u) + (int)((y >> 16) & 255u);
    const int r = (int)((x >> 16) & 255u) - (int)((y >> 16) & 255u);
    const int g = (int)((x >> 8) & 255u) - (int)((y >> 8) & 255u);
    const int b = (int)(x & 255u) - (int)(y & 255u);
    return (uint32_t)((1024 + rs) * r * r + 2048 * g * g + (1534 - rs) * b * b);
}
__device__ __forceinline__ uint32_t cube_level(uint32_t v, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, uint32_t t4)
{
    return (v >= t0) + (v >= t1) + (v >= t2) + (v >= t3) + (v >= t4);
}
__device__ __forceinline__ uint32_t ansi256_from_rgb(uint32_t r, uint32_t g, uint32_t b)   // ANSIRGB.h:141-189
{
    if (r == g && g == b) return d_grey_lut.v[b];
    const uint32_t rgb = (r << 16) | (g << 8) | b;
    const uint32_t lum = (3567664u * r + 11998547u * g + 1211005u * b + (1u << 23)) >> 24;   // :126-134
    const uint32_t grey_index = d_grey_lut.v[lum & 255u];
    const uint32_t grey_distance = pal_distance(rgb, pal_rgb(grey_index));
    const uint32_t ir = cube_level(r, 38, 115, 155, 196, 235);    // :18-20
    const uint32_t ig = cube_level(g, 36, 116, 154, 195, 235);    // :25-27
    const uint32_t ib = cube_level(b, 35, 115, 155, 195, 235);    // :32-34
    const uint32_t cube_idx = 16u + 36u * ir + 6u * ig + ib;
    return pal_distance(rgb, pal_rgb(cube_idx)) < grey_distance ? cube_idx : grey_index;   // :188
}

// 68-step ASCII ramp (RayTracing.h:97-115).
static __device__ const char d_ascii_ramp[69] = " .`^\",:;Il!i><~+_-?*][}{1)(|/tfjrxnuvczmwXYUJCLqpdbkhao#%ZO8B$0QM&W@";

// x^32 for x in [0,1] (RayTracing.cu:73 `pow(x, 32.0f)`): five squarings in binary64, rounded
// once to binary32 -- the correctly rounded result, which is what the host libm's powf returns
// in all but vanishingly rare double-rounding cases (DESIGN.md "pow").
__device__ __forceinline__ float pow32(float x)
{
    if (!(x == x)) return x;
    double v = (double)x;
    v = __dmul_rn(v, v); v = __dmul_rn(v, v); v = __dmul_rn(v, v); v = __dmul_rn(v, v); v = __dmul_rn(v, v);
    return __double2float_rn(v);
}

// BlinnPhongShading (RayTracing.cu:41-79) with the constants of its call site (:143-152).
__device__ __forceinline__ V3 blinn_phong(const ShadeParams& sp, V3 kd, V3 point, V3 view, V3 normal)
{
    V3 L = vsub(v3(sp.light[0], sp.light[1], sp.light[2]), point);   // :48, light position :146
    const float len = vlength(L);                               // :50
    const float dist = mul(len, len);                           // :51
    const float inv = rcp(dist);                                // :52
    const float linv = rcp(len);                                // :54  Normalize_GPU recomputes the same length
    L = v3(mul(L.x, linv), mul(L.y, linv), mul(L.z, linv));
    const V3 N = vnormalize_unit(normal);                       // :56  (already normalised twice by the caller)
    const V3 V = vnormalize_unit(view);                         // :57  (already normalised by the caller, :150)
    const float di = clampf_ref(vdot(N, L), 0.0f, 1.0f);        // :60-61
    const float diff = mul(mul(mul(sp.diff_color, di), sp.diff_power), inv);   // :64
    const V3 H = vnormalize(vadd(L, V));                        // :67
    const float si = pow32(clampf_ref(vdot(N, H), 0.0f, 1.0f)); // :72-73
    const float spec = mul(mul(mul(sp.spec_color, si), sp.spec_power), inv);   // :75
    // :77-78  ambient*kd + diffuse*kd + specular*objectSpecular
    return v3(add(add(mul(sp.ambient[0], kd.x), mul(diff, kd.x)), mul(spec, sp.obj_specular)),
              add(add(mul(sp.ambient[1], kd.y), mul(diff, kd.y)), mul(spec, sp.obj_specular)),
              add(add(mul(sp.ambient[2], kd.z), mul(diff, kd.z)), mul(spec, sp.obj_specular)));
}

constexpr int kModeNormalsSaturate = 100 + RTC_RGB_NORMALS;   // `mode` of shade_pixel for RGB_NORMALS + RTC_FLAG_NORMALS_SATURATE

// One traced pixel -> colour key (RGB modes: R | G << 8 | B << 16; 8-bit modes: xterm-256 index) | glyph << 24.
//   d: the ray direction (CalculateInitialDirection), t / idx: the accepted hit (idx < 0: none), shadowed: the shadow-ray
//   extension found an occluder (ambient term only).  `objs` is the 64-byte object array, 16-byte aligned; obj_kd[i] is
//   object i's colour / 255 (computed by the host at scene upload, rtc_api.cu: upload_scene).
//   (BIT8 / GLYPH follow from `mode`; they are separate arguments so that callers with compile-time modes fold them.)
__device__ __forceinline__ uint32_t shade_pixel(const bool BIT8, const bool GLYPH, int mode, const ShadeParams& sp,
                                                const rtc_object* __restrict__ objs, const float4* __restrict__ obj_kd, V3 cam,
                                                float far_dist, V3 d, float t, int idx, bool shadowed)
{
    uint32_t c0 = BIT8 ? 16u : 0u, c1 = 0u, c2 = 0u, gl = ' ';       // miss cell: ESC[48;5;<NUL>16m / ESC[48;2;0;0;0m (RayTracing.cu:248, :599)
    const bool hit = t <= far_dist;                                   // RayTracing.cu:508
    if (hit && idx >= 0) {
        const float4* q = reinterpret_cast<const float4*>(objs + idx);
        const float4 q0 = __ldg(q);                                   // type, centre
        V3 n;
        if (__float_as_int(q0.x) == RTC_OBJ_SPHERE) {                 // Sphere.cu:67
            n = vnormalize(vsub(vadd(cam, vscale(d, t)), v3(q0.y, q0.z, q0.w)));
        } else {
            const float4 q2 = __ldg(q + 2);                           // Plane.cu:72
            n = v3(q2.x, q2.y, q2.z);
        }
        n = vnormalize_unit(n);                                       // RayTracing.cu:129
        const float shading_value = add(add(mul(n.x, 1.0f), mul(n.y, 0.0f)), mul(n.z, 0.0f));   // :133
        if (GLYPH) {                                                  // GetASCIICharacter, RayTracing.cu:26-39
            int gi = (int)ceilf(mul(shading_value, 67.0f));
            gi = gi < 1 ? 1 : gi;
            gi = gi > 67 ? 67 : gi;                                   // index 68 (one past the table) pinned to 67
            gl = (uint32_t)(unsigned char)d_ascii_ramp[gi];
        }
        uint32_t r8, g8, b8;
        if (mode == RTC_RGB_NORMALS) {                                // RayTracing.cu:669-709
            r8 = to_u8(mul(n.x, 255.0f)); g8 = to_u8(mul(n.y, 255.0f)); b8 = to_u8(mul(n.z, 255.0f));
        } else if (mode == kModeNormalsSaturate) {                    // RTC_FLAG_NORMALS_SATURATE: the CUDA platform's cast
            r8 = to_u8_sat(mul(n.x, 255.0f)); g8 = to_u8_sat(mul(n.y, 255.0f)); b8 = to_u8_sat(mul(n.z, 255.0f));
        } else {
            const float4 k4 = __ldg(obj_kd + idx);                                         // :144 colour / 255, hoisted per object
            const V3 kd = v3(k4.x, k4.y, k4.z);
            const V3 point = vadd(cam, vscale(d, t));                                      // :149
            V3 sh;
            if (shadowed)                                             // extension: occluded -> ambient only
                sh = vcmul(v3(sp.ambient[0], sp.ambient[1], sp.ambient[2]), kd);
            else
                sh = blinn_phong(sp, kd, point, vnormalize_unit(vscale(d, -1.0f)), n);     // :143-152 (d is a unit vector: RayTracing.cu:23)
            sh = vscale(sh, 255.0f);                                                       // :154
            r8 = to_u8(minf_ref(255.0f, sh.x));                                            // :157, :527
            g8 = to_u8(minf_ref(255.0f, sh.y));
            b8 = to_u8(minf_ref(255.0f, sh.z));
        }
        if (BIT8) c0 = ansi256_from_rgb(r8, g8, b8);                  // :210
        else { c0 = r8; c1 = g8; c2 = b8; }
    } else if (hit && GLYPH) {
        gl = '.';   // far plane >= 99999999: a miss prints as a black "hit" with shadingValue 0 -> ASCII[1]
    }
    return c0 | (c1 << 8) | (c2 << 16) | (gl << 24);
}

}  // namespace rtc
