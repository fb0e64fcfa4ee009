/*
 * rt_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's per-frame hot path (everything inside
 * RayTracingManager::Update, reference RayTracingManager.cu:76-154).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it, and only as the checker / the reported CPU baseline -- never as the product path.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so
 * this file is pinned against the reference's own sources compiled unmodified for CPU
 * (oracle/ref_build -> oracle/_ref/libref_cpu.so): tests/test_oracle_vs_reference.py
 * (runs where /root/reference exists) and the committed fixtures in tests/golden/
 * generated from that build by tests/golden/make_golden.py.
 *
 * Arithmetic: IEEE binary32, evaluated strictly left to right as the reference writes it;
 * compile with -ffp-contract=off and without -ffast-math (SURVEY 8c "canonical oracle").
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/ConsoleProject/).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/rtc.h" /* POD layouts + mode/flag enums only */

#define ORC_API __attribute__((visibility("default")))

typedef struct { float x, y, z; } v3;

/* ---- MyMath (MyMath.h:6-168, MyMath.cu:4-67) -------------------------------------------- */
static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }   /* MyMath.h:60 */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }   /* MyMath.h:74 */
static inline v3 v3_scale(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }    /* MyMath.h:89 */
static inline v3 v3_div(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }      /* MyMath.h:103 */
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }       /* MyMath.cu:4-8 */
static inline v3 v3_cmul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }  /* MyMath.cu:22-26 */
/* Normalize_GPU: reciprocal length, no zero check (MyMath.h:139-146). */
static inline v3 v3_normalize_gpu(v3 a)
{
    const float inv = 1.0f / sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3_make(a.x * inv, a.y * inv, a.z * inv);
}
/* Normalize: zero-checked host version used by the Plane ctor (MyMath.h:117-123). */
static inline v3 v3_normalize_host(v3 a)
{
    const float len = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    const float div = len < 0.000001f ? 0.0f : 1.0f / len;
    return v3_make(a.x * div, a.y * div, a.z * div);
}
static inline float v3_length(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }   /* MyMath.h:160-163 */
static inline float clampf(float v, float lo, float hi)                                    /* MyMath.cu:29-34 */
{
    const float r = v < lo ? lo : v;
    return r > hi ? hi : r;
}
static inline int clampi(int v, int lo, int hi) { const int r = v < lo ? lo : v; return r > hi ? hi : r; } /* MyMath.cu:36-41 */
static inline float minf_ref(float a, float b) { return a < b ? a : b; }                   /* MyMath.cu:59-62 */
static inline int float_equals(float a, float b) { return fabsf(a - b) < 1.1920928955078125e-7f; } /* MyMath.cu:43-47, FLT_EPSILON */

/* ---- ray generation: CalculateInitialDirection (RayTracing.cu:9-24) ---------------------- */
static v3 initial_direction(const rtc_params* p, uint32_t row, uint32_t col)
{
    /* size_t row*2 / 2*column are formed in integers, then converted (RayTracing.cu:16-17). */
    const float cy = ((float)p->y - (float)((uint64_t)row * 2u)) / (float)p->y;
    const float cx = ((float)(2u * (uint64_t)col) - (float)p->x) / (float)p->x;
    const float vx = cx * p->element1, vy = cy * p->element2, vz = 1.0f, vw = 0.0f;       /* :20 */
    const float* m = p->inv_view;                                                          /* Matrix::Mult, MyMath.h:303-311 */
    v3 w;
    w.x = m[0] * vx + m[1] * vy + m[2] * vz + m[3] * vw;
    w.y = m[4] * vx + m[5] * vy + m[6] * vz + m[7] * vw;
    w.z = m[8] * vx + m[9] * vy + m[10] * vz + m[11] * vw;
    return v3_normalize_gpu(w);                                                            /* :23 */
}

/* ---- Sphere::Trace (Sphere.cu:30-68) ------------------------------------------------------ */
typedef struct { float a, fourA, divTwoA; } ray_terms;   /* ObjectTraceInputData, Object3D.h:16-25 */

static int sphere_trace(const rtc_object* s, v3 o, v3 d, const ray_terms* rt, float* dist, v3* nrm)
{
    const v3 pos = v3_make(s->center[0], s->center[1], s->center[2]);
    const v3 oc = v3_sub(o, pos);                               /* :34 */
    const float b = 2.0f * v3_dot(d, oc);                       /* :36 */
    const float c = v3_dot(oc, oc) - (s->radius * s->radius);   /* :37 */
    const float disc = b * b - rt->fourA * c;                   /* :39 */
    if (disc < 0.0f) return 0;                                  /* :42 */
    const float sq = sqrtf(disc);
    const float mb = -b;
    float t1 = (mb + sq) * rt->divTwoA;                         /* :52 */
    const float t2 = (mb - sq) * rt->divTwoA;                   /* :53 */
    if (t1 < 0.0f || t2 < 0.0f) return 0;                       /* :57 camera inside / sphere behind */
    t1 = minf_ref(t1, t2);                                      /* :63 */
    *dist = t1;
    *nrm = v3_normalize_gpu(v3_sub(v3_add(o, v3_scale(d, t1)), pos)); /* :67 */
    return 1;
}

/* ---- Plane::Trace (Plane.cu:38-73) -------------------------------------------------------- */
static int plane_trace(const rtc_object* pl, v3 o, v3 d, float* dist, v3* nrm)
{
    const v3 n = v3_make(pl->normal[0], pl->normal[1], pl->normal[2]);
    const v3 pos = v3_make(pl->center[0], pl->center[1], pl->center[2]);
    const float dn = v3_dot(d, n);                                            /* :43 */
    if (dn > 0.0f || float_equals(dn, 0.0f)) return 0;                        /* :47 back side / parallel */
    const float t1 = v3_dot(v3_sub(pos, o), n) / dn;                          /* :52 */
    if (t1 <= 0.0f) return 0;                                                 /* :54 */
    const v3 h = v3_add(o, v3_scale(d, t1));                                  /* :59 */
    const float hw = pl->width * 0.5f, hh = pl->height * 0.5f;                /* :60-61 */
    if ((h.x <= pos.x - hw || h.x >= pos.x + hw) || (h.z <= pos.z - hh || h.z >= pos.z + hh)) /* :64-65 */
        return 0;
    *dist = t1;
    *nrm = n;
    return 1;
}

/* ---- BlinnPhongShading (RayTracing.cu:41-79) with its call-site constants (:143-152) ------- */
/* The light / material constants of the call site (RayTracing.cu:143-152, :77).  The reference hard-codes them; the
 * product exposes them (rtc_set_light), so the checker can be told the same values (orc_set_light; NULL = reference). */
static rtc_light g_light = {{1.0f, 50.0f, 0.0f}, 1.0f, 2000.0f, 1.0f, 3000.0f, {0.2f, 0.2f, 0.2f}, 1.0f};

static v3 blinn_phong(v3 kd, v3 point, v3 view, v3 normal)
{
    const v3 light_pos = v3_make(g_light.pos[0], g_light.pos[1], g_light.pos[2]);   /* :146 */
    const float diff_power = g_light.diffuse_power, spec_power = g_light.specular_power;   /* :147-148 */
    const v3 ones = v3_make(1.0f, 1.0f, 1.0f);
    v3 L = v3_sub(light_pos, point);                            /* :48 */
    float dist = v3_length(L);                                  /* :50 */
    dist = dist * dist;                                         /* :51 */
    const float inv = 1.0f / dist;                              /* :52 */
    L = v3_normalize_gpu(L);                                    /* :54 */
    const v3 N = v3_normalize_gpu(normal);                      /* :56 */
    const v3 V = v3_normalize_gpu(view);                        /* :57 */
    const float ndl = v3_dot(N, L);                             /* :60 */
    const float di = clampf(ndl, 0.0f, 1.0f);                   /* :61 */
    const v3 diffuse = v3_scale(v3_scale(v3_scale(v3_scale(ones, g_light.diffuse_color), di), diff_power), inv); /* :64 (diffuseColor = 1) */
    const v3 H = v3_normalize_gpu(v3_add(L, V));                /* :67 */
    const float ndh = v3_dot(N, H);                             /* :72 */
    const float si = powf(clampf(ndh, 0.0f, 1.0f), 32.0f);      /* :73 */
    const v3 specular = v3_scale(v3_scale(v3_scale(v3_scale(ones, g_light.specular_color), si), spec_power), inv); /* :75 (specColor = 1) */
    const v3 amb = v3_make(g_light.ambient[0], g_light.ambient[1], g_light.ambient[2]);                   /* :77 */
    return v3_add(v3_add(v3_cmul(amb, kd), v3_cmul(diffuse, kd)), v3_cmul(specular, v3_scale(ones, g_light.object_specular))); /* :78 */
}

/* ---- RayTrace (RayTracing.cu:81-168) -------------------------------------------------------- */
typedef struct {
    v3 color;            /* RayTraceReturnData, RayTracing.h:17-23 */
    v3 normal;
    float distance;
    float shading_value;
    int32_t index;       /* accepted object index or -1 (not in the reference; parity hook) */
    int32_t lit;         /* shadow extension: 1 unless the shadow ray was blocked            */
} hit_rec;

/* Shadow-ray EXTENSION (not in the reference, which casts no shadow rays: SURVEY F1), defined with the
 * reference's own Trace functions.  The ray is cast FROM THE LIGHT (1,50,0, RayTracing.cu:146) toward
 * the shaded point lifted off the surface, P' = P + n*1e-3: the point is in shadow iff some object is hit at a
 * distance strictly below |P' - light|.  (Casting from the light gives every shadow ray of a frame the same origin,
 * which is what lets the GPU hoist the per-sphere terms exactly as for primary rays.)                          */
static int shadow_blocked(const rtc_object* objs, uint32_t n, v3 point, v3 normal)
{
    const v3 light_pos = v3_make(g_light.pos[0], g_light.pos[1], g_light.pos[2]);
    const v3 lp = v3_sub(v3_add(point, v3_scale(normal, 1.0e-3f)), light_pos);
    const float len = v3_length(lp);
    const v3 d = v3_scale(lp, 1.0f / len);
    ray_terms rt;
    rt.a = v3_dot(d, d); rt.fourA = 4.0f * rt.a; rt.divTwoA = 1.0f / (2.0f * rt.a);
    for (uint32_t i = 0; i < n; ++i) {
        float t; v3 nn; int hit = 0;
        if (objs[i].type == RTC_OBJ_SPHERE) hit = sphere_trace(&objs[i], light_pos, d, &rt, &t, &nn);
        else if (objs[i].type == RTC_OBJ_PLANE) hit = plane_trace(&objs[i], light_pos, d, &t, &nn);
        if (hit && t < len) return 1;
    }
    return 0;
}

static void ray_trace(const rtc_object* objs, uint32_t n, v3 o, v3 d, uint32_t flags, hit_rec* r)
{
    ray_terms rt;
    rt.a = v3_dot(d, d);                         /* :91 */
    rt.fourA = 4.0f * rt.a;                      /* :92 */
    rt.divTwoA = 1.0f / (2.0f * rt.a);           /* :93 */
    r->color = v3_make(0.f, 0.f, 0.f);
    r->normal = v3_make(0.f, 0.f, 0.f);
    r->distance = 99999999.f;                    /* RayTracing.h:21 */
    r->shading_value = 0.0f;
    r->index = -1;
    r->lit = 1;
    /* objectTraceReturnData lives outside the loop and bHit is never cleared (:95); a stale
     * distance can never be < the running best, so the behaviour equals "test each object". */
    int stale_hit = 0; float stale_dist = 99999999.f; v3 stale_nrm = v3_make(0.f, 0.f, 0.f);
    int hit_any = 0;
    for (uint32_t i = 0; i < n; ++i) {           /* :100 */
        float t; v3 nn;
        int hit = 0;
        if (objs[i].type == RTC_OBJ_PLANE) hit = plane_trace(&objs[i], o, d, &t, &nn);            /* :107-112 */
        else if (objs[i].type == RTC_OBJ_SPHERE) hit = sphere_trace(&objs[i], o, d, &rt, &t, &nn); /* :113-118 */
        if (hit) { stale_hit = 1; stale_dist = t; stale_nrm = nn; }
        if (stale_hit && stale_dist < r->distance) {       /* :123 strict '<': lowest index wins ties */
            hit_any = 1;
            r->distance = stale_dist;
            r->normal = v3_normalize_gpu(stale_nrm);        /* :128-129 (second normalisation) */
            r->shading_value = r->normal.x * 1.0f + r->normal.y * 0.0f + r->normal.z * 0.0f; /* :133 */
            r->color = v3_make(objs[i].color[0], objs[i].color[1], objs[i].color[2]);       /* :134 */
            r->index = (int32_t)i;
        }
    }
    if (!hit_any) return;                        /* :138-141 */
    const v3 point = v3_add(o, v3_scale(d, r->distance));                        /* :149 */
    v3 shading = blinn_phong(v3_div(r->color, 255.0f), point,
                             v3_normalize_gpu(v3_scale(d, -1.0f)), r->normal);   /* :143-152 */
    if ((flags & RTC_FLAG_SHADOWS) && shadow_blocked(objs, n, point, r->normal)) {
        shading = v3_cmul(v3_make(g_light.ambient[0], g_light.ambient[1], g_light.ambient[2]), v3_div(r->color, 255.0f));  /* ambient term only */
        r->lit = 0;
    }
    shading = v3_scale(shading, 255.0f);                                         /* :154 */
    r->color = v3_make(minf_ref(255.0f, shading.x), minf_ref(255.0f, shading.y), minf_ref(255.0f, shading.z)); /* :157 */
}

/* ---- GetASCIICharacter + ASCII ramp (RayTracing.cu:26-39, RayTracing.h:97-115) ---------- */
static const char ASCII_RAMP[68 + 1] =
    " .`^\",:;Il!i><~+_-?*][}{1)(|/tfjrxnuvczmwXYUJCLqpdbkhao#%ZO8B$0QM&W@";
static char ascii_char(float distance, float far_plane, float shading_value)
{
    if (distance > far_plane) return ASCII_RAMP[0];
    int idx = clampi((int)ceilf(shading_value * 67.0f), 1, 68);
    /* Index 68 is one past the reference's table (needs normal.x > 1 by rounding); the
     * reference reads an unspecified byte there.  Pinned here to the last glyph.          */
    if (idx > 67) idx = 67;
    return ASCII_RAMP[idx];
}

/* float -> uint8_t as the reference's host compiler does it (cvttss2si, low byte); the
 * colour path only ever sees [0,255] or NaN (-> 0); RGB_NORMALS sees negatives (wraps). */
static inline uint8_t to_u8(float f)
{
    if (!(f == f)) return 0;
    return (uint8_t)(int32_t)f;
}

/* ---- xterm-256 quantiser, restated from the xterm palette definition (ANSIRGB.h:141-189) --
 * Palette: 16 system colours; 6x6x6 cube on levels {0,95,135,175,215,255}; 24 greys 8+10k. */
static uint32_t g_pal[256];
static uint8_t g_grey_lut[256];
static pthread_once_t g_tab_once = PTHREAD_ONCE_INIT;
static void build_tables(void)
{
    static const uint32_t sys16[16] = {0x000000, 0xcd0000, 0x00cd00, 0xcdcd00, 0x0000ee, 0xcd00cd, 0x00cdcd, 0xe5e5e5,
                                       0x7f7f7f, 0xff0000, 0x00ff00, 0xffff00, 0x5c5cff, 0xff00ff, 0x00ffff, 0xffffff};
    static const uint32_t lv[6] = {0, 95, 135, 175, 215, 255};
    for (int i = 0; i < 16; ++i) g_pal[i] = sys16[i];
    for (int r = 0; r < 6; ++r) for (int g = 0; g < 6; ++g) for (int b = 0; b < 6; ++b)
        g_pal[16 + 36 * r + 6 * g + b] = (lv[r] << 16) | (lv[g] << 8) | lv[b];
    for (int k = 0; k < 24; ++k) { uint32_t v = 8 + 10 * k; g_pal[232 + k] = (v << 16) | (v << 8) | v; }
    /* Grey LUT: nearest of the 30 greys the palette offers (24-step ramp + the 6 cube greys
     * 16,59,102,145,188,231).  Exact ties (v = 13,23,...) go to the darker entry in the lower
     * half (v < 120) and to the brighter entry above -- the behaviour observed from the
     * reference's table, verified entry-by-entry (and on all 2^24 colours) by
     * tests/test_oracle_vs_reference.py.                                                    */
    int cand_idx[30], cand_val[30], nc = 0;
    for (int k = 0; k < 6; ++k) { cand_idx[nc] = 16 + 43 * k; cand_val[nc] = (int)lv[k]; ++nc; }
    for (int k = 0; k < 24; ++k) { cand_idx[nc] = 232 + k; cand_val[nc] = 8 + 10 * k; ++nc; }
    for (int v = 0; v < 256; ++v) {
        int best = 0, bestd = 1 << 30, bestval = 0;
        for (int j = 0; j < nc; ++j) {
            int dd = abs(cand_val[j] - v);
            const int tie_wins = (v < 120) ? (cand_val[j] < bestval) : (cand_val[j] > bestval);
            if (dd < bestd || (dd == bestd && tie_wins)) { bestd = dd; best = cand_idx[j]; bestval = cand_val[j]; }
        }
        g_grey_lut[v] = (uint8_t)best;
    }
}
static inline uint32_t pal_distance(uint32_t x, uint32_t y)          /* ANSIRGB.h:118-124 */
{
    const int32_t rs = (int32_t)((x >> 16) & 255) + (int32_t)((y >> 16) & 255);
    const int32_t r = (int32_t)((x >> 16) & 255) - (int32_t)((y >> 16) & 255);
    const int32_t g = (int32_t)((x >> 8) & 255) - (int32_t)((y >> 8) & 255);
    const int32_t b = (int32_t)(x & 255) - (int32_t)(y & 255);
    return (uint32_t)((1024 + rs) * r * r + 2048 * g * g + (1534 - rs) * b * b);
}
static inline int cube_level(uint32_t v, const uint8_t th[5])
{
    int i = 0;
    while (i < 5 && v >= th[i]) ++i;
    return i;
}
static uint8_t ansi256(uint32_t rgb)
{
    pthread_once(&g_tab_once, build_tables);
    const uint32_t r = (rgb >> 16) & 255, g = (rgb >> 8) & 255, b = rgb & 255;
    if (r == g && g == b) return g_grey_lut[b];                      /* :181-183 */
    const uint32_t lum = (3567664u * r + 11998547u * g + 1211005u * b + (1u << 23)) >> 24; /* :126-134 */
    const uint8_t grey_index = g_grey_lut[lum & 255];
    const uint32_t grey_distance = pal_distance(rgb, g_pal[grey_index]);
    static const uint8_t thr[5] = {38, 115, 155, 196, 235};          /* :18-20 */
    static const uint8_t thg[5] = {36, 116, 154, 195, 235};          /* :25-27 */
    static const uint8_t thb[5] = {35, 115, 155, 195, 235};          /* :32-34 */
    static const uint32_t lv[6] = {0, 95, 135, 175, 215, 255};
    const int ir = cube_level(r, thr), ig = cube_level(g, thg), ib = cube_level(b, thb);
    const uint32_t cube_rgb = (lv[ir] << 16) | (lv[ig] << 8) | lv[ib];
    const uint8_t cube_idx = (uint8_t)(16 + 36 * ir + 6 * ig + ib);
    return pal_distance(rgb, cube_rgb) < grey_distance ? cube_idx : grey_index; /* :188 */
}

/* ---- digit formatter (RayTracing.cu:526-583): NUL-padded 3 decimal digits ------------------ */
static inline void digits3(uint8_t v, char out[3])
{
    out[0] = v >= 100 ? (char)('0' + v / 100) : '\0';
    out[1] = v >= 10 ? (char)('0' + (v / 10) % 10) : '\0';
    out[2] = (char)('0' + v % 10);
}

/* ---- one traced cell --------------------------------------------------------------------- */
typedef struct {
    uint8_t c[3];   /* quantised colour: RGB (RGB modes) or c[0] = xterm index (8-bit modes) */
    uint8_t glyph;  /* cell character                                                        */
    uint8_t hit;    /* distance <= camFarDist (RayTracing.cu:508)                            */
} cell_px;

static void trace_cell(const rtc_object* objs, uint32_t n, const rtc_params* p, rtc_mode mode, uint32_t flags,
                       uint32_t row, uint32_t col, cell_px* px, hit_rec* rec)
{
    const v3 o = v3_make(p->cam_pos[0], p->cam_pos[1], p->cam_pos[2]);
    const v3 d = initial_direction(p, row, col);
    ray_trace(objs, n, o, d, flags, rec);
    px->hit = rec->distance <= p->cam_far;
    px->glyph = ' ';
    px->c[0] = px->c[1] = px->c[2] = 0;
    if (mode == RTC_BIT_ASCII || mode == RTC_RGB_ASCII)
        px->glyph = (uint8_t)ascii_char(rec->distance, p->cam_far, rec->shading_value); /* :204, :368 */
    if (!px->hit) {
        if (mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL) px->c[0] = 16;              /* :248, :328 */
        return;
    }
    uint8_t r8, g8, b8;
    if (mode == RTC_RGB_NORMALS) {                                                      /* :669-709 */
        r8 = to_u8(rec->normal.x * 255); g8 = to_u8(rec->normal.y * 255); b8 = to_u8(rec->normal.z * 255);
    } else {
        r8 = to_u8(rec->color.x); g8 = to_u8(rec->color.y); b8 = to_u8(rec->color.z);   /* :527,:547,:567 */
    }
    if (mode == RTC_BIT_ASCII || mode == RTC_BIT_PIXEL)
        px->c[0] = ansi256(((uint32_t)r8 << 16) + ((uint32_t)g8 << 8) + b8);            /* :210, :291 */
    else { px->c[0] = r8; px->c[1] = g8; px->c[2] = b8; }
}

static inline uint32_t mode_bpp(rtc_mode m) { return (m == RTC_BIT_ASCII || m == RTC_BIT_PIXEL) ? 1u : 3u; }
static inline uint32_t mode_cell(rtc_mode m) { return (m == RTC_BIT_ASCII || m == RTC_BIT_PIXEL) ? 12u : 20u; } /* RayTracing.h:120-123 */
static inline int mode_glyph(rtc_mode m) { return m == RTC_BIT_ASCII || m == RTC_RGB_ASCII; }

/* The cell bytes a RayTrace_* kernel stores (RayTracing.cu:231-251, :312-331, :448-471,
 * :585-608, :727-750).  `fg` selects ESC[38 (ASCII-mode hit) vs ESC[48.                   */
static uint32_t make_cell(rtc_mode mode, const uint8_t* c, uint8_t glyph, int hit, char* out)
{
    const int fg = hit && mode_glyph(mode);
    uint32_t k = 0;
    out[k++] = '\x1b'; out[k++] = '['; out[k++] = fg ? '3' : '4'; out[k++] = '8'; out[k++] = ';';
    if (mode_bpp(mode) == 1) {
        out[k++] = '5'; out[k++] = ';';
        digits3(c[0], out + k); k += 3;
    } else {
        out[k++] = '2'; out[k++] = ';';
        digits3(c[0], out + k); k += 3; out[k++] = ';';
        digits3(c[1], out + k); k += 3; out[k++] = ';';
        digits3(c[2], out + k); k += 3;
    }
    out[k++] = 'm'; out[k++] = (char)glyph;
    return k;
}

/* ---- frame tracing (thread pool over rows; the kernels' 1 thread per cell) -------------- */
typedef struct {
    const rtc_object* objs; uint32_t n; const rtc_params* p; rtc_mode mode; uint32_t flags;
    uint32_t row0, row1, band_row0;
    uint8_t* color; uint8_t* glyph; uint8_t* hitmask; float* dist; int32_t* index; char* raw;
    int tid, nthreads;
} job_t;

static void* trace_rows(void* arg)
{
    job_t* j = (job_t*)arg;
    const uint32_t W = j->p->x - 1, bpp = mode_bpp(j->mode), cs = mode_cell(j->mode);
    for (uint32_t row = j->row0 + (uint32_t)j->tid; row < j->row1; row += (uint32_t)j->nthreads) {
        for (uint32_t col = 0; col < W; ++col) {
            cell_px px; hit_rec rec;
            trace_cell(j->objs, j->n, j->p, j->mode, j->flags, row, col, &px, &rec);
            const size_t i = (size_t)(row - j->band_row0) * W + col;
            if (j->color) for (uint32_t k = 0; k < bpp; ++k) j->color[i * bpp + k] = px.c[k];
            if (j->glyph) j->glyph[i] = px.glyph;
            if (j->hitmask) j->hitmask[i] = px.hit;
            if (j->dist) j->dist[i] = rec.distance;
            if (j->index) j->index[i] = rec.index;
            if (j->raw && j->mode != RTC_SDL)   /* result[row*(x*SIZE) + column*SIZE] (RayTracing.cu:594) */
                make_cell(j->mode, px.c, px.glyph, px.hit,
                          j->raw + ((size_t)row * j->p->x * cs + (size_t)col * cs));
        }
    }
    return NULL;
}

static void run_rows(job_t* proto, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].nthreads = nthreads; }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, trace_rows, &jobs[t]);
    trace_rows(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* Trace rows [row0,row1) into planes indexed relative to row0: colour ((x-1)*bpp per row),
 * glyph, hit mask, distance, object index.  Any output may be NULL.                        */
ORC_API int orc_trace_planes(const rtc_object* objs, uint32_t n, const rtc_params* p, int mode, uint32_t flags,
                             uint32_t row0, uint32_t row1, int nthreads,
                             uint8_t* color, uint8_t* glyph, uint8_t* hitmask, float* dist, int32_t* index)
{
    if (!p || p->x < 1 || row1 > p->y || row0 > row1) return -1;
    pthread_once(&g_tab_once, build_tables);
    job_t j; memset(&j, 0, sizeof j);
    j.objs = objs; j.n = n; j.p = p; j.mode = (rtc_mode)mode; j.flags = flags;
    j.row0 = row0; j.row1 = row1; j.band_row0 = row0;
    j.color = color; j.glyph = glyph; j.hitmask = hitmask; j.dist = dist; j.index = index;
    run_rows(&j, nthreads);
    return 0;
}

/* The raw, un-minimised device buffer of a RayTrace_* launch: 20*x*y bytes, zeroed
 * (RayTracingManager.cu:161-165) then one cell per traced pixel.                           */
ORC_API int orc_trace_raw(const rtc_object* objs, uint32_t n, const rtc_params* p, int mode, uint32_t flags,
                          int nthreads, char* raw /* 20*x*y */)
{
    if (!p || !raw) return -1;
    pthread_once(&g_tab_once, build_tables);
    memset(raw, 0, (size_t)20 * p->x * p->y);
    job_t j; memset(&j, 0, sizeof j);
    j.objs = objs; j.n = n; j.p = p; j.mode = (rtc_mode)mode; j.flags = flags;
    j.row0 = 0; j.row1 = p->y; j.band_row0 = 0; j.raw = raw;
    run_rows(&j, nthreads);
    return 0;
}

/* ---- MinimizeRGB / Minimize8bit (RayTracingManager.cu:251-319 / :181-249): the host's
 * serial byte scan, restated literally.  size = 20*x*y (PrintMachine::GetMaxSize).        */
ORC_API size_t orc_minimize(const char* raw, size_t size, uint32_t x, uint32_t y, int mode, char* out)
{
    const size_t cs = mode_cell((rtc_mode)mode);
    const int ncol = cs == 12 ? 3 : 11;   /* colour bytes compared: offsets 7..9 (8-bit) or 7..17 minus ';' */
    size_t newlines = 0, added = 0;
    const char* latest = NULL;
    for (size_t i = 0; i < size;) {
        const char cur = raw[i];
        if (cur == '\x1b') {
            int differs = (latest == NULL);
            if (!differs) {
                for (int k = 0; k < ncol; ++k) {
                    if (cs == 20 && (k == 3 || k == 7)) continue;   /* the ';' separators are not compared */
                    if (latest[k] != raw[i + 7 + k]) { differs = 1; break; }
                }
            }
            if (differs) {
                latest = raw + i + 7;
                memcpy(out + added, raw + i, cs);   /* whole cell INCLUDING its NUL padding */
                added += cs;
            } else {
                out[added++] = raw[i + cs - 1];     /* only the character */
            }
            i += cs;
        } else if (((i + 1) % (cs * x)) == 0) {
            ++newlines;
            out[added++] = '\n';
            ++i;
            if (newlines == y) break;
        } else {
            ++i;
        }
    }
    return added;
}

/* The same stream from the planes by the LOCAL rule the GPU encoder uses:
 * full cell iff colour key != key of the previous traced cell in raster order (carried
 * across rows), first cell always full, one '\n' per row.  tests/ prove it byte-identical
 * to orc_minimize(orc_trace_raw(...)).                                                     */
ORC_API size_t orc_encode_planes(const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y, int mode, char* out)
{
    const rtc_mode m = (rtc_mode)mode;
    size_t added = 0;
    if (m == RTC_SDL) { for (uint32_t r = 0; r < y; ++r) out[added++] = '\n'; return added; }
    const uint32_t W = x - 1, bpp = mode_bpp(m);
    const int has_glyph = mode_glyph(m) && glyph != NULL;
    for (uint32_t r = 0; r < y; ++r) {
        for (uint32_t c = 0; c < W; ++c) {
            const size_t i = (size_t)r * W + c;
            const uint8_t g = has_glyph ? glyph[i] : (uint8_t)' ';
            int full = (i == 0);
            if (!full) full = memcmp(color + i * bpp, color + (i - 1) * bpp, bpp) != 0;
            if (full) added += make_cell(m, color + i * bpp, g, /*hit=*/g != ' ', out + added);
            else out[added++] = (char)g;
        }
        out[added++] = '\n';
    }
    return added;
}

/* RayTracingManager::Update minus the physics step: trace + minimise.  Returns stream size. */
ORC_API size_t orc_render(const rtc_object* objs, uint32_t n, const rtc_params* p, int mode, uint32_t flags,
                          int nthreads, char* raw_scratch /* 20*x*y */, char* out)
{
    orc_trace_raw(objs, n, p, mode, flags, nthreads, raw_scratch);
    return orc_minimize(raw_scratch, (size_t)20 * p->x * p->y, p->x, p->y, mode, out);
}

/* ---- UpdateObjects / Sphere::Update (RayTracingManager.cu:10-44, :89-107; Sphere.cu:15-23) */
ORC_API int orc_update_objects(rtc_object* objs, uint32_t n, double dt, uint32_t flags)
{
    /* block = count threads: CUDA rejects count > 1024 (and 0), the kernel never runs. */
    if ((flags & RTC_FLAG_UPDATE_REF_LAUNCH_LIMIT) && (n > 1024 || n == 0)) return 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (objs[i].type != RTC_OBJ_SPHERE) continue;            /* Plane::Update is a no-op (Plane.cu:14-18) */
        /* `speed * mover * dt` with dt long double: on the device (and MSVC) long double is
         * binary64, so the product and the sum are formed in double and rounded to float. */
        const double prod = (double)(objs[i].speed * (float)objs[i].mover) * dt;
        float yy = (float)((double)objs[i].center[1] + prod);    /* Sphere.cu:17 */
        if (yy < -10.0f || yy > 10.0f) {                         /* :18 */
            yy = clampf(yy, -10.0f, 10.0f);                      /* :20 */
            objs[i].mover *= -1;                                 /* :21 */
        }
        objs[i].center[1] = yy;
    }
    return 0;
}

/* ---- Camera3D (Camera3D.cpp:8-48 Init, :51-98 Update, :207-376 GetInverseVMatrix) and the
 * parameter block Engine3D::Render assembles (Engine3D.cpp:88-97).                         */
ORC_API int orc_camera_params(uint32_t x, uint32_t y, const float pos[3], const float rot[3],
                              float pixel_aspect, rtc_params* out)
{
    if (pixel_aspect == 0.0f) pixel_aspect = 0.01f;              /* Camera3D.cpp:17 */
    const float fov = (float)(3.14159265358979323846) / 1.5f;    /* Camera3D.h:79, .cpp:10 */
    const float width = (float)x, height = (float)y;
    const float aspect = width / (pixel_aspect * width * height);
    const float e = 1.0f / tanf(fov / 2.0f);                     /* :19 */
    const float p = rot[0], yw = rot[1];
    const v3 fwd = v3_make(-sinf(yw), -sinf(p) * cosf(yw), -cosf(p) * cosf(yw));   /* :57-59 */
    const v3 right = v3_make(cosf(yw), -sinf(p) * sinf(yw), -cosf(p) * sinf(yw));  /* :65-67 */
    const v3 up = v3_make(0.0f, cosf(p), -sinf(p));                                /* :73-75 */
    float m[16] = { right.x, up.x, fwd.x, pos[0],                                  /* :79-98 */
                    right.y, up.y, fwd.y, pos[1],
                    right.z, up.z, fwd.z, pos[2],
                    0.0f, 0.0f, 0.0f, 1.0f };
    /* Cofactor inverse: entry (r,c) is the signed 3x3 minor that deletes row c / column r of
     * m, expanded in the fixed six-term order the reference writes out (:210-343), divided by
     * det = row1 . first column of the cofactor matrix (:345-356).                          */
    float inv[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
        int R[3], C[3], k = 0;
        for (int i = 0; i < 4; ++i) if (i != c) R[k++] = i;
        k = 0;
        for (int i = 0; i < 4; ++i) if (i != r) C[k++] = i;
#define M_(i, j) m[4 * (i) + (j)]
        const float t1 = M_(R[0], C[0]) * M_(R[1], C[1]) * M_(R[2], C[2]);
        const float t2 = M_(R[0], C[0]) * M_(R[1], C[2]) * M_(R[2], C[1]);
        const float t3 = M_(R[1], C[0]) * M_(R[0], C[1]) * M_(R[2], C[2]);
        const float t4 = M_(R[1], C[0]) * M_(R[0], C[2]) * M_(R[2], C[1]);
        const float t5 = M_(R[2], C[0]) * M_(R[0], C[1]) * M_(R[1], C[2]);
        const float t6 = M_(R[2], C[0]) * M_(R[0], C[2]) * M_(R[1], C[1]);
#undef M_
        /* Odd cofactors are written with the signs flipped term by term (":219 -a*b*c + ...");
         * that differs from -(...) in the sign of a zero result, so keep both spellings.     */
        inv[4 * r + c] = ((r + c) & 1) ? (-t1 + t2 + t3 - t4 - t5 + t6) : (t1 - t2 - t3 + t4 + t5 - t6);
    }
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0f) return -1;                                  /* :349-352 assert */
    det = 1.0f / det;
    for (int i = 0; i < 16; ++i) out->inv_view[i] = inv[i] * det;
    out->cam_pos[0] = pos[0]; out->cam_pos[1] = pos[1]; out->cam_pos[2] = pos[2];
    out->x = x; out->y = y;
    out->element1 = e / aspect;   /* m_pMatrix.row1.x (:29) */
    out->element2 = e;            /* m_pMatrix.row2.y (:35) */
    out->cam_far = 250.0f;        /* Camera3D.h:75 */
    return 0;
}

/* The reference default scene (Scene3D.cpp:28-33). */
ORC_API uint32_t orc_default_scene(rtc_object* out, uint32_t cap)
{
    static const float S[5][7] = { {7.0f, 0.0f, 10.0f, 20.0f, 255.0f, 1.0f, 1.0f},
                                   {6.0f, 5.0f, 10.0f, 20.0f, 1.0f, 255.0f, 1.0f},
                                   {10.0f, 10.0f, 10.0f, 40.0f, 1.0f, 1.0f, 255.0f},
                                   {3.0f, 5.0f, 10.0f, 20.0f, 225.0f, 210.0f, 20.0f},
                                   {4.0f, -5.0f, 10.0f, 40.0f, 225.0f, 10.0f, 220.0f} };
    if (cap < 6) return 0;
    memset(out, 0, 6 * sizeof *out);
    for (int i = 0; i < 5; ++i) {
        out[i].type = RTC_OBJ_SPHERE; out[i].radius = S[i][0];
        out[i].center[0] = S[i][1]; out[i].center[1] = S[i][2]; out[i].center[2] = S[i][3];
        out[i].color[0] = S[i][4]; out[i].color[1] = S[i][5]; out[i].color[2] = S[i][6];
        out[i].speed = 1.0f; out[i].mover = -1;                  /* Sphere.cu:9 (speed is rand() there) */
    }
    out[5].type = RTC_OBJ_PLANE;
    out[5].center[0] = 0.0f; out[5].center[1] = -3.0f; out[5].center[2] = 30.0f;
    const v3 nn = v3_normalize_host(v3_make(0.0f, 1.0f, 0.0f));  /* Plane.cu:9 */
    out[5].normal[0] = nn.x; out[5].normal[1] = nn.y; out[5].normal[2] = nn.z;
    out[5].color[0] = out[5].color[1] = out[5].color[2] = 100.0f;
    out[5].width = 10.0f; out[5].height = 20.0f;
    return 6;
}

/* ---- per-function known-answer hooks (mirrors of oracle/ref_build/ref_driver.cpp) -------- */
ORC_API int orc_sphere_trace(const float c[3], float radius, const float o[3], const float d[3], float* dist, float nrm[3])
{
    rtc_object s; memset(&s, 0, sizeof s); s.type = RTC_OBJ_SPHERE; s.radius = radius;
    s.center[0] = c[0]; s.center[1] = c[1]; s.center[2] = c[2];
    const v3 dd = v3_make(d[0], d[1], d[2]);
    ray_terms rt; rt.a = v3_dot(dd, dd); rt.fourA = 4.0f * rt.a; rt.divTwoA = 1.0f / (2.0f * rt.a);
    v3 n = v3_make(0.f, 0.f, 0.f); float t = 99999999.f;
    const int hit = sphere_trace(&s, v3_make(o[0], o[1], o[2]), dd, &rt, &t, &n);
    *dist = t; nrm[0] = n.x; nrm[1] = n.y; nrm[2] = n.z;
    return hit;
}
ORC_API int orc_plane_normal(const float normal[3], float out[3])
{
    const v3 n = v3_normalize_host(v3_make(normal[0], normal[1], normal[2]));
    out[0] = n.x; out[1] = n.y; out[2] = n.z; return 0;
}
ORC_API int orc_plane_trace(const float c[3], const float normal[3], float w, float h,
                            const float o[3], const float d[3], float* dist, float nrm[3])
{
    rtc_object p; memset(&p, 0, sizeof p); p.type = RTC_OBJ_PLANE; p.width = w; p.height = h;
    p.center[0] = c[0]; p.center[1] = c[1]; p.center[2] = c[2];
    orc_plane_normal(normal, p.normal);
    v3 n = v3_make(0.f, 0.f, 0.f); float t = 99999999.f;
    const int hit = plane_trace(&p, v3_make(o[0], o[1], o[2]), v3_make(d[0], d[1], d[2]), &t, &n);
    *dist = t; nrm[0] = n.x; nrm[1] = n.y; nrm[2] = n.z;
    return hit;
}
ORC_API int orc_initial_direction(const rtc_params* p, uint32_t row, uint32_t col, float d[3])
{
    const v3 v = initial_direction(p, row, col); d[0] = v.x; d[1] = v.y; d[2] = v.z; return 0;
}
ORC_API int orc_raytrace(const rtc_object* objs, uint32_t n, const float o[3], const float d[3], uint32_t flags,
                         float* dist, float nrm[3], float col[3], float* shading_value, int32_t* index)
{
    hit_rec r; ray_trace(objs, n, v3_make(o[0], o[1], o[2]), v3_make(d[0], d[1], d[2]), flags, &r);
    *dist = r.distance; *shading_value = r.shading_value; if (index) *index = r.index;
    nrm[0] = r.normal.x; nrm[1] = r.normal.y; nrm[2] = r.normal.z;
    col[0] = r.color.x; col[1] = r.color.y; col[2] = r.color.z;
    return 0;
}
ORC_API int orc_blinn_phong(const float kd[3], const float point[3], const float view[3], const float nrm[3], float out[3])
{
    const v3 r = blinn_phong(v3_make(kd[0], kd[1], kd[2]), v3_make(point[0], point[1], point[2]),
                             v3_make(view[0], view[1], view[2]), v3_make(nrm[0], nrm[1], nrm[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; return 0;
}
ORC_API int orc_set_light(const rtc_light* l)
{
    static const rtc_light def = {{1.0f, 50.0f, 0.0f}, 1.0f, 2000.0f, 1.0f, 3000.0f, {0.2f, 0.2f, 0.2f}, 1.0f};
    g_light = l ? *l : def;
    return 0;
}
ORC_API int orc_ansi256_range(uint32_t first, uint32_t count, uint8_t* out)
{
    for (uint32_t i = 0; i < count; ++i) out[i] = ansi256(first + i);
    return 0;
}
ORC_API int orc_ascii_char(float distance, float far_plane, float shading_value)
{
    return (int)(unsigned char)ascii_char(distance, far_plane, shading_value);
}
ORC_API int orc_digits3(uint32_t v, char out[3]) { digits3((uint8_t)v, out); return 0; }
ORC_API int orc_make_cell(int mode, const uint8_t* c, uint8_t glyph, int hit, char* out)
{
    return (int)make_cell((rtc_mode)mode, c, glyph, hit, out);
}

/* ---- CPU-baseline timing leg: trace `rows` rows starting at row0 with nthreads threads --- */
ORC_API double orc_time_trace(const rtc_object* objs, uint32_t n, const rtc_params* p, int mode, uint32_t flags,
                              uint32_t row0, uint32_t row1, int nthreads)
{
    const uint32_t W = p->x - 1;
    uint8_t* color = (uint8_t*)malloc((size_t)(row1 - row0) * W * 3);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    orc_trace_planes(objs, n, p, mode, flags, row0, row1, nthreads, color, NULL, NULL, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    free(color);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
