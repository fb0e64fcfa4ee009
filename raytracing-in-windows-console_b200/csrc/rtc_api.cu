// rtc_api.cu -- the C-ABI of include/rtc.h: context, scene store, frame driver.
//
// Replaces the host side of RayTracingManager (reference RayTracingManager.cu:53-165) and the
// device-array bookkeeping of Scene3D (Scene3D.cpp:7-164).  Differences by design:
//   * the scene lives in ONE device array of 64-byte PODs uploaded in a single async copy
//     when dirty (the reference issues 2 blocking cudaMemcpy per object, Scene3D.cpp:47-59);
//   * nothing is memset per frame (the reference clears 20*x*y bytes, :161-165) because the
//     20-byte raw cell buffer does not exist;
//   * only the minimised stream crosses PCIe (the reference copies the whole raw buffer to
//     pageable memory, :143, then minimises on one host thread, :146);
//   * errors are returned, never exit()ed (pch.h:45-53).
// There is no CPU fallback anywhere in this file: without a CUDA device rtc_create fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "rtc_ctx.h"

namespace rtc {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace rtc

using rtc::DevBuf;
using rtc::PinBuf;
using rtc::fail;
using rtc::mode_bpp;
using rtc::mode_cell;
using rtc::mode_has_glyph;
using rtc::mode_is_8bit;

namespace {

// Sphere slots in Morton (Z-curve) order of their centres, so that the 4 spheres that share a packed group -- and a
// bounding cone for per-tile culling -- are neighbours in space.  The order is irrelevant to the result (the accept
// rule is the lexicographic minimum of (distance, object index)); it is camera-independent, hence done here on the
// host when the scene changes rather than per frame.
void morton_order(const std::vector<rtc_object>& objs, std::vector<int32_t>& sphere_obj)
{
    const size_t n = sphere_obj.size();
    if (n < 8) return;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int32_t i : sphere_obj)
        for (int a = 0; a < 3; ++a) {
            const float v = objs[i].center[a];
            if (v == v && fabsf(v) < 1.0e30f) { lo[a] = v < lo[a] ? v : lo[a]; hi[a] = v > hi[a] ? v : hi[a]; }
        }
    std::vector<std::pair<uint32_t, int32_t>> keyed(n);
    for (size_t k = 0; k < n; ++k) {
        uint32_t code = 0;
        for (int a = 0; a < 3; ++a) {
            const float v = objs[sphere_obj[k]].center[a];
            const float span = hi[a] - lo[a];
            uint32_t q = 0;
            if (v == v && fabsf(v) < 1.0e30f && span > 0.0f) {
                const float t = (v - lo[a]) / span * 1023.0f;
                q = t <= 0.0f ? 0u : t >= 1023.0f ? 1023u : (uint32_t)t;
            }
            q = (q | (q << 16)) & 0x030000FFu; q = (q | (q << 8)) & 0x0300F00Fu;          // spread 10 bits to every third bit
            q = (q | (q << 4)) & 0x030C30C3u;  q = (q | (q << 2)) & 0x09249249u;
            code |= q << a;
        }
        keyed[k] = std::make_pair(code, sphere_obj[k]);
    }
    std::sort(keyed.begin(), keyed.end());                      // ties fall back to the object index: deterministic
    for (size_t k = 0; k < n; ++k) sphere_obj[k] = keyed[k].second;
}

int upload_scene(rtc_ctx* c)
{
    if (!c->scene_dirty) return RTC_OK;
    const size_t n = c->objs.size();
    c->sphere_obj.clear();
    c->plane_obj.clear();
    // The Morton order only depends on which objects are spheres and where their centres are: a scene that is uploaded
    // again with the same geometry (per-frame uploads of a static or colour-animated scene) keeps its order.
    bool same_geometry = c->order_key.size() == 4 * n;
    for (size_t i = 0; i < n && same_geometry; ++i) {
        const rtc_object& o = c->objs[i];
        const float* k = &c->order_key[4 * i];
        same_geometry = memcmp(k, o.center, 12) == 0 && k[3] == (float)o.type;
    }
    if (same_geometry && n > 0) {
        c->sphere_obj = c->sphere_order;
        c->plane_obj = c->plane_order;
    } else {
        for (size_t i = 0; i < n; ++i) {
            if (c->objs[i].type == RTC_OBJ_SPHERE) c->sphere_obj.push_back((int32_t)i);
            else if (c->objs[i].type == RTC_OBJ_PLANE) c->plane_obj.push_back((int32_t)i);
        }
        morton_order(c->objs, c->sphere_obj);
        c->order_key.resize(4 * n);
        for (size_t i = 0; i < n; ++i) {
            memcpy(&c->order_key[4 * i], c->objs[i].center, 12);
            c->order_key[4 * i + 3] = (float)c->objs[i].type;
        }
        c->sphere_order = c->sphere_obj;
        c->plane_order = c->plane_obj;
    }
    const size_t n_slots = (c->sphere_obj.size() + 3) & ~(size_t)3;
    // Stage in pinned memory (two slots: the previous upload may still be in flight), then ONE async copy.
    const size_t b_objs = (n * sizeof(rtc_object) + 63) & ~(size_t)63, b_sph = (n_slots * sizeof(int32_t) + 63) & ~(size_t)63,
                 b_pl = (c->plane_obj.size() * sizeof(int32_t) + 63) & ~(size_t)63, b_kd = n * sizeof(float4);
    const size_t b_all = b_objs + b_sph + b_pl + b_kd + 64;
    // Next ring slot: stage in pinned memory, copy on the upload stream, make the frame stream wait for the copy.
    const int b = (c->scene_cur + 1) % rtc_ctx::kSceneRing;
    if (c->scene_up_pending[b]) {                               // the previous upload out of this staging slot (4 uploads ago)
        CK(cudaEventSynchronize(c->ev_scene_up[b]));
        c->scene_up_pending[b] = false;
    }
    PinBuf<unsigned char>& st = c->h_scene[b];
    if (b_all > st.cap || b_all > c->d_scene[b].cap) {
        // Grow EVERY ring slot now, in one stall (pinned allocations synchronise every device of the process; a scene that
        // grows would otherwise stall four times, once per slot).  After the two synchronisations nothing reads any slot,
        // and the current slot's contents are dead: this upload replaces the scene (the host copy is authoritative here).
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaStreamSynchronize(c->upload_stream));
        for (int i = 0; i < rtc_ctx::kSceneRing; ++i) {
            c->scene_up_pending[i] = false;
            if (b_all > c->h_scene[i].cap) CK(c->h_scene[i].ensure(b_all * 2));
            if (b_all > c->d_scene[i].cap) CK(c->d_scene[i].ensure(b_all * 2));
        }
    }
    unsigned char* h = st.p;
    if (n) memcpy(h, c->objs.data(), n * sizeof(rtc_object));
    if (!c->sphere_obj.empty()) memcpy(h + b_objs, c->sphere_obj.data(), c->sphere_obj.size() * sizeof(int32_t));
    if (!c->plane_obj.empty()) memcpy(h + b_objs + b_sph, c->plane_obj.data(), c->plane_obj.size() * sizeof(int32_t));
    // kd = colour / 255 (RayTracing.cu:144 divides per pixel; the quotient only depends on the object).  IEEE division,
    // no contraction (-ffp-contract=off): the same float the device's __fdiv_rn would give.
    float* kd = reinterpret_cast<float*>(h + b_objs + b_sph + b_pl);
    for (size_t i = 0; i < n; ++i) {
        kd[4 * i + 0] = c->objs[i].color[0] / 255.0f; kd[4 * i + 1] = c->objs[i].color[1] / 255.0f;
        kd[4 * i + 2] = c->objs[i].color[2] / 255.0f; kd[4 * i + 3] = 0.0f;
    }
    if (c->scene_rd_recorded[b]) CK(cudaStreamWaitEvent(c->upload_stream, c->ev_scene_rd[b], 0));   // frames still reading this device slot
    CK(cudaMemcpyAsync(c->d_scene[b].p, h, b_all - 64, cudaMemcpyHostToDevice, c->upload_stream));
    CK(cudaEventRecord(c->ev_scene_up[b], c->upload_stream));
    c->scene_up_pending[b] = true;
    CK(cudaStreamWaitEvent(c->stream, c->ev_scene_up[b], 0));
    c->scene_cur = b;
    c->d_objs.p = reinterpret_cast<rtc_object*>(c->d_scene[b].p);
    c->d_sphere_obj.p = reinterpret_cast<int32_t*>(c->d_scene[b].p + b_objs);
    c->d_plane_obj.p = reinterpret_cast<int32_t*>(c->d_scene[b].p + b_objs + b_sph);
    c->d_kd.p = reinterpret_cast<float4*>(c->d_scene[b].p + b_objs + b_sph + b_pl);
    c->scene_dirty = false;
    return RTC_OK;
}

// Everything enqueued so far on the frame stream may read the current scene slot: the upload stream must not overwrite it
// before this point of the frame stream has been reached.
int mark_scene_read(rtc_ctx* c)
{
    CK(cudaEventRecord(c->ev_scene_rd[c->scene_cur], c->stream));
    c->scene_rd_recorded[c->scene_cur] = true;
    return RTC_OK;
}

int refresh_host_scene(rtc_ctx* c)
{
    if (!c->host_stale) return RTC_OK;
    if (!c->objs.empty()) {
        CK(cudaMemcpyAsync(c->objs.data(), c->d_objs.p, c->objs.size() * sizeof(rtc_object), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    c->host_stale = false;
    return RTC_OK;
}

rtc::FrameParams make_frame(const rtc_params* p, uint32_t row0, uint32_t row1)
{
    rtc::FrameParams f;
    for (int i = 0; i < 12; ++i) f.m[i] = p->inv_view[i];
    for (int i = 0; i < 3; ++i) f.cam[i] = p->cam_pos[i];
    f.e1 = p->element1; f.e2 = p->element2; f.far_dist = p->cam_far;
    f.fx = (float)p->x; f.fy = (float)p->y;     // size_t -> float in the reference (RayTracing.cu:16-17)
    f.x = p->x; f.y = p->y; f.row0 = row0; f.row1 = row1;
    return f;
}

}  // namespace

namespace rtc {

// The encoder as one step: scratch, parity, launches.  The scratch is zeroed when (re)allocated; afterwards every count
// launch zeroes the accumulators of the NEXT one, so the parity may flip only when a count kernel is really enqueued
// (SDL frames, 1-column consoles and failed renders launch none and must leave it alone).
int do_encode(rtc_ctx* c, const uint8_t* d_color, const uint8_t* d_glyph, uint32_t x, uint32_t rows, int mode, char* d_out,
              size_t cap, unsigned long long* d_total, bool continues)
{
    const bool counts = mode != RTC_SDL && x > 1u && rows > 0u;
    if (counts) {
        const size_t need = rtc::encode_state_bytes((uint64_t)(x - 1u) * rows);
        if (need > c->d_desc.cap) {
            CK(c->d_desc.ensure(need + need / 2));              // head-room: growing frames do not reallocate every time
            CK(cudaMemsetAsync(c->d_desc.p, 0, c->d_desc.cap, c->stream));
        }
    }
    const uint32_t parity = c->enc_parity ^ (counts ? 1u : 0u);
    const cudaError_t e = rtc::launch_encode(c->stream, d_color, d_glyph, x, rows, mode, d_out, cap, d_total, c->d_desc.p,
                                             c->d_desc.cap, parity, continues);
    if (e != cudaSuccess) {
        // a launch may or may not have run: put the accumulators of both parities back into the known state
        if (c->d_desc.p) cudaMemsetAsync(c->d_desc.p, 0, c->d_desc.cap, c->stream);
        return fail(RTC_ERR_CUDA, "ANSI encoder launch failed: %s", cudaGetErrorString(e));
    }
    c->enc_parity = parity;
    return RTC_OK;
}

// trace (scene hoist in its prologue, shade in its tile epilogue) for rows [row0,row1) into colour/glyph planes (band-relative).
int trace_shade(rtc_ctx* c, const rtc_params* p, int mode, uint32_t flags, uint32_t row0, uint32_t row1,
                uint8_t* d_color, uint8_t* d_glyph, bool record_events)
{
    if (!p) return fail(RTC_ERR_INVALID, "params is NULL");
    if (mode < RTC_BIT_ASCII || mode > RTC_SDL) return fail(RTC_ERR_INVALID, "invalid rendering mode %d", mode);
    if (p->x < 1 || p->y < 1) return fail(RTC_ERR_INVALID, "invalid console size %ux%u", p->x, p->y);
    if (row0 > row1 || row1 > p->y) return fail(RTC_ERR_INVALID, "invalid row band [%u,%u) of %u", row0, row1, p->y);
    int rc = upload_scene(c);
    if (rc) return rc;
    const uint32_t W = p->x - 1u;
    const size_t n_px = (size_t)(row1 - row0) * W;
    const int n_spheres = (int)c->sphere_obj.size();
    const int n_slots = (n_spheres + 3) & ~3;
    const int n_planes = (int)c->plane_obj.size();
    c->last_launches = 0;
    if (record_events) CK(cudaEventRecord(c->ev[0], c->stream));
    const bool cull = (flags & RTC_FLAG_CULL) != 0;
    // test counts of this frame: parity p; every launch of the frame zeroes parity p ^ 1 for the next one (so the parity
    // only moves when something is launched)
    if (n_px > 0) c->stats_parity ^= 1u;
    unsigned long long* stats = c->d_counters.p + rtc::kStatsCounter + 2 * c->stats_parity;
    unsigned long long* stats_zero = c->d_counters.p + rtc::kStatsCounter + 2 * (c->stats_parity ^ 1u);
    // The screen-affine packed filter (rtc_trace.cu) assumes what every camera gives: a near-orthonormal 3x3 inverse view
    // matrix (its error bound is relative to |w| = |col2 + vx col0 + vy col1|; cancellation between skewed columns would
    // void it).  Anything else -- the C-ABI accepts arbitrary matrices -- runs the dot-product form of the filter.
    const rtc::FrameParams fp0 = make_frame(p, row0, row1);
    bool affine = getenv("RTC_NO_AFFINE") == nullptr;
    for (int a = 0; a < 3 && affine; ++a)
        for (int b = a; b < 3; ++b) {
            const double dot = (double)fp0.m[a] * fp0.m[b] + (double)fp0.m[4 + a] * fp0.m[4 + b] + (double)fp0.m[8 + a] * fp0.m[8 + b];
            if (!(fabs(dot - (a == b ? 1.0 : 0.0)) < 1.0e-3)) affine = false;
        }
    const bool packet = affine && (flags & RTC_FLAG_PACKET) != 0;   // (the packet filter is a form of the screen-affine one)
    const bool shadows = (flags & RTC_FLAG_SHADOWS) != 0 && mode != RTC_SDL && mode != RTC_RGB_NORMALS;
    // Without shadow rays the ray kernel shades + quantises in its tile epilogue (one launch, no hit-record round trip);
    // hit records are then written only on request.  With shadow rays the records feed the light-origin pass and the
    // stand-alone shade kernel runs after it.
    const bool fused = !shadows && mode != RTC_SDL;
    const int shade_mode = (mode == RTC_RGB_NORMALS && (flags & RTC_FLAG_NORMALS_SATURATE)) ? rtc::kModeNormalsSaturate : mode;
    const bool keep_hits = (flags & RTC_FLAG_KEEP_HITS) != 0 || shadows;
    if (record_events) CK(cudaEventRecord(c->ev[1], c->stream));
    c->hits_valid = false;
    if (n_px > 0) {
        const rtc::FrameParams fp = make_frame(p, row0, row1);
        const rtc::TracePlan plan = rtc::plan_trace(p->x, row1 - row0, n_slots, c->sm_count, packet);
        const int n_chunks = n_slots == 0 ? 1 : (n_slots + plan.max_slots - 1) / plan.max_slots;
        if (n_chunks > rtc::kMaxChunks) return fail(RTC_ERR_CAPACITY, "too many spheres (%d)", n_spheres);
        const unsigned long long tickets = rtc::trace_tickets(p->x, row1 - row0, c->sm_count, plan.threads, plan.rays);
        if (keep_hits || n_chunks > 1) {
            CK(c->d_hit_t.ensure(n_px));
            CK(c->d_hit_idx.ensure(n_px));
        }
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int s0 = ch * plan.max_slots;
            const int slots = n_slots - s0 < plan.max_slots ? n_slots - s0 : plan.max_slots;
            const int sph = n_spheres - s0 < slots ? n_spheres - s0 : slots;
            const bool last = ch == n_chunks - 1;
            CK(rtc::launch_trace(c->stream, c->sm_count, fp, c->d_sphere_obj.p + s0,
                                 sph, slots, c->d_objs.p, c->d_plane_obj.p, last ? n_planes : 0, c->d_hit_t.p,
                                 c->d_hit_idx.p, c->d_counters.p + ch, c->ticket_base[ch], ch > 0 ? 1 : 0, nullptr, nullptr, plan.threads,
                                 cull, stats, stats_zero, c->shade, fused && last ? shade_mode : -1, d_color, d_glyph,
                                 !last || keep_hits, c->d_kd.p, affine, plan.rays, packet));
            c->ticket_base[ch] += tickets;
            c->last_launches++;
        }
        c->hits_valid = keep_hits;
        if (shadows) {                                          // second pass: one ray per shaded pixel, cast from the light
            CK(c->d_shadow.ensure(n_px));
            for (int ch = 0; ch < n_chunks; ++ch) {
                const int s0 = ch * plan.max_slots;
                const int slots = n_slots - s0 < plan.max_slots ? n_slots - s0 : plan.max_slots;
                const int sph = n_spheres - s0 < slots ? n_spheres - s0 : slots;
                const bool last = ch == n_chunks - 1;
                CK(rtc::launch_trace(c->stream, c->sm_count, fp, c->d_sphere_obj.p + s0, sph, slots,
                                     c->d_objs.p, c->d_plane_obj.p, last ? n_planes : 0,
                                     c->d_hit_t.p, c->d_hit_idx.p, c->d_counters.p + rtc::kMaxChunks + ch, c->ticket_base[rtc::kMaxChunks + ch],
                                     ch > 0 ? 1 : 0, c->shade.light, c->d_shadow.p, plan.threads, cull, stats + 1, stats_zero, c->shade, -1,
                                     nullptr, nullptr, false, nullptr, false, plan.rays, false));
                c->ticket_base[rtc::kMaxChunks + ch] += tickets;
                c->last_launches++;
            }
        }
        if (record_events) CK(cudaEventRecord(c->ev[2], c->stream));
        if (shadows) {
            CK(rtc::launch_shade(c->stream, fp, c->shade, mode, c->d_objs.p, c->d_kd.p, c->d_hit_t.p, c->d_hit_idx.p, c->d_shadow.p,
                                 d_color, d_glyph));
            c->last_launches++;
        }
        if (record_events) CK(cudaEventRecord(c->ev[3], c->stream));
    } else if (record_events) {
        CK(cudaEventRecord(c->ev[2], c->stream));
        CK(cudaEventRecord(c->ev[3], c->stream));
    }
    return mark_scene_read(c);
}

}  // namespace rtc

using rtc::do_encode;
using rtc::trace_shade;

extern "C" {

const char* rtc_last_error(void) { return rtc::g_err; }
const char* rtc_version(void) { return "rtc_b200 0.1 (sm_100a)"; }

uint32_t rtc_mode_bpp(rtc_mode mode) { return mode_bpp(mode); }
uint32_t rtc_mode_has_glyph(rtc_mode mode) { return mode_has_glyph(mode) ? 1u : 0u; }
size_t rtc_encode_capacity(uint32_t x, uint32_t y, rtc_mode mode)
{
    if (x == 0) return 0;
    return (size_t)(x - 1u) * y * mode_cell(mode) + y + 64;
}

int rtc_create(rtc_ctx** out, int device)
{
    if (!out) return fail(RTC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(RTC_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n_dev) return fail(RTC_ERR_INVALID, "device %d out of range (0..%d)", device, n_dev - 1);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(RTC_ERR_CUDA, "device %d is sm_%d%d; this build targets sm_100a only", device, prop.major, prop.minor);
    rtc_ctx* c = new (std::nothrow) rtc_ctx();
    if (!c) return fail(RTC_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    cudaDeviceGetAttribute(&c->clock_khz, cudaDevAttrClockRate, device);
#define CKC(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e2_ = (call);                                                                         \
        if (e2_ != cudaSuccess) {                                                                         \
            rtc_destroy(c);                                                                               \
            return fail(RTC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e2_), __FILE__, __LINE__); \
        }                                                                                                 \
    } while (0)
    CKC(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto& ev : c->ev) CKC(cudaEventCreate(&ev));
    CKC(rtc::configure_trace());
    CKC(rtc::configure_encode());
    CKC(c->d_counters.ensure(rtc::kNumCounters));
    CKC(cudaMemset(c->d_counters.p, 0, rtc::kNumCounters * sizeof(unsigned long long)));
    CKC(c->d_total.ensure(2));
    CKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& ev : c->ev_total) CKC(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c->ev_scene_up) CKC(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : c->ev_scene_rd) CKC(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CKC(cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking));
    CKC(c->d_sink.ensure(4));
    CKC(c->h_total.ensure(2));
#undef CKC
    *out = c;
    return RTC_OK;
}

void rtc_destroy(rtc_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->upload_stream) { cudaStreamSynchronize(c->upload_stream); cudaStreamDestroy(c->upload_stream); }
    for (auto& b : c->d_scene) b.release();
    c->d_shadow.release();
    c->d_hit_t.release(); c->d_hit_idx.release(); c->d_color.release(); c->d_glyph.release(); c->d_out[0].release(); c->d_out[1].release();
    c->d_desc.release(); c->d_counters.release(); c->d_total.release(); c->d_sink.release();
    c->h_total.release(); c->h_out[0].release(); c->h_out[1].release(); for (auto& b : c->h_scene) b.release();
    c->h_color.release(); c->h_glyph.release();
    for (auto& ev : c->ev_total) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_scene_up) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_scene_rd) if (ev) cudaEventDestroy(ev);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    c->h_hit_t.release(); c->h_hit_idx.release();
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int rtc_set_stream(rtc_ctx* c, void* cuda_stream)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return RTC_OK;
}

int rtc_device_info(rtc_ctx* c, int* sm_count, int* clock_khz, size_t* smem_optin)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (sm_count) *sm_count = c->sm_count;
    if (clock_khz) *clock_khz = c->clock_khz;
    if (smem_optin) *smem_optin = c->smem_optin;
    return RTC_OK;
}

int rtc_synchronize(rtc_ctx* c)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return RTC_OK;
}

int rtc_set_light(rtc_ctx* c, const rtc_light* l)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    static_assert(sizeof(rtc_light) == sizeof(rtc::ShadeParams), "rtc_light and ShadeParams are the same 11 floats");
    const rtc_light def = {{1.0f, 50.0f, 0.0f}, 1.0f, 2000.0f, 1.0f, 3000.0f, {0.2f, 0.2f, 0.2f}, 1.0f};
    memcpy(&c->shade, l ? l : &def, sizeof c->shade);
    return RTC_OK;
}

int rtc_resize(rtc_ctx* c, uint32_t x, uint32_t y)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (x < 1 || y < 1) return fail(RTC_ERR_INVALID, "invalid console size %ux%u", x, y);
    if ((uint64_t)(x - 1u) * y >= (1ull << 31)) return fail(RTC_ERR_CAPACITY, "console size %ux%u too large", x, y);
    c->x = x; c->y = y;
    return RTC_OK;
}

// ---- scene ----------------------------------------------------------------------------------
int rtc_scene_clear(rtc_ctx* c)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    c->objs.clear();
    c->scene_dirty = true; c->host_stale = false;
    return RTC_OK;
}

int rtc_scene_add_sphere(rtc_ctx* c, const float center[3], float radius, const float rgb[3], float speed, int mover)
{
    if (!c || !center || !rgb) return fail(RTC_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    int rc = refresh_host_scene(c);
    if (rc) return rc;
    rtc_object o;
    memset(&o, 0, sizeof o);
    o.type = RTC_OBJ_SPHERE;
    for (int i = 0; i < 3; ++i) { o.center[i] = center[i]; o.color[i] = rgb[i]; }
    o.radius = radius; o.speed = speed; o.mover = mover;
    c->objs.push_back(o);
    c->scene_dirty = true;
    return RTC_OK;
}

int rtc_scene_add_plane(rtc_ctx* c, const float center[3], const float normal[3], const float rgb[3], float width, float height)
{
    if (!c || !center || !normal || !rgb) return fail(RTC_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    int rc = refresh_host_scene(c);
    if (rc) return rc;
    rtc_object o;
    memset(&o, 0, sizeof o);
    o.type = RTC_OBJ_PLANE;
    for (int i = 0; i < 3; ++i) { o.center[i] = center[i]; o.color[i] = rgb[i]; }
    // Vector3::Normalize with its zero check (Plane.cu:9, MyMath.h:117-123).
    const float len = sqrtf(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
    const float div = len < 0.000001f ? 0.0f : 1.0f / len;
    for (int i = 0; i < 3; ++i) o.normal[i] = normal[i] * div;
    o.width = width; o.height = height;
    c->objs.push_back(o);
    c->scene_dirty = true;
    return RTC_OK;
}

int rtc_scene_set_objects(rtc_ctx* c, const rtc_object* objs, uint32_t n)
{
    if (!c || (n && !objs)) return fail(RTC_ERR_INVALID, "NULL argument");
    for (uint32_t i = 0; i < n; ++i)
        if (objs[i].type != RTC_OBJ_SPHERE && objs[i].type != RTC_OBJ_PLANE)
            return fail(RTC_ERR_INVALID, "object %u has unknown type %d", i, objs[i].type);
    c->objs.assign(objs, objs + n);
    c->scene_dirty = true; c->host_stale = false;
    return RTC_OK;
}

int rtc_scene_get_objects(rtc_ctx* c, rtc_object* out, uint32_t cap, uint32_t* n)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    int rc = refresh_host_scene(c);
    if (rc) return rc;
    if (n) *n = (uint32_t)c->objs.size();
    if (out) memcpy(out, c->objs.data(), sizeof(rtc_object) * (c->objs.size() < cap ? c->objs.size() : cap));
    return RTC_OK;
}

int rtc_scene_count(rtc_ctx* c, uint32_t* n)
{
    if (!c || !n) return fail(RTC_ERR_INVALID, "NULL argument");
    *n = (uint32_t)c->objs.size();
    return RTC_OK;
}

int rtc_update_objects(rtc_ctx* c, double dt, uint32_t flags)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    const size_t n = c->objs.size();
    // The reference launches UpdateObjects with block = count threads: invalid for count > 1024
    // (and for 0), so nothing moves (RayTracingManager.cu:89-107).
    if ((flags & RTC_FLAG_UPDATE_REF_LAUNCH_LIMIT) && (n > 1024 || n == 0)) return RTC_OK;
    int rc = upload_scene(c);
    if (rc) return rc;
    CK(rtc::launch_update_objects(c->stream, c->d_objs.p, (int)n, dt));
    c->host_stale = true;
    return mark_scene_read(c);
}

// ---- frame ------------------------------------------------------------------------------------
int rtc_render(rtc_ctx* c, const rtc_params* p, rtc_mode mode, uint32_t flags)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (!p) return fail(RTC_ERR_INVALID, "params is NULL");
    CK(cudaSetDevice(c->device));
    if (p->x < 1 || p->y < 1) return fail(RTC_ERR_INVALID, "invalid console size %ux%u", p->x, p->y);
    if ((uint64_t)(p->x - 1u) * p->y >= (1ull << 31)) return fail(RTC_ERR_CAPACITY, "console size too large");
    if (mode < RTC_BIT_ASCII || mode > RTC_SDL) return fail(RTC_ERR_INVALID, "invalid rendering mode %d", (int)mode);
    const uint32_t W = p->x - 1u;
    const size_t n_px = (size_t)W * p->y;
    const size_t cap = rtc_encode_capacity(p->x, p->y, mode);
    CK(c->d_color.ensure(n_px * mode_bpp(mode) + 16));
    if (mode_has_glyph(mode)) CK(c->d_glyph.ensure(n_px + 16));
    const int slot = c->cur ^ 1;                               // the other slot may still be draining to the host
    CK(c->d_out[slot].ensure(cap));
    c->have_frame = false;
    int rc = trace_shade(c, p, mode, flags, 0, p->y, c->d_color.p, mode_has_glyph(mode) ? c->d_glyph.p : nullptr, true);
    if (rc) return rc;
    rc = do_encode(c, c->d_color.p, mode_has_glyph(mode) ? c->d_glyph.p : nullptr, p->x, p->y, mode, c->d_out[slot].p, cap,
                   c->h_total.p + slot, false);        // the emit kernel writes the stream length straight into pinned host memory
    if (rc) return rc;
    c->last_launches += (mode == RTC_SDL || p->x <= 1u) ? 1u : 2u;   // newline kernel, or count + emit
    CK(cudaEventRecord(c->ev[4], c->stream));
    CK(cudaEventRecord(c->ev_total[slot], c->stream));
    c->cur = slot;
    c->slot_cap[slot] = cap;
    c->have_frame = true;
    c->timings_valid = true;
    c->last_mode = mode; c->last_x = p->x; c->last_y = p->y;
    return RTC_OK;
}

int rtc_frame_ansi_device(rtc_ctx* c, const char** dev_ptr, size_t* n_bytes)
{
    if (!c || !dev_ptr || !n_bytes) return fail(RTC_ERR_INVALID, "NULL argument");
    if (!c->have_frame) return fail(RTC_ERR_INVALID, "no frame rendered");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)c->h_total.p[c->cur];
    if (n > c->slot_cap[c->cur]) return fail(RTC_ERR_CAPACITY, "stream (%zu B) exceeds capacity (%zu B)", n, c->slot_cap[c->cur]);
    *dev_ptr = c->d_out[c->cur].p; *n_bytes = n;
    return RTC_OK;
}

int rtc_frame_ansi(rtc_ctx* c, const char** host_ptr, size_t* n_bytes)
{
    const char* dptr = nullptr;
    size_t n = 0;
    int rc = rtc_frame_ansi_device(c, &dptr, &n);
    if (rc) return rc;
    PinBuf<char>& h = c->h_out[c->cur];
    if (n > h.cap) CK(h.ensure(n + n / 2 + 4096));
    if (n) CK(cudaMemcpyAsync(h.p, dptr, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *host_ptr = h.p; *n_bytes = n;
    return RTC_OK;
}

// ---- pipelined frames: submit frame k+1, then collect frame k ------------------------------------------------
int rtc_submit(rtc_ctx* c, const rtc_params* p, rtc_mode mode, double dt, uint32_t flags)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (c->fifo_n >= 2) return fail(RTC_ERR_INVALID, "two frames are already in flight: rtc_collect one first");
    int rc = rtc_update_objects(c, dt, flags);
    if (rc) return rc;
    rc = rtc_render(c, p, mode, flags);
    if (rc) return rc;
    c->fifo[c->fifo_n++] = c->cur;
    return RTC_OK;
}

int rtc_collect(rtc_ctx* c, const char** host_ptr, size_t* n_bytes)
{
    if (!c || !host_ptr || !n_bytes) return fail(RTC_ERR_INVALID, "NULL argument");
    if (c->fifo_n == 0) return fail(RTC_ERR_INVALID, "no frame in flight");
    CK(cudaSetDevice(c->device));
    const int slot = c->fifo[0];
    c->fifo[0] = c->fifo[1];
    --c->fifo_n;
    CK(cudaEventSynchronize(c->ev_total[slot]));               // this frame's kernels are done; the next frame's keep running
    const size_t n = (size_t)c->h_total.p[slot];
    if (n > c->slot_cap[slot]) return fail(RTC_ERR_CAPACITY, "stream (%zu B) exceeds capacity (%zu B)", n, c->slot_cap[slot]);
    PinBuf<char>& h = c->h_out[slot];
    if (n > h.cap) CK(h.ensure(n + n / 2 + 4096));
    if (n) CK(cudaMemcpyAsync(h.p, c->d_out[slot].p, n, cudaMemcpyDeviceToHost, c->copy_stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    *host_ptr = h.p; *n_bytes = n;
    return RTC_OK;
}

int rtc_frame_color(rtc_ctx* c, const uint8_t** host_color, uint32_t* bpp, const uint8_t** host_glyph)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (!c->have_frame) return fail(RTC_ERR_INVALID, "no frame rendered");
    CK(cudaSetDevice(c->device));
    const size_t n_px = (size_t)(c->last_x - 1u) * c->last_y;
    const uint32_t b = mode_bpp(c->last_mode);
    CK(c->h_color.ensure(n_px * b + 1));
    if (n_px && c->last_mode != RTC_SDL)
        CK(cudaMemcpyAsync(c->h_color.p, c->d_color.p, n_px * b, cudaMemcpyDeviceToHost, c->stream));
    const bool gl = mode_has_glyph(c->last_mode);
    if (gl) {
        CK(c->h_glyph.ensure(n_px + 1));
        if (n_px) CK(cudaMemcpyAsync(c->h_glyph.p, c->d_glyph.p, n_px, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    if (host_color) *host_color = c->h_color.p;
    if (bpp) *bpp = b;
    if (host_glyph) *host_glyph = gl ? c->h_glyph.p : nullptr;
    return RTC_OK;
}

int rtc_frame_hits(rtc_ctx* c, const float** host_dist, const int32_t** host_index)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (!c->have_frame) return fail(RTC_ERR_INVALID, "no frame rendered");
    if (!c->hits_valid) return fail(RTC_ERR_INVALID, "the last frame kept no hit records: render it with RTC_FLAG_KEEP_HITS");
    CK(cudaSetDevice(c->device));
    const size_t n_px = (size_t)(c->last_x - 1u) * c->last_y;
    CK(c->h_hit_t.ensure(n_px + 1));
    CK(c->h_hit_idx.ensure(n_px + 1));
    if (n_px) {
        CK(cudaMemcpyAsync(c->h_hit_t.p, c->d_hit_t.p, n_px * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h_hit_idx.p, c->d_hit_idx.p, n_px * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    if (host_dist) *host_dist = c->h_hit_t.p;
    if (host_index) *host_index = c->h_hit_idx.p;
    return RTC_OK;
}

int rtc_update(rtc_ctx* c, const rtc_params* p, rtc_mode mode, double dt, uint32_t flags,
               const char** host_ptr, size_t* n_bytes)
{
    // RayTracingManager::Update runs the physics step unconditionally, even for dt == 0 (where it
    // still clamps every sphere's y into [-10,10], Sphere.cu:18-22) -- so does this.
    int rc = rtc_update_objects(c, dt, flags);
    if (rc) return rc;
    rc = rtc_render(c, p, mode, flags);
    if (rc) return rc;
    return rtc_frame_ansi(c, host_ptr, n_bytes);
}

int rtc_last_timings(rtc_ctx* c, rtc_timings* out)
{
    if (!c || !out) return fail(RTC_ERR_INVALID, "NULL argument");
    if (!c->timings_valid) return fail(RTC_ERR_INVALID, "no frame rendered");
    CK(cudaSetDevice(c->device));
    CK(cudaEventSynchronize(c->ev[4]));
    CK(cudaEventElapsedTime(&out->prep_ms, c->ev[0], c->ev[1]));
    CK(cudaEventElapsedTime(&out->trace_ms, c->ev[1], c->ev[2]));
    CK(cudaEventElapsedTime(&out->shade_ms, c->ev[2], c->ev[3]));
    CK(cudaEventElapsedTime(&out->encode_ms, c->ev[3], c->ev[4]));
    CK(cudaEventElapsedTime(&out->total_ms, c->ev[0], c->ev[4]));
    out->launches = c->last_launches;
    // packed ray-sphere tests the last frame executed: warps x groups x (4 spheres x 256 rays), both passes
    unsigned long long groups[2] = {0ull, 0ull};
    CK(cudaMemcpyAsync(groups, c->d_counters.p + rtc::kStatsCounter + 2 * c->stats_parity, sizeof groups, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out->sphere_tests = (groups[0] + groups[1]) * 4ull * 32ull;        // the kernel counts groups x rays per thread
    return RTC_OK;
}

// ---- stage-level entry points -------------------------------------------------------------------
int rtc_trace_band(rtc_ctx* c, const rtc_params* p, rtc_mode mode, uint32_t flags, uint32_t row0, uint32_t row1,
                   uint8_t* dev_color, uint8_t* dev_glyph)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (!dev_color && mode != RTC_SDL) return fail(RTC_ERR_INVALID, "dev_color is NULL");
    if (mode_has_glyph(mode) && !dev_glyph) return fail(RTC_ERR_INVALID, "dev_glyph is NULL in an ASCII mode");
    CK(cudaSetDevice(c->device));
    return trace_shade(c, p, mode, flags, row0, row1, dev_color, dev_glyph, false);
}

int rtc_trace_raw(rtc_ctx* c, const rtc_params* p, rtc_mode mode, uint32_t flags, char* dev_result)
{
    if (!c || !p || !dev_result) return fail(RTC_ERR_INVALID, "NULL argument");
    if (mode < RTC_BIT_ASCII || mode > RTC_SDL) return fail(RTC_ERR_INVALID, "invalid rendering mode %d", (int)mode);   // (the reference asserts)
    if (p->x < 1 || p->y < 1) return fail(RTC_ERR_INVALID, "invalid console size %ux%u", p->x, p->y);
    if ((uint64_t)(p->x - 1u) * p->y >= (1ull << 31)) return fail(RTC_ERR_CAPACITY, "console size too large");
    CK(cudaSetDevice(c->device));
    const size_t n_px = (size_t)(p->x - 1u) * p->y;
    CK(c->d_color.ensure(n_px * mode_bpp(mode) + 16));
    if (mode_has_glyph(mode)) CK(c->d_glyph.ensure(n_px + 16));
    c->have_frame = false;
    int rc = trace_shade(c, p, mode, flags, 0, p->y, c->d_color.p, mode_has_glyph(mode) ? c->d_glyph.p : nullptr, false);
    if (rc) return rc;
    CK(rtc::launch_expand_raw(c->stream, c->d_color.p, mode_has_glyph(mode) ? c->d_glyph.p : nullptr, p->x, p->y, mode, dev_result));
    return RTC_OK;
}

size_t rtc_raw_size(uint32_t x, uint32_t y) { return (size_t)20 * x * y; }

int rtc_encode(rtc_ctx* c, const uint8_t* dev_color, const uint8_t* dev_glyph, uint32_t x, uint32_t y, rtc_mode mode,
               char* dev_out, size_t cap, unsigned long long* dev_total)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (!dev_out || !dev_total) return fail(RTC_ERR_INVALID, "NULL output");
    if (x < 1 || y < 1) return fail(RTC_ERR_INVALID, "invalid console size %ux%u", x, y);
    if (mode < RTC_BIT_ASCII || mode > RTC_SDL) return fail(RTC_ERR_INVALID, "invalid rendering mode %d", mode);
    if (mode != RTC_SDL && x > 1 && !dev_color) return fail(RTC_ERR_INVALID, "dev_color is NULL");
    if ((uint64_t)(x - 1u) * y >= (1ull << 31)) return fail(RTC_ERR_CAPACITY, "console size too large");
    CK(cudaSetDevice(c->device));
    return do_encode(c, dev_color, dev_glyph, x, y, mode, dev_out, cap, dev_total, false);
}

int rtc_encode_band(rtc_ctx* c, const uint8_t* dev_color, const uint8_t* dev_glyph, uint32_t x, uint32_t rows, rtc_mode mode,
                    int continues, char* dev_out, size_t cap, unsigned long long* dev_total)
{
    if (!c) return fail(RTC_ERR_INVALID, "ctx is NULL");
    if (!dev_out || !dev_total) return fail(RTC_ERR_INVALID, "NULL output");
    if (x < 1) return fail(RTC_ERR_INVALID, "invalid console width %u", x);
    if (mode < RTC_BIT_ASCII || mode > RTC_SDL) return fail(RTC_ERR_INVALID, "invalid rendering mode %d", mode);
    if (mode != RTC_SDL && x > 1 && rows > 0 && !dev_color) return fail(RTC_ERR_INVALID, "dev_color is NULL");
    if ((uint64_t)(x - 1u) * rows >= (1ull << 31)) return fail(RTC_ERR_CAPACITY, "band too large");
    CK(cudaSetDevice(c->device));
    if (rows == 0) {                                            // an empty band contributes an empty stream
        CK(cudaMemsetAsync(dev_total, 0, sizeof(unsigned long long), c->stream));
        return RTC_OK;
    }
    return do_encode(c, dev_color, dev_glyph, x, rows, mode, dev_out, cap, dev_total, continues != 0);
}

int rtc_debug_ansi256_cube(rtc_ctx* c, uint8_t* dev_out)
{
    if (!c || !dev_out) return fail(RTC_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(rtc::launch_ansi256_cube(c->stream, dev_out));
    return RTC_OK;
}

int rtc_camera_params(uint32_t x, uint32_t y, const float pos[3], const float rot[3], float pixel_aspect, rtc_params* out)
{
    const int rc = rtc::camera_params(x, y, pos, rot, pixel_aspect, out);
    if (rc) return fail(rc, "camera_params: invalid argument or singular view matrix");
    return RTC_OK;
}

int rtc_fp32_peak(rtc_ctx* c, int variant, int iters, float* tflops, float* ms)
{
    if (!c || !tflops) return fail(RTC_ERR_INVALID, "NULL argument");
    CK(cudaSetDevice(c->device));
    const int n_ctas = c->sm_count * 4;
    CK(rtc::launch_fp32_peak(c->stream, variant, n_ctas, iters > 8 ? 8 : iters, c->d_sink.p));   // warm-up
    CK(cudaEventRecord(c->ev[5], c->stream));
    CK(rtc::launch_fp32_peak(c->stream, variant, n_ctas, iters, c->d_sink.p));
    CK(cudaEventRecord(c->ev[4], c->stream));                   // (ev[4] is re-recorded by the next rtc_render)
    CK(cudaEventSynchronize(c->ev[4]));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, c->ev[5], c->ev[4]));
    c->timings_valid = false;
    if (ms) *ms = t;
    *tflops = (float)(rtc::fp32_peak_flops(variant, n_ctas, iters) / (t * 1e-3) / 1e12);
    return RTC_OK;
}

}  // extern "C"
