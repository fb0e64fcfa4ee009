// rtc_ctx.h -- internals of librtc_b200 shared by the single-GPU C-ABI (rtc_api.cu) and the multi-GPU frame driver
// (rtc_mgpu.cu): the context object and the two frame stages they both enqueue.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <vector>

#include "rtc_device.cuh"
#include "rtc_kernels.h"
#include "rtc_shade.cuh"

namespace rtc {

int fail(int code, const char* fmt, ...);           // sets the calling thread's rtc_last_error() text, returns code

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return rtc::fail(RTC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline bool mode_is_8bit(int m) { return m == RTC_BIT_ASCII || m == RTC_BIT_PIXEL; }
inline bool mode_has_glyph(int m) { return m == RTC_BIT_ASCII || m == RTC_RGB_ASCII; }
inline uint32_t mode_bpp(int m) { return mode_is_8bit(m) ? 1u : 3u; }
inline uint32_t mode_cell(int m) { return mode_is_8bit(m) ? 12u : 20u; }

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

}  // namespace rtc

struct rtc_ctx {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t x = 0, y = 0;

    // scene (host master copy + device mirror)
    std::vector<rtc_object> objs;
    std::vector<int32_t> sphere_obj, plane_obj;
    std::vector<int32_t> sphere_order, plane_order;   // the index lists of the last upload ...
    std::vector<float> order_key;                     // ... and the geometry they were built from (centre, type per object)
    bool scene_dirty = true;       // host -> device upload pending
    bool host_stale = false;       // device physics ran; host copy must be refreshed before use
    // device scene: ONE blob [objects | sphere index list | plane index list] so that an upload is a single copy
    // (a ring: the upload of the NEXT scene runs on its own stream, ahead of and beside the kernels of the frames in flight,
    //  so a per-frame scene upload is never on a frame's critical path -- in-stream it waited behind the D2H traffic of
    //  the previous frame's stream on the same PCIe link)
    static constexpr int kSceneRing = 4;
    rtc::DevBuf<unsigned char> d_scene[kSceneRing];
    int scene_cur = 0;                               // ring slot the views below point into
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t ev_scene_up[kSceneRing] = {};        // the upload into ring slot i has finished (upload stream)
    cudaEvent_t ev_scene_rd[kSceneRing] = {};        // everything enqueued so far that reads ring slot i has finished (frame stream)
    bool scene_up_pending[kSceneRing] = {}, scene_rd_recorded[kSceneRing] = {};
    struct View { rtc_object* p = nullptr; } d_objs;
    struct ViewI { int32_t* p = nullptr; } d_sphere_obj, d_plane_obj;
    struct ViewK { float4* p = nullptr; } d_kd;     // per object: colour / 255 (RayTracing.cu:144), computed on the host at upload
    rtc::DevBuf<uint8_t> d_shadow;  // 1 byte per pixel: occluded

    // frame buffers
    rtc::DevBuf<float> d_hit_t;
    rtc::DevBuf<int32_t> d_hit_idx;
    rtc::DevBuf<uint8_t> d_color, d_glyph;
    rtc::DevBuf<char> d_out[2];                  // two frame slots: the stream of frame k is copied out while k+1 is encoded
    rtc::DevBuf<unsigned char> d_desc;           // encoder scratch: per-tile counts, group accumulators (two parities)
    uint32_t enc_parity = 0;
    rtc::DevBuf<unsigned long long> d_counters;  // rtc_kernels.h: tile tickets (never reset) + two parities of test counts
    unsigned long long ticket_base[2 * rtc::kMaxChunks] = {};   // tickets each counter has handed out so far
    uint32_t stats_parity = 0;                   // parity of the last frame's test counts
    rtc::DevBuf<unsigned long long> d_total;     // [2]
    rtc::DevBuf<float> d_sink;
    rtc::PinBuf<unsigned long long> h_total;     // [2]
    rtc::PinBuf<char> h_out[2];
    rtc::PinBuf<unsigned char> h_scene[kSceneRing];   // pinned staging of the scene upload (objects + index lists), one per ring slot
    cudaStream_t copy_stream = nullptr;     // D2H of finished streams, concurrent with the next frame's kernels
    cudaEvent_t ev_total[2] = {nullptr, nullptr};   // slot's encode finished and its length is on the host
    int cur = 0;                            // slot of the last rtc_render
    int fifo[2] = {0, 0}, fifo_n = 0;       // submitted, not yet collected slots (oldest first)
    size_t slot_cap[2] = {0, 0};
    rtc::PinBuf<uint8_t> h_color, h_glyph;
    rtc::PinBuf<float> h_hit_t;
    rtc::PinBuf<int32_t> h_hit_idx;

    // light / material block of the shading stage (rtc_set_light; defaults = the reference's constants)
    rtc::ShadeParams shade = {{1.0f, 50.0f, 0.0f}, 1.0f, 2000.0f, 1.0f, 3000.0f, {0.2f, 0.2f, 0.2f}, 1.0f};

    // last frame
    bool have_frame = false;
    bool hits_valid = false;                // the last rtc_render left hit records behind (RTC_FLAG_KEEP_HITS, shadows)
    int last_mode = RTC_RGB_PIXEL;
    uint32_t last_x = 0, last_y = 0;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool timings_valid = false;
    uint32_t last_launches = 0;
};

namespace rtc {
// trace (+ hoist, + shade) for rows [row0,row1) into colour / glyph planes (band-relative), on c->stream.
int trace_shade(rtc_ctx* c, const rtc_params* p, int mode, uint32_t flags, uint32_t row0, uint32_t row1,
                uint8_t* d_color, uint8_t* d_glyph, bool record_events);
// The ANSI encoder as one step (scratch, parity, the two launches), on c->stream.
int do_encode(rtc_ctx* c, const uint8_t* d_color, const uint8_t* d_glyph, uint32_t x, uint32_t rows, int mode, char* d_out,
              size_t cap, unsigned long long* d_total, bool continues);
}  // namespace rtc
