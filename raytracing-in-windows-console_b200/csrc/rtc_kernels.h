// rtc_kernels.h -- host-side launchers of the sm_100a kernels (internal to librtc_b200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtc.h"

namespace rtc {

struct FrameParams;
struct ShadeParams;

// The persistent ray-kernel CTA (one per SM) exists with 24 and with 28 warps; plan_trace picks per launch.
struct TracePlan { int threads; int max_slots; int rays; };   // threads per CTA; sphere slots resident in shared memory per launch; rays per thread (8 or 4)
TracePlan plan_trace(uint32_t x, uint32_t rows, int n_slots, int n_ctas, bool packet);
constexpr int kMaxChunks = 28;             // sphere-list chunks per pass (one ticket counter each): 69k spheres at 24 warps
// Device counters of a context, all 64-bit, zeroed once at creation and never reset: [0 .. kMaxChunks) tile tickets of the
// primary pass (one per sphere chunk), [kMaxChunks .. 2 kMaxChunks) of the shadow pass -- a launch draws its tickets above
// the base the host keeps (trace_tickets) --, then two frame parities x (primary, shadow) counts of sphere groups tested.
constexpr int kStatsCounter = 2 * kMaxChunks;
constexpr int kNumCounters = kStatsCounter + 4;

// kernel 1 (rtc_trace.cu): per-frame scene hoist (every CTA, into its shared memory) + ray generation + nearest hit
// (+ shade/quantise epilogue)
cudaError_t configure_trace();
size_t trace_smem_bytes(int n_slots, int threads, int rays);
unsigned long long trace_tickets(uint32_t x, uint32_t rows, int n_ctas, int threads, int rays);
cudaError_t launch_trace(cudaStream_t st, int n_ctas, const FrameParams& fp, const int32_t* sphere_obj, int n_spheres,
                         int n_slots, const rtc_object* objs, const int32_t* plane_obj, int n_planes, float* hit_t,
                         int32_t* hit_idx, unsigned long long* tile_counter, unsigned long long ticket_base, int carry_in,
                         const float* light /* NULL: primary rays */, uint8_t* shadow, int threads, bool cull,
                         unsigned long long* groups_tested, unsigned long long* stats_zero,
                         const ShadeParams& sp, int shade_mode /* >= 0: shade + quantise in the tile epilogue; -1: no */,
                         uint8_t* color, uint8_t* glyph, bool write_hits, const float4* obj_kd /* per object colour / 255 */,
                         bool affine /* screen-affine packed filter (primary rays only) */, int rays /* per thread: 8 or 4 */,
                         bool packet /* RTC_FLAG_PACKET: the filter on the end rays of every thread's packet (needs affine) */);

// kernel 2 (rtc_shade.cu): stand-alone shade + quantise, used only after a shadow pass
cudaError_t launch_shade(cudaStream_t st, const FrameParams& fp, const ShadeParams& sp, int mode,
                         const rtc_object* objs, const float4* obj_kd, const float* hit_t, const int32_t* hit_idx,
                         const uint8_t* shadow /* NULL: every point is lit */, uint8_t* color, uint8_t* glyph);

cudaError_t launch_ansi256_cube(cudaStream_t st, uint8_t* out /* 2^24 bytes */);

// kernel 3 (rtc_encode.cu)
cudaError_t configure_encode();
size_t encode_state_bytes(uint64_t n_cells);
// `scratch`: >= encode_state_bytes() of device memory, ZEROED when allocated; `parity` must alternate between
// consecutive launches on the same scratch (each launch zeroes the other parity's accumulators).  Two launches.
cudaError_t launch_encode(cudaStream_t st, const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y,
                          int mode, char* out, size_t cap, unsigned long long* total, void* scratch, size_t scratch_bytes,
                          uint32_t parity, bool continues = false /* the cell stored before `color` precedes cell 0 */);

// planes -> the reference's raw 20*x*y-byte cell buffer (compatibility / parity hook; rtc_encode.cu)
cudaError_t launch_expand_raw(cudaStream_t st, const uint8_t* color, const uint8_t* glyph, uint32_t x, uint32_t y, int mode, char* out);

// physics (rtc_shade.cu)
cudaError_t launch_update_objects(cudaStream_t st, rtc_object* objs, int n, double dt);

// microbenchmarks (rtc_microbench.cu)
cudaError_t launch_fp32_peak(cudaStream_t st, int variant, int n_ctas, int iters, float* sink);

}  // namespace rtc

namespace rtc {
double fp32_peak_flops(int variant, int n_ctas, int iters);
// Camera3D math on the host (rtc_camera.cpp)
int camera_params(uint32_t x, uint32_t y, const float pos[3], const float rot[3], float pixel_aspect, rtc_params* out);
}  // namespace rtc
