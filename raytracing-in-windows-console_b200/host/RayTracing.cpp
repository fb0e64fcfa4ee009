#include "RayTracing.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "../../include/rtc.h"
#include "Scene3D.h"

namespace {
void gpuAssert(int rc, const char* file, int line)            // the reference's convention, pch.h:45-53
{
    if (rc == RTC_OK) return;
    if (getenv("RTC_FACADE_THROW")) throw std::runtime_error(rtc_last_error());
    fprintf(stderr, "GPUassert: %s %s %d\n", rtc_last_error(), file, line);
    exit(rc);
}
#define gpuErrchk(ans) gpuAssert((ans), __FILE__, __LINE__)
}  // namespace

void RayTracing::RayTrace(const dim3&, const dim3&, Object3D* DEVICE_MEMORY_PTR const objects, const unsigned int,
                          const RayTracingCPUToGPUData* params, char* resultArray, const RenderingMode mode)
{
    SceneBackend* be = reinterpret_cast<SceneBackend*>(objects);
    if (!be || !be->ctx) { fprintf(stderr, "RayTracing::RayTrace needs the single-GPU backend (RTC_GPUS unset)\n"); exit(1); }
    assert(mode >= BIT_ASCII && mode <= SDL);                  // reference RayTracing.cu:862-864
    rtc_params p{};
    const MyMath::Vector4* rows[4] = {&params->inverseVMatrix.row1, &params->inverseVMatrix.row2,
                                      &params->inverseVMatrix.row3, &params->inverseVMatrix.row4};
    for (int r = 0; r < 4; ++r) {
        p.inv_view[4 * r + 0] = rows[r]->x; p.inv_view[4 * r + 1] = rows[r]->y;
        p.inv_view[4 * r + 2] = rows[r]->z; p.inv_view[4 * r + 3] = rows[r]->w;
    }
    p.cam_pos[0] = params->camPos.x; p.cam_pos[1] = params->camPos.y; p.cam_pos[2] = params->camPos.z;
    p.x = (uint32_t)params->x; p.y = (uint32_t)params->y;
    p.element1 = params->element1; p.element2 = params->element2; p.cam_far = params->camFarDist;
    gpuErrchk(rtc_trace_raw(be->ctx, &p, (rtc_mode)mode, RTC_FLAG_CULL | RTC_FLAG_PACKET, resultArray));
}

void RayTracing::Synchronize(Object3D* DEVICE_MEMORY_PTR const objects)
{
    SceneBackend* be = reinterpret_cast<SceneBackend*>(objects);
    if (be && be->ctx) gpuErrchk(rtc_synchronize(be->ctx));
}
