#include "RayTracingManager.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/rtc.h"
#include "PrintMachine.h"
#include "Scene3D.h"

namespace {
// The reference's error convention (pch.h:45-53): print "GPUassert: ..." and exit.  The C-ABI
// returns codes; the facade restores the reference behaviour unless RTC_FACADE_THROW is set.
void gpuAssert(int rc, const char* file, int line)
{
    if (rc == RTC_OK) return;
    if (getenv("RTC_FACADE_THROW")) throw std::runtime_error(rtc_last_error());
    fprintf(stderr, "GPUassert: %s %s %d\n", rtc_last_error(), file, line);
    exit(rc);
}
#define gpuErrchk(ans) gpuAssert((ans), __FILE__, __LINE__)
}  // namespace

RayTracingManager::RayTracingManager()
{
    SceneBackend* b = Scene3D::Backend();
    if (b->ctx) gpuErrchk(rtc_resize(b->ctx, (uint32_t)PrintMachine::GetWidth(), (uint32_t)PrintMachine::GetHeight()));
    if (getenv("RTC_FACADE_SYNC")) m_pipelined = false;
    if (getenv("RTC_FACADE_NO_CULL")) m_culling = false;
    if (getenv("RTC_FACADE_NO_PACKET")) m_packets = false;
}

RayTracingManager::~RayTracingManager() {}   // (a frame still in flight is dropped with the context)

void RayTracingManager::SetRenderingMode(const RenderingMode newRenderMode) { currentRenderingMode = newRenderMode; }

void RayTracingManager::Update(const RayTracingCPUToGPUData& params, const DeviceObjectArray<Object3D*>& objects, double dt)
{
    SceneBackend* be = reinterpret_cast<SceneBackend*>(objects.m_deviceArray);
    rtc_params p{};
    const MyMath::Vector4* rows[4] = {&params.inverseVMatrix.row1, &params.inverseVMatrix.row2,
                                      &params.inverseVMatrix.row3, &params.inverseVMatrix.row4};
    for (int r = 0; r < 4; ++r) {
        p.inv_view[4 * r + 0] = rows[r]->x; p.inv_view[4 * r + 1] = rows[r]->y;
        p.inv_view[4 * r + 2] = rows[r]->z; p.inv_view[4 * r + 3] = rows[r]->w;
    }
    p.cam_pos[0] = params.camPos.x; p.cam_pos[1] = params.camPos.y; p.cam_pos[2] = params.camPos.z;
    p.x = (uint32_t)params.x; p.y = (uint32_t)params.y;
    p.element1 = params.element1; p.element2 = params.element2; p.cam_far = params.camFarDist;
    uint32_t flags = (m_shadows ? RTC_FLAG_SHADOWS : 0u) | (m_culling ? RTC_FLAG_CULL : 0u) | (m_packets ? RTC_FLAG_PACKET : 0u) |
                     (m_fixLaunchLimit ? 0u : RTC_FLAG_UPDATE_REF_LAUNCH_LIMIT);
    const char* stream = nullptr;
    size_t size = 0;
    m_backend = be;
    const rtc_mode mode = (rtc_mode)currentRenderingMode;
    if (m_pipelined) {
        gpuErrchk(be->mgpu ? rtc_mgpu_submit(be->mgpu, &p, mode, dt, flags) : rtc_submit(be->ctx, &p, mode, dt, flags));   // frame k
        if (m_inFlight) {                                                                                                // frame k-1
            gpuErrchk(be->mgpu ? rtc_mgpu_collect(be->mgpu, &stream, &size) : rtc_collect(be->ctx, &stream, &size));
            PrintMachine::SetDataInBackBuffer(stream, size);
        }
        m_inFlight = true;
        return;
    }
    // physics step + trace + shade + ANSI encode + stream to (pinned) host memory
    gpuErrchk(be->mgpu ? rtc_mgpu_update(be->mgpu, &p, mode, dt, flags, &stream, &size) : rtc_update(be->ctx, &p, mode, dt, flags, &stream, &size));
    PrintMachine::SetDataInBackBuffer(stream, size);             // reference RayTracingManager.cu:150
}

void RayTracingManager::SetPipelined(bool on)
{
    if (!on) Flush();
    m_pipelined = on;
}

void RayTracingManager::Flush()
{
    if (!m_inFlight || !m_backend) return;
    const char* stream = nullptr;
    size_t size = 0;
    SceneBackend* be = reinterpret_cast<SceneBackend*>(m_backend);
    gpuErrchk(be->mgpu ? rtc_mgpu_collect(be->mgpu, &stream, &size) : rtc_collect(be->ctx, &stream, &size));
    PrintMachine::SetDataInBackBuffer(stream, size);
    m_inFlight = false;
}
