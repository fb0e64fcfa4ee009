// RayTracingManager.h -- THE DROP-IN BOUNDARY, API as the reference's (reference
// RayTracingManager.h:9-54): per-frame Update(params, objects, dt) that ends in
// PrintMachine::SetDataInBackBuffer(stream, size).
#pragma once
#include <cstddef>

#include "MyMath.h"
#include "Object3D.h"

struct RayTracingCPUToGPUData
{
    MyMath::Matrix inverseVMatrix;
    MyMath::Vector3 camPos;
    size_t x;
    size_t y;
    float element1;
    float element2;
    float camFarDist;
};

enum RenderingMode { BIT_ASCII = 0, BIT_PIXEL, RGB_ASCII, RGB_PIXEL, RGB_NORMALS, SDL };

class RayTracingManager
{
public:
    RayTracingManager();     // PrintMachine::Start must already have run (reference RayTracingManager.cu:58)
    ~RayTracingManager();

    void Update(const RayTracingCPUToGPUData& params, const DeviceObjectArray<Object3D*>& objects, double dt);
    void SetRenderingMode(const RenderingMode newRenderMode);

    // Extensions.  Shadows: opt-in shadow rays (not in the reference).  FixLaunchLimit(true): let
    // UpdateObjects move more than 1024 objects (the reference's launch is rejected there).
    void SetShadows(bool on) { m_shadows = on; }
    // Per-tile sphere culling (RTC_FLAG_CULL): identical frames, far fewer ray-sphere tests.  ON by default -- it is
    // bit-identical by test (tests/test_gpu_parity.py::test_culling_is_invisible).
    void SetCulling(bool on) { m_culling = on; }
    // Packet filter (RTC_FLAG_PACKET): the ray-sphere filter once per 8-ray packet instead of once per ray; identical
    // frames (tests/test_gpu_parity.py::test_packet_filter_is_invisible).  ON by default (RTC_FACADE_NO_PACKET=1: off).
    void SetPacketFilter(bool on) { m_packets = on; }
    // Pipelined sink (rtc_submit / rtc_collect, rtc_mgpu_submit / rtc_mgpu_collect): Update(k) enqueues frame k and hands
    // frame k-1 to PrintMachine, so the copy of k-1 to the host runs under the kernels of k -- one frame of latency, like
    // the reference's own print thread behind SetDataInBackBuffer.  ON by default (RTC_FACADE_SYNC=1 or
    // SetPipelined(false) restore the reference's synchronous hand-over); Flush() delivers the frame still in flight.
    void SetPipelined(bool on);
    void Flush();
    void FixLaunchLimit(bool on) { m_fixLaunchLimit = on; }

private:
    RenderingMode currentRenderingMode = BIT_ASCII;    // reference RayTracingManager.h:53
    bool m_shadows = false;
    bool m_culling = true;
    bool m_packets = true;
    bool m_pipelined = true;
    bool m_inFlight = false;
    void* m_backend = nullptr;
    bool m_fixLaunchLimit = false;
};
