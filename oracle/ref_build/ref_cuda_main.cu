// TEST INFRASTRUCTURE -- throughput baseline only.  Runs the reference's OWN CUDA kernels (compiled
// unmodified for sm_100 with the vcxproj's flags: -rdc=true --use_fast_math) on a scene file written
// by bench.py, and reports CUDA-event timings of (a) RayTracingManager::Update (memset + kernels +
// sync + full-buffer D2H + host minimise) and (b) the RayTrace_* kernel alone.
//   ref_cuda_sm100 <scene.bin> <frames> [dump.raw]
// With a third argument the raw cell buffer the kernel leaves in m_deviceResultArray (20*x*y bytes, after the
// reference's own per-frame memset) is written to that file: the cross-check of the new path against the reference's
// REAL platform (tests/test_ref_cuda_crosscheck.py).
// scene.bin: u32 n_objs, u32 mode, rtc_params (96 B), rtc_object[n_objs] (64 B each).
#include "pch.h"
#define private public
#define protected public
#include "Scene3D.h"
#include "RayTracingManager.h"
#undef private
#undef protected
#include "RayTracing.h"
#include "PrintMachine.h"
#include "../../include/rtc.h"

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ref_cuda_sm100 scene.bin frames\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    uint32_t n = 0, mode = 3;
    rtc_params p;
    if (fread(&n, 4, 1, f) != 1 || fread(&mode, 4, 1, f) != 1 || fread(&p, sizeof p, 1, f) != 1) return 2;
    std::vector<rtc_object> objs(n);
    if (n && fread(objs.data(), sizeof(rtc_object), n, f) != n) return 2;
    fclose(f);
    const int frames = atoi(argv[2]);

    PrintMachine::Start(p.x, p.y);
    RayTracingManager mgr;
    mgr.SetRenderingMode((RenderingMode)mode);
    Scene3D scene;
    scene.Init();
    scene.m_deviceObjects.count = 0; scene.m_devicePlanes.count = 0; scene.m_deviceSpheres.count = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const rtc_object& o = objs[i];
        MyMath::Vector3 c(o.center[0], o.center[1], o.center[2]), col(o.color[0], o.color[1], o.color[2]);
        if (o.type == RTC_OBJ_SPHERE) scene.CreateSphere(o.radius, c, col);
        else scene.CreatePlane(c, MyMath::Vector3(o.normal[0], o.normal[1], o.normal[2]), col, o.width, o.height);
    }
    RayTracingCPUToGPUData q;
    const float* m = p.inv_view;
    q.inverseVMatrix.row1 = MyMath::Vector4(m[0], m[1], m[2], m[3]);
    q.inverseVMatrix.row2 = MyMath::Vector4(m[4], m[5], m[6], m[7]);
    q.inverseVMatrix.row3 = MyMath::Vector4(m[8], m[9], m[10], m[11]);
    q.inverseVMatrix.row4 = MyMath::Vector4(m[12], m[13], m[14], m[15]);
    q.camPos = MyMath::Vector3(p.cam_pos[0], p.cam_pos[1], p.cam_pos[2]);
    q.x = p.x; q.y = p.y; q.element1 = p.element1; q.element2 = p.element2; q.camFarDist = p.cam_far;

    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    // (b) kernel alone, through the reference's inner seam RayTracing::RayTrace
    DeviceObjectArray<Object3D*> arr = scene.GetObjects();
    dim3 grid((unsigned)std::ceil((p.x + 1) / 16.0), (unsigned)std::ceil(p.y / 16.0), 1), block(16, 16, 1);
    cudaMemcpy(mgr.m_deviceRayTracingData, &q, sizeof q, cudaMemcpyHostToDevice);
    float kernel_ms = 0.f;
    for (int it = 0; it < frames + 1; ++it) {
        cudaEventRecord(a);
        RayTracing::RayTrace(grid, block, arr.m_deviceArray, arr.count, mgr.m_deviceRayTracingData, mgr.m_deviceResultArray, (RenderingMode)mode);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it) kernel_ms += ms;
    }
    cudaError_t e = cudaGetLastError();
    if (argc >= 4) {
        const size_t bytes = (size_t)20 * p.x * p.y;
        std::vector<char> raw(bytes);
        cudaMemset(mgr.m_deviceResultArray, 0, bytes);           // RayTracingManager::ResetDeviceBackBuffer (RayTracingManager.cu:161-165)
        RayTracing::RayTrace(grid, block, arr.m_deviceArray, arr.count, mgr.m_deviceRayTracingData, mgr.m_deviceResultArray, (RenderingMode)mode);
        cudaDeviceSynchronize();
        cudaMemcpy(raw.data(), mgr.m_deviceResultArray, bytes, cudaMemcpyDeviceToHost);
        FILE* o = fopen(argv[3], "wb");
        if (!o) { perror(argv[3]); return 2; }
        fwrite(raw.data(), 1, bytes, o);
        fclose(o);
    }
    // (a) the whole Update; for > 1024 objects its UpdateObjects launch is invalid (block = count) --
    // clear the sticky-less launch error first so the reference's own gpuErrchk does not exit on it.
    double update_ms = 0.0;
    size_t stream_bytes = 0;
    for (int it = 0; it < frames + 1; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        mgr.Update(q, arr, 0.0);
        cudaGetLastError();
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (it) update_ms += ms;
        stream_bytes = PrintMachine::GetPrintSize();
    }
    const double rays = (double)(p.x - 1) * p.y;
    printf("{\"impl\": \"ref_cuda_sm100\", \"objects\": %u, \"x\": %u, \"y\": %u, \"mode\": %u, \"frames\": %d, "
           "\"kernel_ms\": %.4f, \"kernel_mrays_s\": %.2f, \"update_ms\": %.3f, \"update_mrays_s\": %.2f, "
           "\"stream_bytes\": %zu, \"last_cuda_error\": \"%s\"}\n",
           n, p.x, p.y, mode, frames, kernel_ms / frames, rays / (kernel_ms / frames * 1e-3) / 1e6,
           update_ms / frames, rays / (update_ms / frames * 1e-3) / 1e6, stream_bytes, cudaGetErrorString(e));
    return 0;
}
