// facade_test.cpp -- drives the C++ facade exactly the way the reference's Entrypoint/Engine3D do
// and dumps what lands in PrintMachine's back buffer, for the parity tests (tests/test_facade.py).
//   facade_test <outdir>             -> default scene, 240x64, every mode  -> <outdir>/default_240x64_m<k>.bin
//   facade_test <outdir> engine N    -> Engine3D::Start(240,64) + N frames (dt = 0) in RGB_PIXEL, facade defaults
//                                       (pipelined sink + culling); with RTC_GPUS=N the frames come from the multi-GPU driver
//   facade_test <outdir> sync N      -> the same with the reference's synchronous hand-over (SetPipelined(false)), no culling
//   facade_test <outdir> raw         -> RayTracing::RayTrace (the inner seam): raw 20*x*y-byte cell buffers, every mode
//   facade_test <outdir> print N     -> N frames through the print thread into <outdir>/printed.bin
#include <cuda_runtime.h>

#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Engine3D.h"
#include "PrintMachine.h"
#include "RayTracing.h"

static void dump(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    fwrite(PrintMachine::GetBackBuffer(), 1, PrintMachine::GetPrintSize(), f);
    fclose(f);
}

static RayTracingCPUToGPUData make_params(const Camera3D& camera)
{
    RayTracingCPUToGPUData params;
    params.inverseVMatrix = camera.GetInverseVMatrix();
    params.camPos = camera.GetPos();
    params.x = PrintMachine::GetWidth();
    params.y = PrintMachine::GetHeight();
    params.element1 = camera.GetPMatrix().row1.x;
    params.element2 = camera.GetPMatrix().row2.y;
    params.camFarDist = camera.GetFarPlaneDistance();
    return params;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: facade_test <outdir> [engine N | sync N | raw | print N]\n"); return 2; }
    const std::string out = argv[1];
    const std::string what = argc >= 3 ? argv[2] : "";
    if (argc >= 4 && (what == "engine" || what == "sync" || what == "print")) {
        Engine3D engine;
        engine.Start(240, 64);
        engine.SetFixedDt(0.0);
        engine.Manager().SetRenderingMode(RGB_PIXEL);
        if (what == "sync") { engine.Manager().SetPipelined(false); engine.Manager().SetCulling(false); }
        FILE* sink = nullptr;
        if (what == "print") {
            sink = fopen((out + "/printed.bin").c_str(), "wb");
            if (!sink) { perror("printed.bin"); return 2; }
            PrintMachine::StartPrintThread(sink, true);
        }
        const int n = atoi(argv[3]);
        for (int i = 0; i < n && engine.Run(); ++i) {}
        engine.Manager().Flush();
        if (what == "print") {
            for (int spin = 0; spin < 20000 && PrintMachine::FramesPrinted() == 0; ++spin) usleep(100);   // let the thread pick the last frame up
            usleep(20000);
            PrintMachine::JoinPrintThread();
            fclose(sink);
            printf("printed %zu frames\n", PrintMachine::FramesPrinted());
        }
        dump(out + "/" + what + "_240x64_m3.bin");
        engine.CleanUp();
        return 0;
    }
    // The reference's start-up order (Engine3D.cpp:6-28), spelled out.
    PrintMachine::Start(240, 64);
    RayTracingManager manager;
    Camera3D camera;
    camera.Init();
    camera.Update();
    Scene3D scene;
    if (what == "raw") {
        // What the reference's manager does around its inner seam (RayTracingManager.cu:60, :119-143): own a 20*x*y-byte
        // device buffer, launch RayTracing::RayTrace into it, synchronise, copy it to the host.
        const size_t bytes = 20 * PrintMachine::GetWidth() * PrintMachine::GetHeight();
        char* dev = nullptr;
        if (cudaMalloc(&dev, bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 3; }
        std::vector<char> host(bytes);
        for (int mode = BIT_ASCII; mode <= SDL; ++mode) {
            scene.Init();
            const RayTracingCPUToGPUData params = make_params(camera);
            const DeviceObjectArray<Object3D*> objects = scene.GetObjects();
            cudaMemset(dev, 0x5a, bytes);                       // the launcher must define every byte
            RayTracing::RayTrace(dim3(15, 4, 1), dim3(16, 16, 1), objects.m_deviceArray, objects.count, &params, dev, (RenderingMode)mode);
            RayTracing::Synchronize(objects.m_deviceArray);
            if (cudaMemcpy(host.data(), dev, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "cudaMemcpy failed\n"); return 3; }
            FILE* f = fopen((out + "/raw_240x64_m" + std::to_string(mode) + ".bin").c_str(), "wb");
            fwrite(host.data(), 1, bytes, f);
            fclose(f);
        }
        cudaFree(dev);
        scene.CleanUp();
        printf("facade_test raw ok\n");
        return 0;
    }
    manager.SetPipelined(false);                                // frame k must be in the back buffer when Update(k) returns
    for (int mode = BIT_ASCII; mode <= SDL; ++mode) {
        scene.Init();                                 // fresh default scene (Update moves/clamps the spheres)
        manager.SetRenderingMode((RenderingMode)mode);
        const RayTracingCPUToGPUData params = make_params(camera);
        manager.Update(params, scene.GetObjects(), 0.0);
        dump(out + "/default_240x64_m" + std::to_string(mode) + ".bin");
    }
    scene.CleanUp();
    printf("facade_test ok\n");
    return 0;
}
