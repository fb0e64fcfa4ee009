// facade_test.cpp -- drives the C++ facade exactly the way the reference's Entrypoint/Engine3D do
// and dumps what lands in PrintMachine's back buffer, for the parity tests (tests/test_facade.py).
//   facade_test <outdir>           -> default scene, 240x64, every mode  -> <outdir>/default_240x64_m<k>.bin
//   facade_test <outdir> engine N  -> Engine3D::Start(240,64) + N frames (dt = 0) in RGB_PIXEL
//   facade_test <outdir> pipelined N -> the same through the pipelined sink (SetPipelined + Flush), culling on
#include <cstdio>
#include <cstring>
#include <string>

#include "Engine3D.h"
#include "PrintMachine.h"

static void dump(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    fwrite(PrintMachine::GetBackBuffer(), 1, PrintMachine::GetPrintSize(), f);
    fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: facade_test <outdir> [engine N]\n"); return 2; }
    const std::string out = argv[1];
    if (argc >= 4 && (!strcmp(argv[2], "engine") || !strcmp(argv[2], "pipelined"))) {
        const bool pipelined = !strcmp(argv[2], "pipelined");
        Engine3D engine;
        engine.Start(240, 64);
        engine.SetFixedDt(0.0);
        engine.Manager().SetRenderingMode(RGB_PIXEL);
        if (pipelined) { engine.Manager().SetPipelined(true); engine.Manager().SetCulling(true); }
        const int n = atoi(argv[3]);
        for (int i = 0; i < n && engine.Run(); ++i) {}
        engine.Manager().Flush();
        dump(out + (pipelined ? "/pipelined_240x64_m3.bin" : "/engine_240x64_m3.bin"));
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
    for (int mode = BIT_ASCII; mode <= SDL; ++mode) {
        scene.Init();                                 // fresh default scene (Update moves/clamps the spheres)
        manager.SetRenderingMode((RenderingMode)mode);
        RayTracingCPUToGPUData params;
        params.inverseVMatrix = camera.GetInverseVMatrix();
        params.camPos = camera.GetPos();
        params.x = PrintMachine::GetWidth();
        params.y = PrintMachine::GetHeight();
        params.element1 = camera.GetPMatrix().row1.x;
        params.element2 = camera.GetPMatrix().row2.y;
        params.camFarDist = camera.GetFarPlaneDistance();
        manager.Update(params, scene.GetObjects(), 0.0);
        dump(out + "/default_240x64_m" + std::to_string(mode) + ".bin");
    }
    scene.CleanUp();
    printf("facade_test ok\n");
    return 0;
}
