// Engine3D.h -- app loop, API as the reference's Engine3D (reference Engine3D.h/.cpp), headless:
// the Win32 keyboard/mouse polling is replaced by an injectable input hook.
#pragma once
#include <functional>
#include <memory>

#include "Camera3D.h"
#include "RayTracingManager.h"
#include "Scene3D.h"
#include "Timer.h"

class Engine3D
{
public:
    Engine3D() = default;
    ~Engine3D() = default;

    void Start();                         // PrintMachine::Start(400,150), manager, camera, scene (reference Engine3D.cpp:6-28)
    void Start(size_t x, size_t y);       // extension: explicit console size
    bool Run();                           // one frame (reference Engine3D.cpp:30-79)
    void CleanUp();

    // Extensions for headless drivers.
    Camera3D& Camera() { return *m_camera; }
    Scene3D& Scene() { return *m_scene; }
    RayTracingManager& Manager() { return *m_rayTracingManager; }
    void SetInputHook(std::function<void(Engine3D&, long double)> f) { m_input = std::move(f); }
    void SetFixedDt(long double dt) { m_fixedDt = dt; }
    void Quit() { m_bShouldQuit = true; }

private:
    void Render(const long double dt);
    void CheckKeyboard(const long double dt);

    std::unique_ptr<Time> m_timer;
    std::unique_ptr<Camera3D> m_camera;
    std::unique_ptr<Scene3D> m_scene;
    std::unique_ptr<RayTracingManager> m_rayTracingManager;
    std::function<void(Engine3D&, long double)> m_input;
    long double m_fpsTimer = 0.0;
    long double m_fixedDt = -1.0;
    int m_fps = 0;
    bool m_bShouldQuit = false;
};
