#include "Engine3D.h"

#include <cstdlib>

#include "PrintMachine.h"

void Engine3D::Start() { Start(400, 150); }     // reference Engine3D.cpp:16

void Engine3D::Start(size_t x, size_t y)
{
    m_timer = std::make_unique<Time>();
    m_camera = std::make_unique<Camera3D>();
    m_scene = std::make_unique<Scene3D>();
    PrintMachine::Start(x, y);                   // must precede the manager (it sizes itself from the sink)
    m_rayTracingManager = std::make_unique<RayTracingManager>();
    m_camera->Init();
    m_camera->Update();
    m_scene->Init();
}

bool Engine3D::Run()
{
    if (!PrintMachine::CheckIfRunning()) { CleanUp(); return false; }
    if (m_bShouldQuit) return false;
    m_timer->Update();
    const long double dt = m_fixedDt >= 0.0 ? m_fixedDt : m_timer->DeltaTime();
    m_fpsTimer += dt;
    m_fps++;
    CheckKeyboard(dt);
    m_camera->Move(dt);
    Render(dt);
    if (m_fpsTimer >= 1.0f) {
        // "Create a sphere every second for testing purposes" (reference Engine3D.cpp:63)
        m_scene->CreateSphere(static_cast<float>(rand() % 10),
                              MyMath::Vector3(rand() % 100 - 50, rand() % 100 - 50, rand() % 100 - 50),
                              MyMath::Vector3(rand() % 255, rand() % 255, rand() % 255));
        PrintMachine::UpdateRenderingFPS(m_fps);
        m_fpsTimer = 0.0f;
        m_fps = 0;
    }
    return true;
}

void Engine3D::Render(const long double dt)     // reference Engine3D.cpp:81-107
{
    m_camera->Update();
    m_scene->Update(dt);
    RayTracingCPUToGPUData params;
    params.inverseVMatrix = m_camera->GetInverseVMatrix();
    params.camPos = m_camera->GetPos();
    params.x = PrintMachine::GetWidth();
    params.y = PrintMachine::GetHeight();
    params.element1 = m_camera->GetPMatrix().row1.x;
    params.element2 = m_camera->GetPMatrix().row2.y;
    params.camFarDist = m_camera->GetFarPlaneDistance();
    DeviceObjectArray<Object3D*> objects = m_scene->GetObjects();
    m_rayTracingManager->Update(params, objects, (double)dt);
}

void Engine3D::CheckKeyboard(const long double dt)
{
    // The reference polls GetKeyState/GetCursorPos here (Engine3D.cpp:110-240); headless builds
    // inject input instead (keys -> m_camera->m_Keys, F1..F5 -> SetRenderingMode, Esc -> Quit()).
    if (m_input) m_input(*this, dt);
}

void Engine3D::CleanUp()
{
    PrintMachine::CleanUp();
    if (m_scene) m_scene->CleanUp();
}
