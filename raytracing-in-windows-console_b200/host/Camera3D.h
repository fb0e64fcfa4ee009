// Camera3D.h -- host camera, API as the reference's Camera3D (reference Camera3D.h/.cpp).
// The matrix math is delegated to the library's host routine (rtc_camera_params, which restates
// Camera3D::Init/Update/GetInverseVMatrix bit for bit), so facade and C-ABI cannot drift apart.
#pragma once
#include "MyMath.h"

struct COORD { short X, Y; };   // stands in for the Win32 type the reference's API mentions

class Camera3D
{
public:
    Camera3D() = default;
    ~Camera3D() = default;

    void Init();
    void Update();
    void SetRot(const float p, const float y, const float r);
    void SetPos(const float x, const float y, const float z);
    void Move(const long double dt);
    void AddRot(const long double dt, const short p, const short y, const short r);

    const MyMath::Matrix& GetVMatrix() const { return m_vMatrix; }
    const MyMath::Matrix GetInverseVMatrix() const;
    const MyMath::Matrix& GetPMatrix() const { return m_pMatrix; }
    const MyMath::Vector3& GetPos() const { return m_pos; }
    const MyMath::Vector3& GetRot() const { return m_rot; }
    const MyMath::Vector3& GetRight() const { return m_right; }
    const MyMath::Vector3& GetUp() const { return m_up; }
    const MyMath::Vector3& GetForward() const { return m_forward; }
    const float GetFarPlaneDistance() const { return m_screenFar; }
    const MyMath::Vector4 GetFrustum() const { return MyMath::Vector4(m_wNear, m_hNear, m_wFar, m_hFar); }
    void SetMouseCoords(const COORD& c) { m_mouseCoords = c; }
    const COORD& GetMouseCoords() { return m_mouseCoords; }

    // Extension: the reference hard-codes 0.01f (Camera3D.cpp:17) and says to retune it per
    // resolution; 0 keeps the reference value.
    void SetPixelAspect(const float k) { m_pixelAspect = k; }

    struct PressedKeys { int W = 0, A = 0, S = 0, D = 0, Space = 0, Shift = 0; };
    PressedKeys m_Keys;

private:
    MyMath::Matrix m_vMatrix, m_pMatrix;
    MyMath::Vector3 m_right = MyMath::Vector3(-1.0f, 0.0f, 0.0f);
    MyMath::Vector3 m_up = MyMath::Vector3(0.0f, 1.0f, 0.0f);
    MyMath::Vector3 m_forward = MyMath::Vector3(0.0f, 0.0f, 1.0f);
    MyMath::Vector3 m_staticRight = MyMath::Vector3(-1.0f, 0.0f, 0.0f);
    MyMath::Vector3 m_staticForward = MyMath::Vector3(0.0f, 0.0f, 1.0f);
    MyMath::Vector3 m_pos;
    MyMath::Vector3 m_rot = MyMath::Vector3(0.0f, 3.14159274101257324f, 0.0f);   // (0, (float)M_PI, 0)
    float m_hNear = 0.0f, m_wNear = 0.0f, m_hFar = 0.0f, m_wFar = 0.0f;
    COORD m_mouseCoords = {-1, -1};
    float m_pixelAspect = 0.0f;
    const float m_screenNear = 0.1f;
    const float m_screenFar = 250.0f;
    const float m_FOV = 1.5f;
};
