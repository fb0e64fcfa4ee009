#include "Camera3D.h"

#include <cmath>

#include "../../include/rtc.h"
#include "PrintMachine.h"

namespace {
// One call computes what Init + Update + GetInverseVMatrix compute in the reference.
rtc_params blockFor(const MyMath::Vector3& pos, const MyMath::Vector3& rot, float pixelAspect)
{
    rtc_params p{};
    const float po[3] = {pos.x, pos.y, pos.z}, ro[3] = {rot.x, rot.y, rot.z};
    rtc_camera_params((uint32_t)PrintMachine::GetWidth(), (uint32_t)PrintMachine::GetHeight(), po, ro, pixelAspect, &p);
    return p;
}
}  // namespace

void Camera3D::Init()     // reference Camera3D.cpp:8-48
{
    const rtc_params p = blockFor(m_pos, m_rot, m_pixelAspect);
    const float fov = 3.14159274101257324f / m_FOV;
    const float width = (float)PrintMachine::GetWidth(), height = (float)PrintMachine::GetHeight();
    const float aspect = width / ((m_pixelAspect == 0.0f ? 0.01f : m_pixelAspect) * width * height);
    m_hNear = 2.0f * std::tan(fov / 2.0f) * m_screenNear; m_wNear = m_hNear * aspect;
    m_hFar = 2.0f * std::tan(fov / 2.0f) * m_screenFar;  m_wFar = m_hFar * aspect;
    m_pMatrix = MyMath::Matrix();
    m_pMatrix.row1.x = p.element1;
    m_pMatrix.row2.y = p.element2;
    m_pMatrix.row3.z = (m_screenFar + m_screenNear) / (m_screenNear - m_screenFar);
    m_pMatrix.row3.w = (2.0f * m_screenFar * m_screenNear) / (m_screenNear - m_screenFar);
    m_pMatrix.row4.z = -1.0f;
}

void Camera3D::Update()   // reference Camera3D.cpp:51-98
{
    const float p = m_rot.x, y = m_rot.y;
    m_forward = MyMath::Vector3(-std::sin(y), -std::sin(p) * std::cos(y), -std::cos(p) * std::cos(y));
    m_staticForward = MyMath::Vector3(-std::sin(y), -std::cos(y), -std::cos(y));
    m_right = MyMath::Vector3(std::cos(y), -std::sin(p) * std::sin(y), -std::cos(p) * std::sin(y));
    m_staticRight = MyMath::Vector3(std::cos(y), -std::sin(y), -std::sin(y));
    m_up = MyMath::Vector3(0.0f, std::cos(p), -std::sin(p));
    m_vMatrix.row1 = MyMath::Vector4(m_right.x, m_up.x, m_forward.x, m_pos.x);
    m_vMatrix.row2 = MyMath::Vector4(m_right.y, m_up.y, m_forward.y, m_pos.y);
    m_vMatrix.row3 = MyMath::Vector4(m_right.z, m_up.z, m_forward.z, m_pos.z);
    m_vMatrix.row4 = MyMath::Vector4(0.0f, 0.0f, 0.0f, 1.0f);
}

const MyMath::Matrix Camera3D::GetInverseVMatrix() const   // reference Camera3D.cpp:207-376
{
    const rtc_params p = blockFor(m_pos, m_rot, m_pixelAspect);
    const float* m = p.inv_view;
    return MyMath::Matrix(MyMath::Vector4(m[0], m[1], m[2], m[3]), MyMath::Vector4(m[4], m[5], m[6], m[7]),
                          MyMath::Vector4(m[8], m[9], m[10], m[11]), MyMath::Vector4(m[12], m[13], m[14], m[15]));
}

void Camera3D::SetRot(const float p, const float y, const float r) { m_rot = MyMath::Vector3(p, y, r); }
void Camera3D::SetPos(const float x, const float y, const float z) { m_pos = MyMath::Vector3(x, y, z); }

void Camera3D::Move(const long double dt)   // reference Camera3D.cpp:142-163
{
    const float deltaSpeed = static_cast<float>(dt) * 10.0f;
    MyMath::Vector3 total = m_staticRight * static_cast<float>(m_Keys.D - m_Keys.A) +
                            m_staticForward * static_cast<float>(m_Keys.W - m_Keys.S);
    total.Normalize_InPlace();
    m_pos.x = m_pos.x + (total.x * deltaSpeed);
    m_pos.z = m_pos.z + (total.z * deltaSpeed);
    m_pos.y += (m_Keys.Space - m_Keys.Shift) * deltaSpeed;
}

void Camera3D::AddRot(const long double, const short p, const short y, const short r)   // reference Camera3D.cpp:166-187
{
    const float deltaSpeed = 0.002f;
    m_rot.x -= ((float)p * deltaSpeed);
    m_rot.y += ((float)y * deltaSpeed);
    m_rot.z += ((float)r * deltaSpeed);
    if (m_rot.x > static_cast<float>(M_PI / 2.0)) m_rot.x = static_cast<float>((M_PI / 2.0) - 0.0001);
    if (m_rot.x < static_cast<float>(-M_PI / 2.0)) m_rot.x = static_cast<float>((-M_PI / 2.0) + 0.0001);
}
