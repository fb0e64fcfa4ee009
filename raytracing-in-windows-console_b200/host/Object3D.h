// Object3D.h -- host-side mirror of the reference's scene object types (reference Object3D.h,
// Sphere.h, Plane.h).  The device-side Trace()/Update() members of the reference live in the
// CUDA library now (csrc/rtc_trace.cu, rtc_shade.cu); these classes are the host API only.
#pragma once
#include <cstdlib>

#include "MyMath.h"

#define DEVICE_MEMORY_PTR *

// Handle to the objects resident on the device (reference Object3D.h:6-12).  m_deviceArray is
// opaque here: the scene lives in the rtc context, not in a pointer table.
template <typename T>
struct DeviceObjectArray {
    T DEVICE_MEMORY_PTR m_deviceArray;
    unsigned int allocatedBytes;
    unsigned int count;
};

enum class ObjectType { None = 0, PlaneType, SphereType };

class Object3D
{
public:
    Object3D() = delete;
    Object3D(const MyMath::Vector3& center, const ObjectType type, const MyMath::Vector3& color)
        : m_center(center), m_type(type), m_color(color) {}
    virtual ~Object3D() noexcept = default;

    ObjectType GetType() const { return m_type; }
    MyMath::Vector3 GetPos() const { return m_center; }
    MyMath::Vector3 GetColor() const { return m_color; }
    void SetType(const ObjectType type) { m_type = type; }
    void SetMiddlePos(const MyMath::Vector3& center) { m_center = center; }

protected:
    MyMath::Vector3 m_center;
    ObjectType m_type;
    MyMath::Vector3 m_color;
};

class Sphere : public Object3D
{
public:
    // reference Sphere.cu:6-13: mover starts at -1, speed = (rand() % 300 + 100) / 100
    Sphere(const MyMath::Vector3& center, const float radius, const MyMath::Vector3& color)
        : Object3D(center, ObjectType::SphereType, color), m_radius(radius), mover(-1)
    {
        const int temp = rand() % 300 + 100;
        speed = static_cast<float>(temp) / 100.0f;
    }
    float GetRadius() const { return m_radius; }
    int GetMover() const { return mover; }
    float GetSpeed() const { return speed; }

private:
    float m_radius;
    int mover;
    float speed;
};

class Plane : public Object3D
{
public:
    // reference Plane.cu:6-12: the stored normal is normalised (zero-checked)
    Plane(const MyMath::Vector3& center, const MyMath::Vector3& normal, const MyMath::Vector3& color,
          const float width, const float height)
        : Object3D(center, ObjectType::PlaneType, color), m_normal(normal.Normalize()), m_width(width), m_height(height) {}
    MyMath::Vector3 GetNormal() const { return m_normal; }
    float GetWidth() const { return m_width; }
    float GetHeight() const { return m_height; }

private:
    MyMath::Vector3 m_normal;
    float m_width;
    float m_height;
};
