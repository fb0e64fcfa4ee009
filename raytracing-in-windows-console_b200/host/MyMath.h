// MyMath.h -- host-side mirror of the reference's MyMath namespace (reference MyMath.h:6-357,
// MyMath.cu:4-67): same type and function names, same arithmetic (binary32, left to right).
// Plain PODs here: the reference's virtual destructors (a vptr in every vector) were an ABI
// accident that forced byte-copies of host objects to the device; nothing crosses to the
// device from these types any more (include/rtc.h PODs do).
#pragma once
#include <cmath>

namespace MyMath
{
    class Vector3
    {
    public:
        Vector3(const float inX, const float inY, const float inZ) : x(inX), y(inY), z(inZ) {}
        Vector3(const int inX, const int inY, const int inZ)
            : x(static_cast<float>(inX)), y(static_cast<float>(inY)), z(static_cast<float>(inZ)) {}
        Vector3() : x(0.0f), y(0.0f), z(0.0f) {}

        Vector3 operator-(const Vector3& o) const { return Vector3(x - o.x, y - o.y, z - o.z); }
        void operator-=(const Vector3& o) { x -= o.x; y -= o.y; z -= o.z; }
        Vector3 operator+(const Vector3& o) const { return Vector3(x + o.x, y + o.y, z + o.z); }
        void operator+=(const Vector3& o) { x += o.x; y += o.y; z += o.z; }
        Vector3 operator*(const float s) const { return Vector3(x * s, y * s, z * s); }
        void operator*=(const float s) { x *= s; y *= s; z *= s; }
        Vector3 operator/(const float s) const { return Vector3(x / s, y / s, z / s); }
        void operator/=(const float s) { x /= s; y /= s; z /= s; }

        // zero-checked (reference MyMath.h:117-123)
        Vector3 Normalize() const
        {
            const float length = std::sqrt(x * x + y * y + z * z);
            const float divider = length < 0.000001f ? 0.0f : 1.0f / length;
            return Vector3(x * divider, y * divider, z * divider);
        }
        Vector3& Normalize_InPlace() { *this = Normalize(); return *this; }
        // no zero check (reference MyMath.h:139-146)
        Vector3 Normalize_GPU() const
        {
            const float length = 1.0f / std::sqrt(x * x + y * y + z * z);
            return Vector3(x * length, y * length, z * length);
        }
        Vector3& Normalize_InPlace_GPU() { *this = Normalize_GPU(); return *this; }
        float Length() const { return std::sqrt(x * x + y * y + z * z); }

        float x, y, z;
    };

    class Vector4
    {
    public:
        Vector4(const float inX, const float inY, const float inZ, const float inW) : x(inX), y(inY), z(inZ), w(inW) {}
        Vector4(const Vector3& v, const float inW) : x(v.x), y(v.y), z(v.z), w(inW) {}
        Vector4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
        Vector3 xyz() const { return Vector3(x, y, z); }
        Vector4 Normalize() const
        {
            const float length = 1.0f / std::sqrt(x * x + y * y + z * z + w * w);
            return Vector4(x * length, y * length, z * length, w * length);
        }
        float x, y, z, w;
    };

    class Matrix
    {
    public:
        Matrix(const Vector4& v1, const Vector4& v2, const Vector4& v3, const Vector4& v4) : row1(v1), row2(v2), row3(v3), row4(v4) {}
        Matrix() {}
        Vector4 Mult(const Vector4& v) const
        {
            return Vector4(row1.x * v.x + row1.y * v.y + row1.z * v.z + row1.w * v.w,
                           row2.x * v.x + row2.y * v.y + row2.z * v.z + row2.w * v.w,
                           row3.x * v.x + row3.y * v.y + row3.z * v.z + row3.w * v.w,
                           row4.x * v.x + row4.y * v.y + row4.z * v.z + row4.w * v.w);
        }
        Vector4 row1, row2, row3, row4;
    };

    inline float Dot(const Vector3& a, const Vector3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    inline float Dot(const Vector4& a, const Vector4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
    inline Vector3 Cross(const Vector3& a, const Vector3& b)
    {
        return Vector3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    }
    inline Vector3 ComponentMul(const Vector3& a, const Vector3& b) { return Vector3(a.x * b.x, a.y * b.y, a.z * b.z); }
    inline float Clamp(const float v, const float lo, const float hi) { const float r = v < lo ? lo : v; return r > hi ? hi : r; }
    inline int Clamp(const int v, const int lo, const int hi) { const int r = v < lo ? lo : v; return r > hi ? hi : r; }
    inline bool FloatEquals(float a, float b) { return std::fabs(a - b) < 1.1920928955078125e-7f; }
    inline int Min(int a, int b) { return a < b ? a : b; }
    inline int Max(int a, int b) { return a < b ? b : a; }
    inline float Min(float a, float b) { return a < b ? a : b; }
    inline float Max(float a, float b) { return a < b ? b : a; }
}
