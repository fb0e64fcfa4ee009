// Scene3D.h -- scene owner, API as the reference's Scene3D (reference Scene3D.h/.cpp).
// Storage differs by design: objects go into the rtc context as 64-byte PODs, uploaded in one
// asynchronous copy when dirty (the reference issues two blocking cudaMemcpy per object).
#pragma once
#include "Object3D.h"

struct rtc_ctx;
struct rtc_mgpu;

// What DeviceObjectArray::m_deviceArray points at here: the scene lives in the library, either in one context (one GPU)
// or replicated by the multi-GPU frame driver (RTC_GPUS > 1).  Exactly one of the two is set.
struct SceneBackend {
    rtc_ctx* ctx = nullptr;
    rtc_mgpu* mgpu = nullptr;
};

#define FIVE_MEGABYTES 5'000'000
#define HUNDRED_MEGABYTES 100'000'000

class Scene3D
{
public:
    Scene3D() = default;
    ~Scene3D() = default;

    void Init();                          // device state + the reference's default scene (Scene3D.cpp:28-33)
    void Update(const long double dt);    // no-op: physics runs on the device (Scene3D.cpp:89-92)
    void CleanUp();
    void CreatePlane(const MyMath::Vector3& middlePos, const MyMath::Vector3& normal, const MyMath::Vector3& color,
                     const float width, const float height);
    void CreateSphere(const float radius, const MyMath::Vector3& middlePos, const MyMath::Vector3& color);
    DeviceObjectArray<Object3D*> GetObjects();

    // Extensions for headless use.
    void InitEmpty();                     // device state only, no default objects
    static rtc_ctx* Context();            // the process-wide rtc context (device 0 or $RTC_DEVICE); NULL in multi-GPU mode
    // The process-wide backend.  Environment: RTC_GPUS=N (N > 1: row bands over N GPUs through rtc_mgpu), RTC_DEVICES="0,1,.."
    // (device ordinals, may repeat), RTC_GATHER=host|p2p, RTC_DEVICE (single-GPU ordinal).
    static SceneBackend* Backend();

private:
    unsigned int m_count = 0;
    unsigned int m_spheres = 0, m_planes = 0;
};
