#include "Scene3D.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/rtc.h"

namespace {
SceneBackend g_backend;
bool g_backend_up = false;
void check(int rc, const char* what)
{
    if (rc != RTC_OK) throw std::runtime_error(std::string(what) + ": " + rtc_last_error());
}
}  // namespace

SceneBackend* Scene3D::Backend()
{
    if (!g_backend_up) {
        const char* gpus = getenv("RTC_GPUS");
        const int n = gpus ? atoi(gpus) : 1;
        if (n > 1) {
            int ids[16], n_ids = 0;
            if (const char* devs = getenv("RTC_DEVICES")) {
                std::string d(devs);
                size_t pos = 0;
                while (n_ids < 16 && pos <= d.size()) {
                    const size_t comma = d.find(',', pos);
                    ids[n_ids++] = atoi(d.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos).c_str());
                    if (comma == std::string::npos) break;
                    pos = comma + 1;
                }
            }
            const char* ga = getenv("RTC_GATHER");
            const int gather = (ga && std::string(ga) == "p2p") ? RTC_GATHER_P2P : RTC_GATHER_HOST;
            check(rtc_mgpu_create(&g_backend.mgpu, n, n_ids == n ? ids : nullptr, gather), "rtc_mgpu_create");
        } else {
            const char* dev = getenv("RTC_DEVICE");
            check(rtc_create(&g_backend.ctx, dev ? atoi(dev) : 0), "rtc_create");   // throws when no B200 is usable: no CPU fallback
        }
        g_backend_up = true;
    }
    return &g_backend;
}

rtc_ctx* Scene3D::Context() { return Backend()->ctx; }

void Scene3D::InitEmpty()
{
    SceneBackend* b = Backend();
    check(b->mgpu ? rtc_mgpu_scene_clear(b->mgpu) : rtc_scene_clear(b->ctx), "rtc_scene_clear");
    m_count = m_spheres = m_planes = 0;
}

void Scene3D::Init()
{
    InitEmpty();
    CreateSphere(7.0f, MyMath::Vector3(0.0f, 10.0f, 20.0f), MyMath::Vector3(255.0f, 1.0f, 1.0f));
    CreateSphere(6.0f, MyMath::Vector3(5.0f, 10.0f, 20.0f), MyMath::Vector3(1.0f, 255.0f, 1.0f));
    CreateSphere(10.0f, MyMath::Vector3(10.0f, 10.0f, 40.0f), MyMath::Vector3(1.0f, 1.0f, 255.0f));
    CreateSphere(3.0f, MyMath::Vector3(5.0f, 10.0f, 20.0f), MyMath::Vector3(225.0f, 210.0f, 20.0f));
    CreateSphere(4.0f, MyMath::Vector3(-5.0f, 10.0f, 40.0f), MyMath::Vector3(225.0f, 10.0f, 220.0f));
    CreatePlane(MyMath::Vector3(0.0f, -3.0f, 30.0f), MyMath::Vector3(0.0f, 1.0f, 0.0f), MyMath::Vector3(100.0f, 100.0f, 100.0f), 10, 20);
}

void Scene3D::CreateSphere(const float radius, const MyMath::Vector3& middlePos, const MyMath::Vector3& color)
{
    // reference Scene3D.cpp:131-141: when the 5 MB typed array is full the object is silently dropped
    if ((size_t)(m_spheres + 1) * 96 > FIVE_MEGABYTES) return;
    // reference Scene3D.cpp:107-116: the pointer table may not reach 100 MB
    if ((size_t)(m_count + 1) * sizeof(void*) * 2 >= HUNDRED_MEGABYTES)
        throw std::runtime_error("Error! Out of dedicated memory when trying to create an object.");
    const Sphere s(middlePos, radius, color);        // draws speed from rand() like the reference
    const float c[3] = {middlePos.x, middlePos.y, middlePos.z}, k[3] = {color.x, color.y, color.z};
    SceneBackend* b = Backend();
    check(b->mgpu ? rtc_mgpu_scene_add_sphere(b->mgpu, c, radius, k, s.GetSpeed(), s.GetMover())
                  : rtc_scene_add_sphere(b->ctx, c, radius, k, s.GetSpeed(), s.GetMover()), "rtc_scene_add_sphere");
    ++m_spheres; ++m_count;
}

void Scene3D::CreatePlane(const MyMath::Vector3& middlePos, const MyMath::Vector3& normal, const MyMath::Vector3& color,
                          const float width, const float height)
{
    if ((size_t)(m_planes + 1) * 96 > FIVE_MEGABYTES) return;
    const float c[3] = {middlePos.x, middlePos.y, middlePos.z}, n[3] = {normal.x, normal.y, normal.z}, k[3] = {color.x, color.y, color.z};
    SceneBackend* b = Backend();
    check(b->mgpu ? rtc_mgpu_scene_add_plane(b->mgpu, c, n, k, width, height) : rtc_scene_add_plane(b->ctx, c, n, k, width, height),
          "rtc_scene_add_plane");
    ++m_planes; ++m_count;
}

void Scene3D::Update(const long double) {}

void Scene3D::CleanUp()
{
    if (g_backend.mgpu) { rtc_mgpu_destroy(g_backend.mgpu); g_backend.mgpu = nullptr; }
    if (g_backend.ctx) { rtc_destroy(g_backend.ctx); g_backend.ctx = nullptr; }
    g_backend_up = false;
    m_count = m_spheres = m_planes = 0;
}

DeviceObjectArray<Object3D*> Scene3D::GetObjects()
{
    DeviceObjectArray<Object3D*> a;
    a.m_deviceArray = reinterpret_cast<Object3D**>(Backend());   // opaque: the objects live in the library
    a.allocatedBytes = m_count * (unsigned)sizeof(rtc_object);
    a.count = m_count;
    return a;
}
