// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
// C-ABI driver around the reference's OWN sources (compiled unmodified from
// /root/reference/ConsoleProject through the shims in this directory).  It exists so
// that (1) the restated oracle (oracle/rt_oracle.c) can be pinned against the real
// reference, (2) golden vectors under tests/golden/ can be generated, and (3) bench.py's
// `--impl reference` / `cpu_baseline` legs can time the reference's CPU path.
// Output: oracle/_ref/libref_cpu.so (git-ignored).  Nothing in the product links this.
#include "pch.h"

// The reference keeps the state we need to set/read (sphere speed/mover, the typed
// object arrays) private; open it up for this TU only.  Layout is unaffected.
#define private public
#define protected public
#include "Scene3D.h"
#include "Camera3D.h"
#include "RayTracingManager.h"
#undef private
#undef protected
#include "RayTracing.h"
#include "PrintMachine.h"

#include "../../include/rtc.h"   // POD layouts only (rtc_object, rtc_params)

thread_local fake_uint3 blockIdx, threadIdx;
thread_local dim3 blockDim, gridDim;
namespace fakecuda { int g_threads = 1; }

// Non-static functions defined in the reference's RayTracing.cu / ANSIRGB.h.
MyMath::Vector3 BlinnPhongShading(
    const MyMath::Vector3&, const MyMath::Vector3&, const MyMath::Vector3&,
    const MyMath::Vector3&, const float, const MyMath::Vector3&, const float,
    const MyMath::Vector3&, const MyMath::Vector3&, const MyMath::Vector3&);
void RayTrace(const RayTraceInputData&, RayTraceReturnData&);
uint8_t ansi256_from_rgb(uint32_t rgb);

namespace {

RayTracingCPUToGPUData to_ref_params(const rtc_params* p)
{
    RayTracingCPUToGPUData q;
    const float* m = p->inv_view;
    q.inverseVMatrix.row1 = MyMath::Vector4(m[0], m[1], m[2], m[3]);
    q.inverseVMatrix.row2 = MyMath::Vector4(m[4], m[5], m[6], m[7]);
    q.inverseVMatrix.row3 = MyMath::Vector4(m[8], m[9], m[10], m[11]);
    q.inverseVMatrix.row4 = MyMath::Vector4(m[12], m[13], m[14], m[15]);
    q.camPos = MyMath::Vector3(p->cam_pos[0], p->cam_pos[1], p->cam_pos[2]);
    q.x = p->x; q.y = p->y;
    q.element1 = p->element1; q.element2 = p->element2; q.camFarDist = p->cam_far;
    return q;
}

void build_scene(Scene3D& scene, const rtc_object* objs, uint32_t n, int use_default)
{
    scene.Init();                      // allocates the three arrays + the default 5 spheres + plane
    if (use_default) return;
    scene.m_deviceObjects.count = 0;   // drop the default objects; arrays are reused
    scene.m_devicePlanes.count = 0;
    scene.m_deviceSpheres.count = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const rtc_object& o = objs[i];
        MyMath::Vector3 c(o.center[0], o.center[1], o.center[2]);
        MyMath::Vector3 col(o.color[0], o.color[1], o.color[2]);
        if (o.type == RTC_OBJ_SPHERE) {
            scene.CreateSphere(o.radius, c, col);
            Sphere& s = scene.m_deviceSpheres.m_deviceArray[scene.m_deviceSpheres.count - 1];
            s.speed = o.speed; s.mover = o.mover;       // replace the rand() draw (Sphere.cu:11-12)
        } else if (o.type == RTC_OBJ_PLANE) {
            MyMath::Vector3 nn(o.normal[0], o.normal[1], o.normal[2]);
            scene.CreatePlane(c, nn, col, o.width, o.height);
            Plane& pl = scene.m_devicePlanes.m_deviceArray[scene.m_devicePlanes.count - 1];
            pl.m_normal = nn;                           // rtc_object carries the final (normalised) normal
        }
    }
}

void read_scene(Scene3D& scene, rtc_object* out, uint32_t cap, uint32_t* n)
{
    uint32_t cnt = scene.m_deviceObjects.count;
    if (n) *n = cnt;
    for (uint32_t i = 0; i < cnt && i < cap; ++i) {
        Object3D* o = scene.m_deviceObjects.m_deviceArray[i];
        rtc_object r; memset(&r, 0, sizeof r);
        r.type = (int32_t)o->GetType();
        MyMath::Vector3 c = o->GetPos(), col = o->GetColor();
        r.center[0] = c.x; r.center[1] = c.y; r.center[2] = c.z;
        r.color[0] = col.x; r.color[1] = col.y; r.color[2] = col.z;
        if (o->GetType() == ObjectType::SphereType) {
            Sphere* s = (Sphere*)o; r.radius = s->GetRadius(); r.speed = s->speed; r.mover = s->mover;
        } else if (o->GetType() == ObjectType::PlaneType) {
            Plane* p = (Plane*)o; MyMath::Vector3 nn = p->GetNormal();
            r.normal[0] = nn.x; r.normal[1] = nn.y; r.normal[2] = nn.z;
            r.width = p->GetWidth(); r.height = p->GetHeight();
        }
        out[i] = r;
    }
}

}  // namespace

extern "C" {

// Full RayTracingManager::Update on the CPU (reference RayTracingManager.cu:76-154).
// raw_out (optional): the un-minimised 20*x*y cell buffer; min_out: the bytes handed to
// PrintMachine::SetDataInBackBuffer.  objs_after (optional, cap n): object state after the
// UpdateObjects step.  Returns seconds spent inside Update (<0 on error).
double ref_update(const rtc_object* objs, uint32_t n, int use_default_scene,
                  const rtc_params* p, int mode, double dt, int nthreads,
                  char* raw_out, size_t raw_cap, char* min_out, size_t min_cap, size_t* min_size,
                  rtc_object* objs_after)
{
    fakecuda::g_threads = nthreads;
    PrintMachine::Start(p->x, p->y);
    double secs = -1.0;
    {
        RayTracingManager mgr;
        mgr.SetRenderingMode((RenderingMode)mode);
        Scene3D scene;
        build_scene(scene, objs, n, use_default_scene);
        RayTracingCPUToGPUData q = to_ref_params(p);
        auto t0 = std::chrono::steady_clock::now();
        mgr.Update(q, scene.GetObjects(), dt);
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const size_t sz = PrintMachine::GetPrintSize();
        if (min_size) *min_size = sz;
        if (min_out && sz <= min_cap) memcpy(min_out, PrintMachine::GetBackBuffer(), sz);
        if (raw_out) {
            const size_t full = PrintMachine::GetMaxSize();
            memcpy(raw_out, mgr.m_hostResultArray.get(), full < raw_cap ? full : raw_cap);
        }
        if (objs_after) { uint32_t cnt; read_scene(scene, objs_after, n ? n : 6, &cnt); }
        scene.CleanUp();
    }
    return secs;
}

// Only the render kernel (RayTracing::RayTrace, RayTracing.cu:797-867) on a window of
// 16-row block rows [brow0, brow1) of the frame -- the bounded sample used by the CPU
// baseline.  Returns seconds; rays_out = rays traced.
double ref_trace_blockrows(const rtc_object* objs, uint32_t n, int use_default_scene,
                           const rtc_params* p, int mode, int nthreads,
                           uint32_t brow0, uint32_t brow1, unsigned long long* rays_out)
{
    fakecuda::g_threads = nthreads;
    PrintMachine::Start(p->x, p->y);
    Scene3D scene;
    build_scene(scene, objs, n, use_default_scene);
    RayTracingCPUToGPUData q = to_ref_params(p);
    char* result = (char*)malloc(PrintMachine::GetMaxSize());
    const unsigned gx = (unsigned)std::ceil((p->x + 1) / 16.0);
    const unsigned gy_full = (unsigned)std::ceil(p->y / 16.0);
    if (brow1 > gy_full) brow1 = gy_full;
    DeviceObjectArray<Object3D*> arr = scene.GetObjects();
    // Run the reference kernel on block rows [brow0,brow1): one launch per block row with
    // blockIdx.y forced by a 1-row grid is not expressible, so loop here exactly like the
    // fake launcher does, but over the window only.
    auto t0 = std::chrono::steady_clock::now();
    {
        const unsigned long long nblocks = (unsigned long long)gx * (brow1 - brow0);
        auto run = [&](unsigned long long lo, unsigned long long hi) {
            gridDim = dim3(gx, gy_full, 1); blockDim = dim3(16, 16, 1);
            for (unsigned long long blk = lo; blk < hi; ++blk) {
                blockIdx.x = (unsigned)(blk % gx); blockIdx.y = brow0 + (unsigned)(blk / gx);
                for (unsigned ty = 0; ty < 16; ++ty) for (unsigned tx = 0; tx < 16; ++tx) {
                    threadIdx.x = tx; threadIdx.y = ty;
                    switch ((RenderingMode)mode) {
                    case BIT_ASCII: RayTrace_ASCII(arr.m_deviceArray, arr.count, &q, result); break;
                    case BIT_PIXEL: RayTrace_PIXEL(arr.m_deviceArray, arr.count, &q, result); break;
                    case RGB_ASCII: RayTrace_RGB_ASCII(arr.m_deviceArray, arr.count, &q, result); break;
                    case RGB_NORMALS: RayTrace_RGB_NORMALS(arr.m_deviceArray, arr.count, &q, result); break;
                    default: RayTrace_RGB_PIXEL(arr.m_deviceArray, arr.count, &q, result); break;
                    }
                }
            }
        };
        int T = nthreads < 1 ? 1 : nthreads;
        if ((unsigned long long)T > nblocks) T = (int)nblocks;
        if (T <= 1) run(0, nblocks);
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < T; ++t) pool.emplace_back(run, nblocks * t / T, nblocks * (t + 1) / T);
            for (auto& th : pool) th.join();
        }
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rays_out) {
        unsigned long long rows = 0;
        for (unsigned b = brow0; b < brow1; ++b) {
            unsigned r0 = b * 16, r1 = r0 + 16; if (r1 > p->y) r1 = p->y; rows += (r1 > r0) ? r1 - r0 : 0;
        }
        *rays_out = rows * (p->x - 1);
    }
    free(result);
    scene.CleanUp();
    return secs;
}

// Engine3D::Render's parameter block from the reference Camera3D (Engine3D.cpp:81-97).
int ref_camera_params(uint32_t x, uint32_t y, const float pos[3], const float rot[3], rtc_params* out)
{
    PrintMachine::Start(x, y);
    Camera3D cam;
    cam.SetPos(pos[0], pos[1], pos[2]);
    cam.SetRot(rot[0], rot[1], rot[2]);
    cam.Init();
    cam.Update();
    MyMath::Matrix inv = cam.GetInverseVMatrix();
    const MyMath::Vector4* rows[4] = { &inv.row1, &inv.row2, &inv.row3, &inv.row4 };
    for (int r = 0; r < 4; ++r) {
        out->inv_view[4 * r + 0] = rows[r]->x; out->inv_view[4 * r + 1] = rows[r]->y;
        out->inv_view[4 * r + 2] = rows[r]->z; out->inv_view[4 * r + 3] = rows[r]->w;
    }
    out->cam_pos[0] = cam.GetPos().x; out->cam_pos[1] = cam.GetPos().y; out->cam_pos[2] = cam.GetPos().z;
    out->x = (uint32_t)PrintMachine::GetWidth(); out->y = (uint32_t)PrintMachine::GetHeight();
    out->element1 = cam.GetPMatrix().row1.x; out->element2 = cam.GetPMatrix().row2.y;
    out->cam_far = cam.GetFarPlaneDistance();
    return 0;
}

// The reference default scene (Scene3D.cpp:28-33) as rtc_objects.
int ref_default_scene(rtc_object* out, uint32_t cap, uint32_t* n)
{
    Scene3D scene; scene.Init();
    read_scene(scene, out, cap, n);
    scene.CleanUp();
    return 0;
}

// ---- per-function known-answer hooks ------------------------------------------------------
// Sphere::Trace (Sphere.cu:30-68) with the per-ray terms computed as RayTrace does (:91-93).
int ref_sphere_trace(const float c[3], float radius, const float o[3], const float d[3],
                     float* dist, float nrm[3])
{
    Sphere s(MyMath::Vector3(c[0], c[1], c[2]), radius, MyMath::Vector3(0.f, 0.f, 0.f));
    ObjectTraceInputData in; ObjectTraceReturnData ret;
    in.origin = MyMath::Vector3(o[0], o[1], o[2]); in.direction = MyMath::Vector3(d[0], d[1], d[2]);
    in.a = MyMath::Dot(in.direction, in.direction); in.fourA = 4.0f * in.a; in.divTwoA = 1.0f / (2.0f * in.a);
    s.Trace(in, ret);
    *dist = ret.distance; nrm[0] = ret.normal.x; nrm[1] = ret.normal.y; nrm[2] = ret.normal.z;
    return ret.bHit ? 1 : 0;
}
// Plane::Trace (Plane.cu:38-73); `normal` is normalised by the Plane ctor (Plane.cu:9).
int ref_plane_trace(const float c[3], const float normal[3], float w, float h,
                    const float o[3], const float d[3], float* dist, float nrm[3])
{
    Plane p(MyMath::Vector3(c[0], c[1], c[2]), MyMath::Vector3(normal[0], normal[1], normal[2]),
            MyMath::Vector3(0.f, 0.f, 0.f), w, h);
    ObjectTraceInputData in; ObjectTraceReturnData ret;
    in.origin = MyMath::Vector3(o[0], o[1], o[2]); in.direction = MyMath::Vector3(d[0], d[1], d[2]);
    p.Trace(in, ret);
    *dist = ret.distance; nrm[0] = ret.normal.x; nrm[1] = ret.normal.y; nrm[2] = ret.normal.z;
    return ret.bHit ? 1 : 0;
}
int ref_plane_normal(const float normal[3], float out[3])
{
    Plane p(MyMath::Vector3(0.f, 0.f, 0.f), MyMath::Vector3(normal[0], normal[1], normal[2]),
            MyMath::Vector3(0.f, 0.f, 0.f), 1.f, 1.f);
    MyMath::Vector3 n = p.GetNormal(); out[0] = n.x; out[1] = n.y; out[2] = n.z; return 0;
}
// CalculateInitialDirection (RayTracing.cu:9-24) for cell (row, col).
int ref_initial_direction(const rtc_params* p, uint32_t row, uint32_t col, float d[3])
{
    RayTracingCPUToGPUData q = to_ref_params(p);
    blockDim = dim3(16, 16, 1); gridDim = dim3(1, 1, 1);
    blockIdx.x = col / 16; blockIdx.y = row / 16; threadIdx.x = col % 16; threadIdx.y = row % 16;
    MyMath::Vector3 v = CalculateInitialDirection(&q);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; return 0;
}
// RayTrace (RayTracing.cu:81-168): nearest hit + shading for one ray.
int ref_raytrace(const rtc_object* objs, uint32_t n, const float o[3], const float d[3],
                 float* dist, float nrm[3], float col[3], float* shading_value)
{
    Scene3D scene; build_scene(scene, objs, n, 0);
    RayTraceInputData in; RayTraceReturnData ret;
    in.origin = MyMath::Vector3(o[0], o[1], o[2]); in.direction = MyMath::Vector3(d[0], d[1], d[2]);
    in.objectCount = scene.GetObjects().count; in.objects = scene.GetObjects().m_deviceArray;
    RayTrace(in, ret);
    *dist = ret.distance; *shading_value = ret.shadingValue;
    nrm[0] = ret.normal.x; nrm[1] = ret.normal.y; nrm[2] = ret.normal.z;
    col[0] = ret.color.x; col[1] = ret.color.y; col[2] = ret.color.z;
    scene.CleanUp();
    return 0;
}
// BlinnPhongShading with the constants of its one call site (RayTracing.cu:143-152).
int ref_blinn_phong(const float kd[3], const float point[3], const float view[3], const float nrm[3], float out[3])
{
    MyMath::Vector3 r = BlinnPhongShading(
        MyMath::Vector3(kd[0], kd[1], kd[2]), MyMath::Vector3(1.0f, 1.0f, 1.0f), MyMath::Vector3(1.0f, 50.0f, 0.0f),
        MyMath::Vector3(1.0f, 1.0f, 1.0f), 2000.0f, MyMath::Vector3(1.0f, 1.0f, 1.0f), 3000.0f,
        MyMath::Vector3(point[0], point[1], point[2]), MyMath::Vector3(view[0], view[1], view[2]),
        MyMath::Vector3(nrm[0], nrm[1], nrm[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; return 0;
}
// ansi256_from_rgb (ANSIRGB.h:141-189) for rgb in [first, first+count).
int ref_ansi256_range(uint32_t first, uint32_t count, uint8_t* out)
{
    for (uint32_t i = 0; i < count; ++i) out[i] = ansi256_from_rgb(first + i);
    return 0;
}
// GetASCIICharacter (RayTracing.cu:26-39).
int ref_ascii_char(float distance, float far_plane, float shading_value)
{
    return (int)(unsigned char)GetASCIICharacter(distance, far_plane, shading_value);
}
int ref_sizes(int* sz)   // ABI facts the survey quotes
{
    sz[0] = (int)sizeof(MyMath::Vector3); sz[1] = (int)sizeof(Object3D); sz[2] = (int)sizeof(Sphere);
    sz[3] = (int)sizeof(Plane); sz[4] = (int)sizeof(RayTracingCPUToGPUData); return 5;
}

}  // extern "C"
