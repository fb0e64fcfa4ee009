// RayTracing.h -- the reference's INNER seam (reference RayTracing.h:25-41): the static launcher RayTracingManager calls.
// Same name, same argument list.  What differs, by design of the new library:
//   * `objects` is the opaque scene handle Scene3D::GetObjects().m_deviceArray carries (the objects live in the library
//     as 64-byte PODs, not behind a device pointer table), `count` is informational;
//   * `params` is read on the HOST (the reference copies the block to the device first, RayTracingManager.cu:83; here it
//     is a kernel parameter);
//   * gridDims / blockDims are ignored: the ray kernel is persistent, one CTA per SM, and walks 16x16 tiles itself;
//   * `resultArray` is DEVICE memory of 20*x*y bytes (4-byte aligned) and receives the reference's raw cell buffer,
//     byte for byte what its RayTrace_* kernels leave there after the per-frame memset -- ready for the reference's own
//     host-side MinimizeRGB.  The launch is asynchronous, as in the reference; Synchronize() waits for it.
// RayTracingManager::Update does NOT go through here: it asks the library for the minimised stream directly.
#pragma once
#include "Object3D.h"
#include "RayTracingManager.h"

#ifndef __VECTOR_TYPES_H__          // no CUDA headers in this translation unit: a stand-in for CUDA's dim3
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int vx = 1, unsigned int vy = 1, unsigned int vz = 1) : x(vx), y(vy), z(vz) {}
};
#endif

class RayTracing
{
public:
    RayTracing() = delete;
    ~RayTracing() = delete;

    static void RayTrace(const dim3& gridDims, const dim3& blockDims, Object3D* DEVICE_MEMORY_PTR const objects,
                         const unsigned int count, const RayTracingCPUToGPUData* params, char* resultArray,
                         const RenderingMode mode);
    static void Synchronize(Object3D* DEVICE_MEMORY_PTR const objects);   // extension: cudaDeviceSynchronize of the reference's Update (:126)
};
