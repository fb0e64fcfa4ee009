// Build shim (test infrastructure): one wrapper TU around the reference's
// RayTracing.cu + RayTracingManager.cu.  Works around two MSVC-isms without
// touching the reference: the opaque `enum RenderingMode;` forward declaration
// (RayTracing.h:5) and the `<<< >>>` launch macro (pch.h:64).
#include "pch.h"
#undef CUDA_KERNEL
#define CUDA_KERNEL(g, b) * fakecuda::Launch{g, b}
#include "RayTracingManager.h"   // defines RenderingMode before RayTracing.h forward-declares it
#include "RayTracing.cu"
#include "RayTracingManager.cu"
