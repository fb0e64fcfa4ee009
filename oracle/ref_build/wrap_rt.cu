// Build shim (test infrastructure, nvcc build of the reference): RayTracing.h forward-declares
// `enum RenderingMode;` (ill-formed outside MSVC); including RayTracingManager.h first defines it.
#include "pch.h"
#include "RayTracingManager.h"
#include "RayTracing.cu"
