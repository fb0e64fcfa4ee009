// Build shim (test infrastructure): stands in for <windows.h> so the reference's
// own pch.h parses on Linux. Provides only the handful of Win32 typedefs that the
// reference headers on the hot path mention (Camera3D.h:33-34,72; PrintMachine.h:39-49).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cfloat>
#include <cstring>
#include <stdexcept>
typedef void* HANDLE;
typedef short SHORT;
typedef unsigned long DWORD;
typedef int BOOL;
struct COORD { SHORT X, Y; };
struct POINT { long x, y; };
#ifndef WINAPI
#define WINAPI
#endif
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
