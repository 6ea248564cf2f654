/*
 * optix.h -- host-side emulation of the handful of OptiX 8 *device* API calls that
 * /root/reference/Core/RayTracer.cu makes, so that the reference's own raygen / intersection /
 * closest-hit / miss programs can be compiled UNMODIFIED by g++ and run on the CPU.
 *
 * TEST INFRASTRUCTURE ONLY.  This is our own file (the OptiX SDK is not in this image and is not copied);
 * it only provides the call signatures RayTracer.cu uses.  The semantics emulated are the documented ones:
 *   optixTrace            closest hit over all primitives in [tmin, tmax]; runs the intersection program per
 *                         candidate, then the closest-hit program of the closest primitive's SBT record, or miss.
 *   optixReportIntersection accepts the hit iff tmin <= t <= current tmax and shrinks tmax.
 * Traversal is brute force over every primitive, in index order (exact by definition; OptiX's own order is
 * unspecified).  See oracle/ref_harness.cpp for the driver.
 */
#pragma once

// CUDA puts the float overloads of sqrt/fabs/... into the global namespace (RayTracer.cu:12,246 rely on it:
// sqrt(float) must stay float).  libstdc++'s <math.h> wrapper does the same on the host; plain <cmath> does not.
#include <math.h>
#include <cmath>
#include <cstdint>
#include <cstring>

// CUDA decorations are meaningless on the host.
#ifndef __CUDACC__
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __host__
#define __host__
#endif
#ifndef __global__
#define __global__
#endif
#ifndef __constant__
#define __constant__
#endif
#ifndef __inline__
#define __inline__ inline
#endif
#endif

#include <vector_types.h>
#include <vector_functions.h>

typedef unsigned long long OptixTraversableHandle;
typedef unsigned long long CUdeviceptr_shim;
typedef unsigned int OptixVisibilityMask;
enum OptixRayFlags { OPTIX_RAY_FLAG_NONE = 0u };

struct RefShimState {
    uint3 launch_index;
    unsigned payload[2];
    // state visible to IS / CH / MS programs
    float3 ray_origin, ray_direction;
    float ray_tmin, ray_tmax;
    unsigned prim_index;
    const void* sbt_data;
    unsigned hit_kind;
    unsigned attr[6];
    // candidate reported by the intersection program currently running
    bool reported;
    float rep_t;
    unsigned rep_kind;
    unsigned rep_attr[6];
    int trace_nesting;
};
extern thread_local RefShimState g_ref_shim;

static inline uint3 optixGetLaunchIndex() { return g_ref_shim.launch_index; }
static inline unsigned optixGetPayload_0() { return g_ref_shim.payload[0]; }
static inline unsigned optixGetPayload_1() { return g_ref_shim.payload[1]; }
static inline CUdeviceptr_shim optixGetSbtDataPointer() { return (CUdeviceptr_shim)(uintptr_t)g_ref_shim.sbt_data; }
static inline unsigned optixGetPrimitiveIndex() { return g_ref_shim.prim_index; }
static inline float3 optixGetWorldRayOrigin() { return g_ref_shim.ray_origin; }
static inline float3 optixGetWorldRayDirection() { return g_ref_shim.ray_direction; }
static inline float optixGetRayTmin() { return g_ref_shim.ray_tmin; }
static inline float optixGetRayTmax() { return g_ref_shim.ray_tmax; }
static inline unsigned optixGetHitKind() { return g_ref_shim.hit_kind; }
static inline unsigned optixGetAttribute_0() { return g_ref_shim.attr[0]; }
static inline unsigned optixGetAttribute_1() { return g_ref_shim.attr[1]; }
static inline unsigned optixGetAttribute_2() { return g_ref_shim.attr[2]; }
static inline unsigned optixGetAttribute_3() { return g_ref_shim.attr[3]; }
static inline unsigned optixGetAttribute_4() { return g_ref_shim.attr[4]; }
static inline unsigned optixGetAttribute_5() { return g_ref_shim.attr[5]; }

static inline bool optixReportIntersection(float t, unsigned kind, unsigned a0, unsigned a1, unsigned a2,
                                           unsigned a3, unsigned a4, unsigned a5) {
    RefShimState& g = g_ref_shim;
    if (!(t >= g.ray_tmin && t <= g.ray_tmax)) return false;
    g.reported = true;
    g.rep_t = t;
    g.rep_kind = kind;
    g.rep_attr[0] = a0; g.rep_attr[1] = a1; g.rep_attr[2] = a2;
    g.rep_attr[3] = a3; g.rep_attr[4] = a4; g.rep_attr[5] = a5;
    return true;
}

// Defined in ref_harness.cpp, after the reference's programs are in scope.
void optixTrace(OptixTraversableHandle handle, float3 rayOrigin, float3 rayDirection, float tmin, float tmax,
                float rayTime, OptixVisibilityMask visibilityMask, unsigned rayFlags, unsigned SBToffset,
                unsigned SBTstride, unsigned missSBTIndex, unsigned& p0, unsigned& p1);

// Device intrinsics used by RayTracer.cu.
// device __powf is an approximation of powf; glibc's headers already declare a __powf, so use a macro.
#define __powf(a, b) powf((a), (b))
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
