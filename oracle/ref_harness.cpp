/*
 * ref_harness.cpp -- runs the reference's OWN device programs on the CPU.
 *
 * TEST INFRASTRUCTURE ONLY.  This translation unit #includes /root/reference/Core/RayTracer.cu (and through it
 * random.cuh, vec_math.h, RayTracer.h) from where it lies -- nothing is copied -- and compiles it with g++
 * against oracle/ref_shim/optix.h.  __raygen__rg, __intersection__hit_sphere, __closesthit__{lambertian,metal,
 * dielectric} and __miss__ms therefore execute exactly as written; the only emulated parts are OptiX's closed-
 * source traversal (brute force over all primitives here) and the SBT (filled the way Renderer.h:452-520 does).
 *
 * Output: oracle/_ref/libvenusaur_ref.so (git-ignored).  tests/golden/gen_golden.py uses it to produce the
 * committed golden vectors that pin oracle.cpp.  /root/reference does not exist on the GPU box, so nothing at
 * run time there depends on this file.
 *
 * Known, unavoidable difference from a device build: g++ evaluates the three random_float() calls inside
 * make_float3(...) (RayTracer.cu:99-115,145) right-to-left, and powf replaces the approximate __powf.
 */
#include <optix.h>

#include <atomic>
#include <thread>
#include <vector>

thread_local RefShimState g_ref_shim;

#include "RayTracer.cu"

namespace {

struct RefPrim {
    SphereHitGroupData data;
    unsigned type;
};

const RefPrim* g_prims = nullptr;
size_t g_nprims = 0;
int g_depth_override = 0;                // 0: keep the reference's compile-time max_depth = 4 (RayTracer.cu:172)
std::atomic<uint64_t> g_segments{0};

}  // namespace

void optixTrace(OptixTraversableHandle, float3 rayOrigin, float3 rayDirection, float tmin, float tmax, float,
                OptixVisibilityMask, unsigned, unsigned, unsigned, unsigned, unsigned& p0, unsigned& p1) {
    RefShimState& g = g_ref_shim;
    const RefShimState saved = g;        // the caller's hit state must survive the nested trace (RayTracer.cu:312,360)
    g.payload[0] = p0;
    g.payload[1] = p1;
    if (saved.trace_nesting == 0 && g_depth_override > 0) {
        // Test hook: lets the reference's closest-hit programs run with a deeper budget than the constant 4.
        PRD* prd = reinterpret_cast<PRD*>(unpackPointer(p0, p1));
        prd->depth = g_depth_override - 1;
    }
    g.trace_nesting = saved.trace_nesting + 1;
    g.ray_origin = rayOrigin;
    g.ray_direction = rayDirection;
    g.ray_tmin = tmin;
    g.ray_tmax = tmax;
    g_segments.fetch_add(1, std::memory_order_relaxed);

    bool any = false;
    size_t best = 0;
    unsigned best_kind = 0, best_attr[6] = {0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < g_nprims; i++) {
        g.prim_index = (unsigned)i;
        g.sbt_data = &g_prims[i].data;
        g.reported = false;
        __intersection__hit_sphere();
        if (g.reported) {
            any = true;
            best = i;
            g.ray_tmax = g.rep_t;
            best_kind = g.rep_kind;
            for (int k = 0; k < 6; k++) best_attr[k] = g.rep_attr[k];
        }
    }
    if (any) {
        g.prim_index = (unsigned)best;
        g.sbt_data = &g_prims[best].data;
        g.hit_kind = best_kind;
        for (int k = 0; k < 6; k++) g.attr[k] = best_attr[k];
        switch (g_prims[best].type) {     // Renderer.h:487-503: the program header chosen per sphere
            case 0: __closesthit__lambertian(); break;
            case 1: __closesthit__metal(); break;
            default: __closesthit__dielectric(); break;
        }
    } else {
        __miss__ms();
    }
    g = saved;
}

extern "C" {

struct ref_sphere { float cx, cy, cz, r, ax, ay, az, fuzz_or_ir; uint32_t type; };
struct ref_params {
    uint32_t width, height, samples_per_pixel, subframe_index;
    float origin[3], u[3], v[3], w[3], lens_radius;
};

// random.cuh / vec_math.h / RayTracer.cu helpers, for known-answer tests
uint32_t ref_tea1(uint32_t a, uint32_t b) { return tea<1>(a, b); }
uint32_t ref_tea4(uint32_t a, uint32_t b) { return tea<4>(a, b); }
uint32_t ref_tea16(uint32_t a, uint32_t b) { return tea<16>(a, b); }
uint32_t ref_lcg(uint32_t* s) { return lcg(*s); }
float ref_rnd(uint32_t* s) { return rnd(*s); }
void ref_normalize(const float* v, float* o) { float3 r = normalize(make_float3(v[0], v[1], v[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void ref_reflect(const float* i, const float* n, float* o) { float3 r = reflect(make_float3(i[0], i[1], i[2]), make_float3(n[0], n[1], n[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void ref_refract(const float* i, const float* n, float eta, float* o) { float3 r = refract(make_float3(i[0], i[1], i[2]), make_float3(n[0], n[1], n[2]), eta); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void ref_lerp(const float* a, const float* b, float t, float* o) { float3 r = lerp(make_float3(a[0], a[1], a[2]), make_float3(b[0], b[1], b[2]), t); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
void ref_make_color(const float* c, uint8_t* o) { uchar4 r = make_color(make_float3(c[0], c[1], c[2])); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
float ref_reflectance(float cosine, float ref_idx) { return reflectance(cosine, ref_idx); }
int ref_near_zero(const float* v) { return near_zero(make_float3(v[0], v[1], v[2])) ? 1 : 0; }
uint32_t ref_sizeof_params() { return (uint32_t)sizeof(Params); }
uint32_t ref_sizeof_sphere_record() { return (uint32_t)sizeof(SphereHitGroupData); }

// Runs __raygen__rg for the listed pixels (all when pixels == NULL).  accum_rgba is read (previous frame, used when
// subframe_index > 0, RayTracer.cu:208-213) and written; image gets the uchar4 pixels.  NOT re-entrant (params is
// the reference's global __constant__ Params).
uint64_t ref_render(const ref_sphere* spheres, uint64_t n, const ref_params* P, const uint32_t* pixels,
                    uint64_t n_pixels, float* accum_rgba, uint8_t* image, int depth_override, unsigned threads) {
    std::vector<RefPrim> prims(n);
    for (uint64_t i = 0; i < n; i++) {
        memset(&prims[i], 0, sizeof(RefPrim));
        prims[i].data.center = make_float3(spheres[i].cx, spheres[i].cy, spheres[i].cz);
        prims[i].data.radius = spheres[i].r;
        prims[i].type = spheres[i].type;
        switch (spheres[i].type) {        // Renderer.h:487-503
            case 0: prims[i].data.mat.albedo = make_float3(spheres[i].ax, spheres[i].ay, spheres[i].az); break;
            case 1: prims[i].data.mat.albedo = make_float3(spheres[i].ax, spheres[i].ay, spheres[i].az);
                    prims[i].data.mat.fuzz = spheres[i].fuzz_or_ir; break;
            default: prims[i].data.mat.ir = spheres[i].fuzz_or_ir; break;
        }
    }
    g_prims = prims.data();
    g_nprims = n;
    g_depth_override = depth_override;
    g_segments = 0;

    params.image = reinterpret_cast<uchar4*>(image);
    params.accum = reinterpret_cast<float4*>(accum_rgba);
    params.width = P->width;
    params.height = P->height;
    params.samples_per_pixel = P->samples_per_pixel;
    params.subframe_index = P->subframe_index;
    params.origin = make_float3(P->origin[0], P->origin[1], P->origin[2]);
    params.u = make_float3(P->u[0], P->u[1], P->u[2]);
    params.v = make_float3(P->v[0], P->v[1], P->v[2]);
    params.w = make_float3(P->w[0], P->w[1], P->w[2]);
    params.lens_radius = P->lens_radius;
    params.handle = 0;

    const uint64_t total = pixels ? n_pixels : (uint64_t)P->width * P->height;
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        while (true) {
            uint64_t b = next.fetch_add(32);
            if (b >= total) break;
            uint64_t e = b + 32 < total ? b + 32 : total;
            for (uint64_t k = b; k < e; k++) {
                uint32_t px = pixels ? pixels[k] : (uint32_t)k;
                memset(&g_ref_shim, 0, sizeof(g_ref_shim));
                g_ref_shim.launch_index = make_uint3(px % P->width, px / P->width, 0);
                __raygen__rg();
            }
        }
    };
    if (threads <= 1) work();
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < threads; t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    g_prims = nullptr;
    g_nprims = 0;
    return g_segments.load();
}

}  // extern "C"
