// vn_math.cuh -- per-path device math of the B200 path tracer: RNG, samplers, camera ray, ray/sphere
// intersection, the three scatter functions, sky, sRGB quantisation.
//
// Two builds of every kernel are made from this one header:
//   VN_EXACT=1  compiled with -fmad=false: each float operation is one IEEE-754 binary32 op in the order the
//               reference writes it (RayTracer.cu / vec_math.h), FP64 where the reference uses FP64.  Bit-identical
//               to the host oracle: the DEFAULT build, the one bench.py measures and the parity tests check.
//   VN_EXACT=0  the opt-in relaxed build (VN_FAST): FMA contraction, approximate reciprocal / rsqrt, FP32
//               replacements for the FP64 fragments.  Checked statistically against the same oracle; not within the image tolerance.
// The header also compiles as plain C++ (g++) so that tests can run the exact math on the CPU; that host build
// is test-only -- the library itself has no CPU path.
//
// Citations: file:line in /root/reference/Core/.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VN_HD __host__ __device__ __forceinline__
#else
#define VN_HD inline
#endif

#ifndef VN_EXACT
#define VN_EXACT 1
#endif

#if VN_EXACT || !defined(__CUDA_ARCH__)
#define VN_FAST_DEVICE 0
#else
#define VN_FAST_DEVICE 1
#endif

namespace vn {

struct f3 { float x, y, z; };

VN_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
VN_HD f3 mk3(float s) { return mk3(s, s, s); }
VN_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
VN_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
VN_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
VN_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
VN_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
VN_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
VN_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            // vec_math.h:527-530

VN_HD uint32_t f2u_early(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}

// ---- mode-dependent scalar primitives
VN_HD float rcp(float a) {
#if VN_FAST_DEVICE
    return __fdividef(1.0f, a);       // MUFU.RCP + FMUL
#else
    return 1.0f / a;
#endif
}
VN_HD float fdiv(float a, float b) {
#if VN_FAST_DEVICE
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
VN_HD float fsqrt(float a) {
#if VN_FAST_DEVICE
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));   // MUFU.SQRT
    return r;
#else
    return sqrtf(a);
#endif
}
// 1/sqrt(a): the reference computes 1.0f / sqrtf(a) (vec_math.h:545-549)
VN_HD float inv_sqrt(float a) {
#if VN_FAST_DEVICE
    return rsqrtf(a);
#else
    return 1.0f / sqrtf(a);
#endif
}

VN_HD f3 normalize(f3 v) { float invLen = inv_sqrt(dot(v, v)); return v * invLen; }   // vec_math.h:545-549
VN_HD f3 reflect(f3 i, f3 n) { return i - 2.0f * n * dot(n, i); }                     // vec_math.h:558-561
VN_HD f3 refract(f3 uv, f3 n, float etai_over_etat) {                                 // vec_math.h:564-570
    float cos_theta = fminf(dot(-uv, n), 1.0f);
    f3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    f3 r_out_parallel = -fsqrt(fabsf(1.0f - dot(r_out_perp, r_out_perp))) * n;
    return r_out_perp + r_out_parallel;
}
VN_HD f3 lerp3(f3 a, f3 b, float t) { return a + t * (b - a); }                        // vec_math.h:500-503
VN_HD float clamp01(float f) { return fmaxf(0.0f, fminf(f, 1.0f)); }                  // vec_math.h:119-122

// ---- RNG: random.cuh:31-67, uint32, bit-exact in both builds
VN_HD uint32_t tea4(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 4; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
VN_HD uint32_t lcg(uint32_t& prev) {
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}
// (float)lcg / (float)0x01000000: dividing by 2^24 is exact, so multiplying by 2^-24 gives the same bits.
VN_HD float rnd(uint32_t& prev) { return (float)lcg(prev) * 5.9604644775390625e-8f; }
// random_float(seed,-1,1) = -1 + (1 - -1) * rnd (RayTracer.cu:93-97).  With n = lcg & 0xFFFFFF every step of that expression is exact:
// n * 2^-24, doubled, minus one = (n - 2^23) * 2^-23.  The same value comes out of three instructions instead of six: the
// state shifted left by 8 drops the 8 bits the mask would clear, flipping bit 31 subtracts 2^31 in two's complement, so the
// signed word is (n - 2^23) * 2^8 -- 24 significant bits, exactly convertible -- and 2^-31 scales it back (n = 2^23 gives +0 either
// way).  The shift and the flip fold into the LCG's multiply-add.  Checked against the literal form for all 2^24 values of n in
// tests/test_host_logic.py.
VN_HD float rnd_pm1(uint32_t& prev) {
    prev = 1664525u * prev + 1013904223u;
    return (float)(int32_t)((prev << 8) ^ 0x80000000u) * 4.656612873077392578125e-10f;
}
// the literal form, kept for the test above
VN_HD float rnd_pm1_literal(uint32_t& prev) { return -1.0f + 2.0f * rnd(prev); }

// RayTracer.cu:117-125; draws sequenced x, y, z (the contract of SURVEY 3.4)
VN_HD f3 random_in_unit_sphere(uint32_t& seed) {
    while (true) {
        f3 p;
        p.x = rnd_pm1(seed);
        p.y = rnd_pm1(seed);
        p.z = rnd_pm1(seed);
        if (dot(p, p) >= 1.0f) continue;
        return p;
    }
}
// RayTracer.cu:141-149
VN_HD void random_in_unit_disk(uint32_t& seed, float& px, float& py) {
    while (true) {
        float x = rnd_pm1(seed);
        float y = rnd_pm1(seed);
        if (x * x + y * y + 0.0f * 0.0f >= 1.0f) continue;
        px = x; py = y;
        return;
    }
}

// ---- launch constants (Params, RayTracer.h:3-17, plus hoisted loop invariants)
struct Camera {
    f3 origin, u, v, w;
    f3 u_unit, v_unit;       // normalize(params.u), normalize(params.v): loop invariants of get_ray (RayTracer.cu:155-156)
    float lens_radius;
    float wm1, hm1;          // float(width-1), float(height-1)
    float inv_wm1, inv_hm1;  // RN(1 / wm1), RN(1 / hm1): the FAST build multiplies by them, the IEEE build uses them in div_by_const()
    uint32_t div_exact;      // both divisors satisfy div_by_const_ok()
};

// a / b for the two per-launch constants of the jitter below (b = float(width - 1), float(height - 1)) without the 10-instruction
// IEEE division: with rb = RN(1 / b) formed once on the host, q = RN(a * rb) is within an ulp of the quotient, r = a - q * b is exact
// in one fma, and RN(q + r * rb) is the correctly rounded a / b (Markstein's theorem; it needs a divisor whose significand is not
// all ones, which is what Camera::div_exact records -- otherwise the plain division runs).  tests/test_host_logic.py compares it
// with a / b over every pixel column of the usual widths and 2^12 jitters each.
VN_HD float div_by_const(float a, float b, float rb) {
    const float q = a * rb;
    const float r = fmaf(-q, b, a);
    return fmaf(r, rb, q);
}
VN_HD bool div_by_const_ok(float b) { return b >= 1.0f && b < 16777216.0f && (f2u_early(b) & 0x007FFFFFu) != 0x007FFFFFu; }

// get_ray + the jitter of __raygen__rg: RayTracer.cu:151-161,173-174
VN_HD void camera_ray(const Camera& c, uint32_t px, uint32_t py, uint32_t& seed, f3& origin, f3& direction) {
    float ju = rnd(seed);
    float jv = rnd(seed);
#if VN_FAST_DEVICE
    float s = 2.0f * ((float)px + ju) * c.inv_wm1 - 1.0f;
    float t = 2.0f * ((float)py + jv) * c.inv_hm1 - 1.0f;
#else
    const float su = 2.0f * ((float)px + ju), tv = 2.0f * ((float)py + jv);
    float s = (c.div_exact ? div_by_const(su, c.wm1, c.inv_wm1) : su / c.wm1) - 1.0f;
    float t = (c.div_exact ? div_by_const(tv, c.hm1, c.inv_hm1) : tv / c.hm1) - 1.0f;
#endif
    float dx, dy;
    random_in_unit_disk(seed, dx, dy);
    float rdx = c.lens_radius * dx, rdy = c.lens_radius * dy;
    f3 offset = c.u_unit * rdx + c.v_unit * rdy;
    origin = c.origin + offset;
    direction = c.w + s * c.u * 0.5f + t * c.v * 0.5f - offset;
}

// ---- ray / sphere: RayTracer.cu:229-270.  Returns the accepted root or -1.  `a` = dot(d,d).
VN_HD float sphere_root(f3 o, f3 d, float a, float inv_a, float cx, float cy, float cz, float r, float t_min, float t_max) {
    f3 oc = o - mk3(cx, cy, cz);
    float half_b = dot(oc, d);
    float c = dot(oc, oc) - r * r;
    float discriminant = half_b * half_b - a * c;
    if (discriminant < 0.0f) return -1.0f;
    float sqrtd = fsqrt(discriminant);
#if VN_FAST_DEVICE
    float root = (-half_b - sqrtd) * inv_a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) * inv_a;
        if (root < t_min || t_max < root) return -1.0f;
    }
#else
    (void)inv_a;
    // The IEEE quotient is only formed when its outcome is open: a numerator below 0.999 * t_min * a gives a root that is
    // certainly < t_min (the true quotient is below t_min by more than the quotient's own rounding error), so the branch the
    // reference takes is known without dividing.  That is the common case for a ray leaving the very sphere it is tested against
    // (both numerators ~ 0: every bounce off the ground tests the ground again), where the division would also drop into its
    // slow path for a zero / tiny numerator (ncu: 2.4 % of all warp instructions).  Same roots, same accept / reject decisions.
    const float lo = (0.999f * t_min) * a;
    float num = -half_b - sqrtd;
    float root = -1.0f;
    bool take = false;
    if (num >= lo) { root = num / a; take = !(root < t_min || t_max < root); }
    if (!take) {
        num = -half_b + sqrtd;
        if (!(num >= lo)) return -1.0f;
        root = num / a;
        if (root < t_min || t_max < root) return -1.0f;
    }
#endif
    return root;
}

// The hit-point gate (oracle/oracle.cpp::hit_gate is the checker's copy).  The reference's intersection program
// only runs for rays that reach the sphere's AABB (custom primitives, Renderer.h:187-200), and far from the origin the float quadratic
// above "hits" spheres the ray misses by a good fraction of their radius (discriminant noise ~1e-7 |o - c|^2).  Which of those phantom
// hits survive would depend on the boxes of whatever BVH sits in front of the test; the gate makes it a function of (ray, sphere)
// alone: a root only counts when the hit point of RayTracer.cu:256 lies inside the sphere's box grown by 0.5 % of the radius plus
// 2^-19 of the coordinates' magnitude.  The builder's leaf boxes (lbvh_core.cuh::leaf_pad) contain that box with room for the slab
// test's rounding, so no BVH node can cull a root the gate would accept.  Applied by the kernels that traverse scenes from L2 / HBM;
// for scenes small enough for the shared-memory wide nodes it never fires (tests) and the headline kernel does not spend the
// instructions.
VN_HD bool hit_gate_ok(f3 o, f3 d, float t, float cx, float cy, float cz, float r) {
    const f3 p = o + d * t;                                // RayTracer.cu:256
    const float g = fabsf(r) * 1.005f + (fabsf(cx) + fabsf(cy) + fabsf(cz) + fabsf(r)) * 1.9073486328125e-06f;
    return fabsf(p.x - cx) <= g && fabsf(p.y - cy) <= g && fabsf(p.z - cz) <= g;
}

// hit point + face-forwarded normal: RayTracer.cu:256-258, 219-224
VN_HD void hit_frame(f3 o, f3 d, float t, float cx, float cy, float cz, float r, f3& p, f3& n, bool& front) {
    p = o + d * t;
    float inv = rcp(r);                                   // vec_math.h:483-487: (p - c) / r = (p - c) * (1/r)
    f3 normal = (p - mk3(cx, cy, cz)) * inv;
    front = dot(d, normal) < 0.0f;
    n = front ? normal : -normal;
}

// ---- materials
// near_zero (RayTracer.cu:8-13) compares fabs(float) < 1e-8 in double.  For a float x that is exactly
// x <= 9.99999993922529e-09f, the largest float below 1e-8 (checked in tests), so no FP64 is needed.
VN_HD bool near_zero(f3 e) {
    const float s = 9.99999993922529e-09f;
    return (fabsf(e.x) <= s) && (fabsf(e.y) <= s) && (fabsf(e.z) <= s);
}
// __closesthit__lambertian, RayTracer.cu:288-291, given the rejection-sampled point in the unit sphere
VN_HD f3 scatter_lambertian(f3 n, f3 in_unit_sphere) {
    f3 dir = n + normalize(in_unit_sphere);
    return near_zero(dir) ? n : dir;
}

// __closesthit__metal, RayTracer.cu:338-341, given normalize(direction) and the rejection-sampled point.
// Returns false when the ray is absorbed.
VN_HD bool scatter_metal(f3 unit_direction, f3 n, float fuzz, f3 in_unit_sphere, f3& dir_out) {
    f3 reflected = reflect(unit_direction, n);
    dir_out = reflected + fuzz * in_unit_sphere;
    return dot(dir_out, n) > 0.0f;
}

// reflectance, RayTracer.cu:373-379 (__powf there).  r0 = ((1 - ref_idx) / (1 + ref_idx))^2 only depends on the sphere, see below.
VN_HD float schlick_r0(float ref_idx) {
    float r0 = fdiv(1.0f - ref_idx, 1.0f + ref_idx);
    return r0 * r0;
}
VN_HD float reflectance_r0(float cosine, float r0) {
#if VN_FAST_DEVICE
    float x = 1.0f - cosine;
    float x2 = x * x;
    return r0 + (1.0f - r0) * (x2 * x2 * x);
#else
    // powf(1 - cosine, 5): x^5 formed in double and rounded once, i.e. the correctly rounded power (glibc's powf in the
    // oracle is within 1 ulp of it; the reference itself uses the approximate __powf)
    const double x = (double)(1.0f - cosine);
    const double x2 = x * x;
    return r0 + (1.0f - r0) * (float)(x2 * x2 * x);
#endif
}
VN_HD float reflectance(float cosine, float ref_idx) { return reflectance_r0(cosine, schlick_r0(ref_idx)); }

// What __closesthit__dielectric derives from the sphere's index of refraction alone, formed once per sphere by the BVH builder with
// the very operations of RayTracer.cu:398-401,375-376 (IEEE division) and stored in the sphere's material record:
// {ir, 1/ir, r0(1/ir) for a front-face hit, r0(ir) for a back-face hit}.  Per hit that removes two IEEE divisions from a branch
// that runs at 2-3 active lanes.
struct DielectricConsts { float ir, inv_ir, r0_front, r0_back; };
VN_HD DielectricConsts dielectric_consts(float ir) {
    DielectricConsts c;
    c.ir = ir;
    c.inv_ir = 1.0f / ir;
    c.r0_front = schlick_r0(c.inv_ir);
    c.r0_back = schlick_r0(ir);
    return c;
}

// __closesthit__dielectric, RayTracer.cu:398-414
VN_HD f3 scatter_dielectric(f3 unit_direction, f3 n, bool front, const DielectricConsts& dc, uint32_t& seed) {
    const float refraction_ratio = front ? dc.inv_ir : dc.ir;
#if VN_FAST_DEVICE
    float cos_theta = fminf(dot(-unit_direction, n), 1.0f);
    float sin2 = 1.0f - cos_theta * cos_theta;
    bool cannot_refract = refraction_ratio * refraction_ratio * sin2 > 1.0f;
    float cos_f = cos_theta;
#else
    const float cos_f = fminf(dot(-unit_direction, n), 1.0f);
    // ratio * sin_theta > 1.0 in double (RayTracer.cu:404-408).  sin_theta = sqrt(1 - cos^2) <= 1 and the product of two doubles
    // <= 1 rounds to <= 1, so a ratio <= 1 (entering glass) can never refuse to refract: the FP64 square root is only needed for
    // the other case.
    bool cannot_refract = false;
    if (refraction_ratio > 1.0f) {
        const double cos_theta = (double)cos_f;
        const double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
        cannot_refract = (double)refraction_ratio * sin_theta > 1.0;
    }
#endif
    // `cannot_refract || reflectance > random_float` short-circuits: no draw when cannot_refract
    bool do_reflect = cannot_refract;
    if (!do_reflect) do_reflect = reflectance_r0(cos_f, front ? dc.r0_front : dc.r0_back) > rnd(seed);
    return do_reflect ? reflect(unit_direction, n) : refract(unit_direction, n, refraction_ratio);
}
VN_HD f3 scatter_dielectric(f3 unit_direction, f3 n, bool front, float ir, uint32_t& seed) {
    return scatter_dielectric(unit_direction, n, front, dielectric_consts(ir), seed);
}

// __miss__ms, RayTracer.cu:442-450.  0.5*(y+1.0) in double then narrowed == the float expression below
// (y+1 is exact in double; halving is exact; one rounding either way).
VN_HD f3 sky(f3 unit_direction) {
    float t = 0.5f * (unit_direction.y + 1.0f);
    return lerp3(mk3(1.0f), mk3(0.5f, 0.7f, 1.0f), t);
}

// ---- toSRGB / quantizeUnsigned8Bits / make_color, RayTracer.cu:16-47
VN_HD float to_srgb(float c) {
    const float invGamma = 1.0f / 2.4f;
    float powed = powf(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}
VN_HD uint32_t quantize8(float x) {
    x = clamp01(x);
    uint32_t v = (uint32_t)(x * 256.0f);
    return v < 255u ? v : 255u;
}
VN_HD uint32_t make_color_u32(f3 c) {   // uchar4 {r,g,b,255} packed little-endian
    uint32_t r = quantize8(to_srgb(clamp01(c.x)));
    uint32_t g = quantize8(to_srgb(clamp01(c.y)));
    uint32_t b = quantize8(to_srgb(clamp01(c.z)));
    return r | (g << 8) | (b << 16) | (255u << 24);
}

// ---- BVH traversal over packed 32-byte nodes (see lbvh.cu for the layout)
//   node i = nodes[2i] = {lo.xyz, link}, nodes[2i+1] = {hi.xyz, aux}; children of an internal node are the
//   64-byte aligned pair (link, link+1); leaf link = 0x80000000 | first<<3 | (count-1).
struct alignas(16) f4 { float x, y, z, w; };
#if defined(__CUDACC__)
typedef float4 node_f4;
#else
typedef f4 node_f4;
#endif

constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr uint32_t kEmptyScene = 0xFFFFFFFFu;
constexpr int kStackSize = 64;
constexpr int kWideStackSize = 128;      // global wide nodes of large scenes: 3 pushes per level, Karras trees of 16 M spheres are ~40 levels deep
constexpr float kTMin = 0.001f;   // RayTracer.cu:194
constexpr float kTMax = 1e16f;    // RayTracer.cu:195

VN_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}

struct TraceCounters { uint32_t nodes, spheres; };

// Slab test against one child box.  Not parity-relevant (it only has to be conservative; boxes are padded at build
// time), so it uses explicit FMAs in both builds; idir comes from slab_idir(), so all products are finite.
VN_HD bool box_hit(const node_f4& lo, const node_f4& hi, f3 idir, f3 ood, float tbest, float& tnear) {
    float t0x = fmaf(lo.x, idir.x, -ood.x), t1x = fmaf(hi.x, idir.x, -ood.x);
    float t0y = fmaf(lo.y, idir.y, -ood.y), t1y = fmaf(hi.y, idir.y, -ood.y);
    float t0z = fmaf(lo.z, idir.z, -ood.z), t1z = fmaf(hi.z, idir.z, -ood.z);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tbest));
    tnear = tn;
    return tn <= tf;
}

// Reciprocal direction for the slab tests.  A component that is exactly (or nearly) zero would give idir = inf and
// o*idir - plane*idir = inf - inf = NaN, which the NaN-dropping min/max could turn into a wrongly culled box; keeping
// |d| >= 1e-30 keeps every product finite and the test conservative.  The slab test only has to be conservative (the
// boxes are padded), so the reciprocal itself is the approximate MUFU.RCP in both builds.
// Approximate reciprocal for the (conservative) slab test: one MUFU.RCP.  __fdividef(1, x) without -ftz wraps the MUFU in a
// denormal-rescue sequence (6 instructions); the operands here are never denormal (|d| >= 1e-30 below, tbest >= 1e-3).
VN_HD float rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
VN_HD f3 slab_idir(f3 d) {
    const float tiny = 1e-30f;
    const float dx = fabsf(d.x) < tiny ? copysignf(tiny, d.x) : d.x;
    const float dy = fabsf(d.y) < tiny ? copysignf(tiny, d.y) : d.y;
    const float dz = fabsf(d.z) < tiny ? copysignf(tiny, d.z) : d.z;
    return mk3(rcp_approx(dx), rcp_approx(dy), rcp_approx(dz));
}

// Same test against a node stored in NEAR/FAR-plane form for the ray's direction octant (see k_render_persistent:
// when shared memory allows, the nodes are staged 8 times, once per octant, with lo/hi already swapped per axis).
// The six per-axis min/max of box_hit disappear: 6 FFMA + 2 FMNMX3 + 2 FMNMX per box instead of 6 FFMA + 10 min/max.
VN_HD bool box_hit_oct(const node_f4& nr, const node_f4& fr, f3 idir, f3 ood, float tbest, float& tnear) {
    const float tn = fmaxf(fmaxf(fmaf(nr.x, idir.x, -ood.x), fmaf(nr.y, idir.y, -ood.y)), fmaxf(fmaf(nr.z, idir.z, -ood.z), 0.0f));
    const float tf = fminf(fminf(fmaf(fr.x, idir.x, -ood.x), fmaf(fr.y, idir.y, -ood.y)), fminf(fmaf(fr.z, idir.z, -ood.z), tbest));
    tnear = tn;
    return tn <= tf;
}
// direction octant: bit a set when component a has its sign bit set (so -0 pairs with idir = -inf)
VN_HD uint32_t ray_octant(f3 d) { return (f2u(d.x) >> 31) | ((f2u(d.y) >> 31) << 1) | ((f2u(d.z) >> 31) << 2); }

// Closest hit in [kTMin, kTMax] = optixTrace(...) of RayTracer.cu:190-202.  prim = index into the SORTED sphere
// arrays, -1 on miss.
template <bool kCount, bool kOct = false>
VN_HD void closest_hit(const node_f4* __restrict__ nodes, const node_f4* __restrict__ geom, uint32_t root_link,
                       f3 o, f3 d, float& t_out, int& prim_out, TraceCounters& cnt, uint32_t oct_stride = 0, bool gate = false) {
    float tbest = kTMax;
    int prim = -1;
    {
        const f3 idir = slab_idir(d);
        if (kOct) nodes += ray_octant(d) * oct_stride;
        const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        const float a = dot(d, d);
        const float inv_a = rcp(a);
        uint32_t stack[kStackSize];
        int sp = 0;
        uint32_t cur = root_link;
        // while-while: every lane first descends through internal nodes until it holds a leaf (or is done); the warp
        // reconverges at the end of the inner loop, so the sphere tests below run with many lanes active instead of
        // being interleaved, a few lanes at a time, with other lanes' box tests.  kEmptyScene doubles as "done".
        for (;;) {
            while (!(cur & kLeafFlag)) {
                const node_f4 l0 = nodes[2 * cur], l1 = nodes[2 * cur + 1];
                const node_f4 r0 = nodes[2 * cur + 2], r1 = nodes[2 * cur + 3];
                if (kCount) cnt.nodes += 1;
                float tl, tr;
                const bool hl = kOct ? box_hit_oct(l0, l1, idir, ood, tbest, tl) : box_hit(l0, l1, idir, ood, tbest, tl);
                const bool hr = kOct ? box_hit_oct(r0, r1, idir, ood, tbest, tr) : box_hit(r0, r1, idir, ood, tbest, tr);
                const uint32_t ll = f2u(l0.w), lr = f2u(r0.w);
                if (hl && hr) {
                    const bool left_first = tl <= tr;
                    cur = left_first ? ll : lr;
                    stack[sp++] = left_first ? lr : ll;
                } else if (hl) {
                    cur = ll;
                } else if (hr) {
                    cur = lr;
                } else {
                    cur = sp ? stack[--sp] : kEmptyScene;
                }
            }
            if (cur == kEmptyScene) break;
            const uint32_t first = (cur & 0x7FFFFFFFu) >> 3;
            const uint32_t count = (cur & 7u) + 1u;
            for (uint32_t k = 0; k < count; k++) {
                const node_f4 g = geom[first + k];
                if (kCount) cnt.spheres += 1;
                const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, g.x, g.y, g.z, g.w))) { tbest = t; prim = (int)(first + k); }
            }
            cur = sp ? stack[--sp] : kEmptyScene;
        }
    }
    t_out = tbest;
    prim_out = prim;
}

// Closest hit over the 4-wide, octant-specialised shared-memory nodes (lbvh_core.cuh::wide_octant_node): one step
// = 7 x 128-bit loads + four slab tests; hit children are pushed far to near in the octant's static order, so the
// step has no distance compare.  Same leaves, same sphere_root() as closest_hit(): the result is identical.
constexpr uint32_t kWideNodeF4 = 7;
VN_HD bool slab_hit(float nx, float ny, float nz, float fx, float fy, float fz, f3 idir, f3 ood, float tbest) {
    const float tn = fmaxf(fmaxf(fmaf(nx, idir.x, -ood.x), fmaf(ny, idir.y, -ood.y)), fmaxf(fmaf(nz, idir.z, -ood.z), 0.0f));
    const float tf = fminf(fminf(fmaf(fx, idir.x, -ood.x), fmaf(fy, idir.y, -ood.y)), fminf(fmaf(fz, idir.z, -ood.z), tbest));
    return tn <= tf;
}
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void push_if(uint32_t* stack, int& sp, uint32_t v, bool pred) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.local.u32 [%0], %1; }" ::"r"((uint32_t)__cvta_generic_to_local(stack + sp)), "r"(v), "r"((uint32_t)pred) : "memory");
    sp += pred ? 1 : 0;
}
#endif
#if defined(__CUDACC__)
// Two slab planes per instruction: Blackwell's packed FFMA2 (fma.rn.f32x2) computes {a0*b + c, a1*b + c} with one issue slot; each
// half is the IEEE fma of the scalar form, so the traversal visits exactly the same nodes as the host emulation's fmaf().  The four
// children's planes arrive as one float4 per axis (LDS.128 fills two aligned register pairs), the ray's reciprocal direction and
// -origin/direction ride along as broadcast operands.
__device__ __forceinline__ void fma2_bcast(float a0, float a1, float b, float c, float& r0, float& r1) {
    // one asm block: the {b, b} / {c, c} packs stay next to the fma, where ptxas folds them into FFMA2's broadcast operands
    // (as separate statements they are hoisted out of the traversal loop and cost four more live registers)
    asm("{ .reg .b64 pa, pb, pc, pd;\n\t"
        "mov.b64 pa, {%2, %3};\n\t"
        "mov.b64 pb, {%4, %4};\n\t"
        "mov.b64 pc, {%5, %5};\n\t"
        "fma.rn.f32x2 pd, pa, pb, pc;\n\t"
        "mov.b64 {%0, %1}, pd; }"
        : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void slab4(const node_f4& q, float id, float nood, float& r0, float& r1, float& r2, float& r3) {
    fma2_bcast(q.x, q.y, id, nood, r0, r1);
    fma2_bcast(q.z, q.w, id, nood, r2, r3);
}
#endif
// One step over a 4-wide node: returns the next node / leaf link (kEmptyScene when the ray is finished).
VN_HD uint32_t wide_node_step(const node_f4* __restrict__ wn, uint32_t cur, f3 idir, f3 ood, float tbest, uint32_t* stack, int& sp) {
    const node_f4* __restrict__ p = wn + kWideNodeF4 * cur;
    const node_f4 nx = p[0], ny = p[1], nz = p[2], fx = p[3], fy = p[4], fz = p[5], lk = p[6];
#if defined(__CUDA_ARCH__)
    float ax[4], ay[4], az[4], bx[4], by[4], bz[4];
    slab4(nx, idir.x, -ood.x, ax[0], ax[1], ax[2], ax[3]);
    slab4(ny, idir.y, -ood.y, ay[0], ay[1], ay[2], ay[3]);
    slab4(nz, idir.z, -ood.z, az[0], az[1], az[2], az[3]);
    slab4(fx, idir.x, -ood.x, bx[0], bx[1], bx[2], bx[3]);
    slab4(fy, idir.y, -ood.y, by[0], by[1], by[2], by[3]);
    slab4(fz, idir.z, -ood.z, bz[0], bz[1], bz[2], bz[3]);
    bool h[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float tn = fmaxf(fmaxf(ax[c], ay[c]), fmaxf(az[c], 0.0f));
        const float tf = fminf(fminf(bx[c], by[c]), fminf(bz[c], tbest));
        h[c] = tn <= tf;
    }
    const bool h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
#else
    const bool h0 = slab_hit(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, idir, ood, tbest);
    const bool h1 = slab_hit(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, idir, ood, tbest);
    const bool h2 = slab_hit(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, idir, ood, tbest);
    const bool h3 = slab_hit(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, idir, ood, tbest);
#endif
    // the nearest hit child (static octant order) becomes the current node straight from registers; the others are
    // pushed far to near.  Only a step without any hit pays the stack load.
    const bool c3 = h3 && (h0 || h1 || h2), c2 = h2 && (h0 || h1), c1 = h1 && h0;
#if defined(__CUDA_ARCH__)
    // predicated stores written in PTX: left to itself the compiler turns the three pushes into six divergent branches
    push_if(stack, sp, f2u(lk.w), c3);
    push_if(stack, sp, f2u(lk.z), c2);
    push_if(stack, sp, f2u(lk.y), c1);
#else
    if (c3) stack[sp++] = f2u(lk.w);
    if (c2) stack[sp++] = f2u(lk.z);
    if (c1) stack[sp++] = f2u(lk.y);
#endif
    uint32_t next = h0 ? f2u(lk.x) : (h1 ? f2u(lk.y) : (h2 ? f2u(lk.z) : f2u(lk.w)));
    if (!(h0 || h1 || h2 || h3)) next = sp ? stack[--sp] : kEmptyScene;
    return next;
}
// ---- the same step with the ray parameter NORMALISED by the closest hit so far (s = t / tbest) and saturating slab products.
// The min/max, compare and select instructions of a node step run on the SM's half-rate ALU pipe, and the step above holds 39 of
// them against 12 FFMA2 + 8 loads: the ALU pipe, not the issue slot, bounds it (ncu: alu 63 %, issue 80 %).  In normalised units
// the valid interval of a slab is [0, 1], which is what FFMA.SAT clamps to for free: with the z planes saturated,
//     tn = max(ax, ay, sat(az)) >= 0        tf = min(bx, by, sat(bz)) <= 1
// need no clamp against 0 and tbest (8 FMNMX per step gone).  A box beyond tbest has tn >= 1 >= tf and a box behind the origin
// has tf <= 0 <= tn, so the test is the STRICT tn < tf (an exact tie of two independently rounded products only occurs for a box
// the ray grazes; boxes are padded, so a sphere that is hit is never lost).  Like slab_hit() this only has to be conservative;
// the scale factors are refreshed whenever a leaf shortens tbest.
struct SlabScale { f3 sdir, nsood; };            // idir / tbest,  -(o * idir) / tbest
VN_HD SlabScale slab_scale(f3 idir, f3 ood, float tbest) {
    const float inv_t = rcp_approx(tbest);
    SlabScale r;
    r.sdir = mk3(idir.x * inv_t, idir.y * inv_t, idir.z * inv_t);
    r.nsood = mk3(-(ood.x * inv_t), -(ood.y * inv_t), -(ood.z * inv_t));
    return r;
}
VN_HD float fma_sat(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
#else
    return fminf(fmaxf(fmaf(a, b, c), 0.0f), 1.0f);          // NaN -> 0 like .sat
#endif
}
// The octant's node copy as the traversal loop sees it.  On the GPU it is a 32-bit shared-memory byte address laundered through an
// empty asm: left as pointer arithmetic, the compiler re-derives it from the ray direction inside the loop (7 ALU instructions per
// step) rather than keep it in a register.  Loads are explicit ld.shared.v4 with immediate offsets: one IMAD + 7 LDS.128 per step.
#if defined(__CUDACC__)
struct WideBase { uint32_t addr; };
__device__ __forceinline__ WideBase wide_base(const node_f4* wnodes, uint32_t octant, uint32_t oct_stride) {
    WideBase b;
    b.addr = (uint32_t)__cvta_generic_to_shared(wnodes + (size_t)octant * oct_stride);
    asm volatile("" : "+r"(b.addr));
    return b;
}
template <int kOff>
__device__ __forceinline__ node_f4 lds_f4(uint32_t addr) {
    node_f4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(kOff));
    return v;
}
#else
struct WideBase { const node_f4* p; };
inline WideBase wide_base(const node_f4* wnodes, uint32_t octant, uint32_t oct_stride) { WideBase b; b.p = wnodes + (size_t)octant * oct_stride; return b; }
#endif
#if !defined(__CUDACC__)
// host emulation of the step (tests/host_harness.cpp); the GPU runs wide_node_step_dev() below: same arithmetic, same order
inline uint32_t wide_node_step_sat(WideBase wb, uint32_t cur, const SlabScale& sc, uint32_t* stack, int& sp) {
    const node_f4* __restrict__ p = wb.p + kWideNodeF4 * cur;
    const node_f4 nx = p[0], ny = p[1], nz = p[2], fx = p[3], fy = p[4], fz = p[5], lk = p[6];
    float ax[4], ay[4], bx[4], by[4];
#if defined(__CUDA_ARCH__)
    slab4(nx, sc.sdir.x, sc.nsood.x, ax[0], ax[1], ax[2], ax[3]);
    slab4(ny, sc.sdir.y, sc.nsood.y, ay[0], ay[1], ay[2], ay[3]);
    slab4(fx, sc.sdir.x, sc.nsood.x, bx[0], bx[1], bx[2], bx[3]);
    slab4(fy, sc.sdir.y, sc.nsood.y, by[0], by[1], by[2], by[3]);
#else
    ax[0] = fmaf(nx.x, sc.sdir.x, sc.nsood.x); ax[1] = fmaf(nx.y, sc.sdir.x, sc.nsood.x); ax[2] = fmaf(nx.z, sc.sdir.x, sc.nsood.x); ax[3] = fmaf(nx.w, sc.sdir.x, sc.nsood.x);
    ay[0] = fmaf(ny.x, sc.sdir.y, sc.nsood.y); ay[1] = fmaf(ny.y, sc.sdir.y, sc.nsood.y); ay[2] = fmaf(ny.z, sc.sdir.y, sc.nsood.y); ay[3] = fmaf(ny.w, sc.sdir.y, sc.nsood.y);
    bx[0] = fmaf(fx.x, sc.sdir.x, sc.nsood.x); bx[1] = fmaf(fx.y, sc.sdir.x, sc.nsood.x); bx[2] = fmaf(fx.z, sc.sdir.x, sc.nsood.x); bx[3] = fmaf(fx.w, sc.sdir.x, sc.nsood.x);
    by[0] = fmaf(fy.x, sc.sdir.y, sc.nsood.y); by[1] = fmaf(fy.y, sc.sdir.y, sc.nsood.y); by[2] = fmaf(fy.z, sc.sdir.y, sc.nsood.y); by[3] = fmaf(fy.w, sc.sdir.y, sc.nsood.y);
#endif
    const float az[4] = {fma_sat(nz.x, sc.sdir.z, sc.nsood.z), fma_sat(nz.y, sc.sdir.z, sc.nsood.z), fma_sat(nz.z, sc.sdir.z, sc.nsood.z), fma_sat(nz.w, sc.sdir.z, sc.nsood.z)};
    const float bz[4] = {fma_sat(fz.x, sc.sdir.z, sc.nsood.z), fma_sat(fz.y, sc.sdir.z, sc.nsood.z), fma_sat(fz.z, sc.sdir.z, sc.nsood.z), fma_sat(fz.w, sc.sdir.z, sc.nsood.z)};
    bool h[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 4; c++) h[c] = fmaxf(fmaxf(ax[c], ay[c]), az[c]) < fminf(fminf(bx[c], by[c]), bz[c]);
    const bool h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3];
    const bool c3 = h3 && (h0 || h1 || h2), c2 = h2 && (h0 || h1), c1 = h1 && h0;
#if defined(__CUDA_ARCH__)
    push_if(stack, sp, f2u(lk.w), c3);
    push_if(stack, sp, f2u(lk.z), c2);
    push_if(stack, sp, f2u(lk.y), c1);
#else
    if (c3) stack[sp++] = f2u(lk.w);
    if (c2) stack[sp++] = f2u(lk.z);
    if (c1) stack[sp++] = f2u(lk.y);
#endif
    uint32_t next = h0 ? f2u(lk.x) : (h1 ? f2u(lk.y) : (h2 ? f2u(lk.z) : f2u(lk.w)));
    if (!(h0 || h1 || h2 || h3)) next = sp ? stack[--sp] : kEmptyScene;
    return next;
}
#endif
#if defined(__CUDACC__)
// Device form of the step: the traversal stack is addressed by its local-memory byte address (`top` = first free word, `base` =
// bottom), and the select / push / pop logic is written out in PTX: four compares, three predicated stores whose addresses advance
// by 4 * hit, and -- only when the nearest child (slot 0 of the octant order) was missed -- a predicated pop of the entry just
// pushed (or of an older one).  20 instructions where the compiler's version of `h3 && (h0 || h1 || h2)` etc. took 33, 14 of them on
// the ALU pipe instead of 23.  Same visiting order as wide_node_step(): hit children far to near onto the stack, nearest first.
__device__ __forceinline__ uint32_t wide_node_step_dev(WideBase wb, uint32_t cur, const SlabScale& sc, uint32_t& top, uint32_t base) {
    const uint32_t pa = wb.addr + cur * (kWideNodeF4 * 16u);
    const node_f4 nx = lds_f4<0>(pa), ny = lds_f4<16>(pa), nz = lds_f4<32>(pa), fx = lds_f4<48>(pa), fy = lds_f4<64>(pa), fz = lds_f4<80>(pa), lk = lds_f4<96>(pa);
    float ax[4], ay[4], bx[4], by[4];
    slab4(nx, sc.sdir.x, sc.nsood.x, ax[0], ax[1], ax[2], ax[3]);
    slab4(ny, sc.sdir.y, sc.nsood.y, ay[0], ay[1], ay[2], ay[3]);
    slab4(fx, sc.sdir.x, sc.nsood.x, bx[0], bx[1], bx[2], bx[3]);
    slab4(fy, sc.sdir.y, sc.nsood.y, by[0], by[1], by[2], by[3]);
    const float az[4] = {fma_sat(nz.x, sc.sdir.z, sc.nsood.z), fma_sat(nz.y, sc.sdir.z, sc.nsood.z), fma_sat(nz.z, sc.sdir.z, sc.nsood.z), fma_sat(nz.w, sc.sdir.z, sc.nsood.z)};
    const float bz[4] = {fma_sat(fz.x, sc.sdir.z, sc.nsood.z), fma_sat(fz.y, sc.sdir.z, sc.nsood.z), fma_sat(fz.z, sc.sdir.z, sc.nsood.z), fma_sat(fz.w, sc.sdir.z, sc.nsood.z)};
    float tn[4], tf[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { tn[c] = fmaxf(fmaxf(ax[c], ay[c]), az[c]); tf[c] = fminf(fminf(bx[c], by[c]), bz[c]); }
    uint32_t next;
    asm volatile(
        "{\n\t"
        ".reg .pred h0, h1, h2, h3, q;\n\t"
        ".reg .u32 n;\n\t"
        "setp.lt.f32 h3, %2, %3;\n\t"
        "setp.lt.f32 h2, %4, %5;\n\t"
        "setp.lt.f32 h1, %6, %7;\n\t"
        "setp.lt.f32 h0, %8, %9;\n\t"
        "@h3 st.local.u32 [%1], %10;\n\t"
        "selp.u32 n, 4, 0, h3;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "@h2 st.local.u32 [%1], %11;\n\t"
        "selp.u32 n, 4, 0, h2;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "@h1 st.local.u32 [%1], %12;\n\t"
        "selp.u32 n, 4, 0, h1;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "setp.gt.u32 q, %1, %14;\n\t"
        "and.pred q, q, !h0;\n\t"
        "selp.u32 %0, %13, 0xFFFFFFFF, h0;\n\t"
        "@q ld.local.u32 %0, [%1+-4];\n\t"
        "@q add.u32 %1, %1, -4;\n\t"
        "}"
        : "=r"(next), "+r"(top)
        : "f"(tn[3]), "f"(tf[3]), "f"(tn[2]), "f"(tf[2]), "f"(tn[1]), "f"(tf[1]), "f"(tn[0]), "f"(tf[0]),
          "r"(f2u(lk.w)), "r"(f2u(lk.z)), "r"(f2u(lk.y)), "r"(f2u(lk.x)), "r"(base)
        : "memory");
    return next;
}
__device__ __forceinline__ uint32_t stack_pop_dev(uint32_t& top, uint32_t base) {
    uint32_t next;
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.gt.u32 q, %1, %2;\n\t"
        "mov.u32 %0, 0xFFFFFFFF;\n\t"
        "@q ld.local.u32 %0, [%1+-4];\n\t"
        "@q add.u32 %1, %1, -4;\n\t"
        "}"
        : "=r"(next), "+r"(top) : "r"(base) : "memory");
    return next;
}
#endif
// ---- 16-bit links (k_render_lean, path_kernels.cu).  The traversal stack of the kernels above lives in local memory: 64 words per
// lane, i.e. one 128-byte line per warp and stack level, next to 227 KB of shared memory that leave the SM 28 KB of L1 -- ncu
// (profiles/r01s3_path_kernel_ncu.txt) showed the warps waiting for those loads more than for anything else (long_scoreboard 2.7
// warps per issue cycle; 24 % of the loads missed L1).  A scene whose wide nodes fit in shared memory has < 32768 nodes and < 4096
// spheres, so a link fits 16 bits: bit 15 = leaf (or done), leaf = 0x8000 | first << 3 | (count - 1), done / empty = 0xFFFF.  Two
// stack entries then share a word, the stack's footprint in L1 halves, and with kDone16 parked at the bottom of the stack a pop
// needs no emptiness test: popping the sentinel IS "traversal finished".
constexpr uint32_t kLeaf16 = 0x8000u;
constexpr uint32_t kDone16 = 0xFFFFu;
constexpr uint32_t kLink16MaxPrims = 4096u;
VN_HD uint32_t link16(uint32_t link) {
    if (link == kEmptyScene) return kDone16;
    if (link & kLeafFlag) return kLeaf16 | (link & 0x7FFFu);         // first << 3 | (count - 1), first < 4096
    return link;                                                      // wide-node index
}
#if defined(__CUDACC__)
// wide_node_step_dev() with 16-bit links and the next node chosen in registers: the nearest hit child (slot order = the octant's
// front-to-back order) becomes the current node directly, only the OTHER hit children are pushed (far to near), and the stack is read
// only when no child was hit.  The step above pushed every hit child but slot 0 and popped the nearest one straight back whenever
// slot 0 was missed -- 70 % of all steps went through a dependent store/load pair in local memory.  Same visiting order.
__device__ __forceinline__ uint32_t wide_node_step16_dev(WideBase wb, uint32_t cur, const SlabScale& sc, uint32_t& top, uint32_t& tos) {
    const uint32_t pa = wb.addr + cur * (kWideNodeF4 * 16u);
    const node_f4 nx = lds_f4<0>(pa), ny = lds_f4<16>(pa), nz = lds_f4<32>(pa), fx = lds_f4<48>(pa), fy = lds_f4<64>(pa), fz = lds_f4<80>(pa), lk = lds_f4<96>(pa);
    float ax[4], ay[4], bx[4], by[4];
    slab4(nx, sc.sdir.x, sc.nsood.x, ax[0], ax[1], ax[2], ax[3]);
    slab4(ny, sc.sdir.y, sc.nsood.y, ay[0], ay[1], ay[2], ay[3]);
    slab4(fx, sc.sdir.x, sc.nsood.x, bx[0], bx[1], bx[2], bx[3]);
    slab4(fy, sc.sdir.y, sc.nsood.y, by[0], by[1], by[2], by[3]);
    const float az[4] = {fma_sat(nz.x, sc.sdir.z, sc.nsood.z), fma_sat(nz.y, sc.sdir.z, sc.nsood.z), fma_sat(nz.z, sc.sdir.z, sc.nsood.z), fma_sat(nz.w, sc.sdir.z, sc.nsood.z)};
    const float bz[4] = {fma_sat(fz.x, sc.sdir.z, sc.nsood.z), fma_sat(fz.y, sc.sdir.z, sc.nsood.z), fma_sat(fz.z, sc.sdir.z, sc.nsood.z), fma_sat(fz.w, sc.sdir.z, sc.nsood.z)};
    float tn[4], tf[4];
#pragma unroll
    for (int c = 0; c < 4; c++) { tn[c] = fmaxf(fmaxf(ax[c], ay[c]), az[c]); tf[c] = fminf(fminf(bx[c], by[c]), bz[c]); }
    // The newest stack entry lives in a register (`tos`), the older ones in local memory below `top`: a push stores the old top entry and
    // keeps the new one, a pop hands out the register at once and refills it with a load whose latency nobody waits for -- the loaded
    // value is first needed by the NEXT pop or push, a node step (or a sphere test) later.
    uint32_t next;
    asm volatile(
        "{\n\t"
        ".reg .pred h0, h1, h2, h3, t01, t012, c1, c2, c3, any;\n\t"
        ".reg .u32 n;\n\t"
        "setp.lt.f32 h3, %3, %4;\n\t"
        "setp.lt.f32 h2, %5, %6;\n\t"
        "setp.lt.f32 h1, %7, %8;\n\t"
        "setp.lt.f32 h0, %9, %10;\n\t"
        "or.pred t01, h0, h1;\n\t"
        "or.pred t012, t01, h2;\n\t"
        "and.pred c3, h3, t012;\n\t"
        "and.pred c2, h2, t01;\n\t"
        "and.pred c1, h1, h0;\n\t"
        "or.pred any, t012, h3;\n\t"
        "@c3 st.local.b16 [%1], %2;\n\t"           // (a 32-bit source register: the low half is stored)
        "selp.u32 %2, %11, %2, c3;\n\t"
        "selp.u32 n, 2, 0, c3;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "@c2 st.local.b16 [%1], %2;\n\t"
        "selp.u32 %2, %12, %2, c2;\n\t"
        "selp.u32 n, 2, 0, c2;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "@c1 st.local.b16 [%1], %2;\n\t"
        "selp.u32 %2, %13, %2, c1;\n\t"
        "selp.u32 n, 2, 0, c1;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "selp.u32 %0, %12, %11, h2;\n\t"
        "selp.u32 %0, %13, %0, h1;\n\t"
        "selp.u32 %0, %14, %0, h0;\n\t"
        "@!any mov.u32 %0, %2;\n\t"
        "@!any ld.local.u16 %2, [%1+-2];\n\t"      // (zero-extended into the 32-bit register)
        "@!any add.u32 %1, %1, -2;\n\t"
        "}"
        : "=&r"(next), "+r"(top), "+r"(tos)
        : "f"(tn[3]), "f"(tf[3]), "f"(tn[2]), "f"(tf[2]), "f"(tn[1]), "f"(tf[1]), "f"(tn[0]), "f"(tf[0]),
          "r"(f2u(lk.w)), "r"(f2u(lk.z)), "r"(f2u(lk.y)), "r"(f2u(lk.x))
        : "memory");
    return next;
}
// pop: the register entry is the result, the refill load runs behind it.  The bottom of the stack holds kDone16, so no emptiness test.
__device__ __forceinline__ uint32_t stack_pop16_dev(uint32_t& top, uint32_t& tos) {
    const uint32_t next = tos;
    asm volatile(
        "ld.local.u16 %0, [%1+-2];\n\t"
        "add.u32 %1, %1, -2;"
        : "=r"(tos), "+r"(top) : : "memory");
    return next;
}
#endif
#if defined(__CUDACC__)
// One step over a PAIR node fetched from L2 / HBM (k_render_lean<kGlobal>): the two children's boxes arrive as one 64-byte line, the
// nearer hit child becomes the current node, the other one is pushed; like the 16-bit stack above, the newest entry lives in a register
// and the bottom of the stack holds kEmptyScene, so a pop needs no emptiness test and nobody waits for its refill load.  Same test,
// same order as the loop of closest_hit(): the closest hit it finds is the same.
__device__ __forceinline__ uint32_t pair_node_step_dev(const node_f4* __restrict__ nodes, uint32_t cur, f3 idir, f3 ood, float tbest, uint32_t& top, uint32_t& tos) {
    // the pair = one aligned 64-byte line, fetched with two 256-bit loads (LDG.E.256, new on sm_100): half the requests and tag look-ups
    // of four 128-bit loads in an L1 that this kernel keeps 75 % busy (ncu, 1 M spheres)
    const node_f4* __restrict__ q = nodes + 2ull * cur;
    node_f4 l0, l1, r0, r1;
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(l0.x), "=f"(l0.y), "=f"(l0.z), "=f"(l0.w), "=f"(l1.x), "=f"(l1.y), "=f"(l1.z), "=f"(l1.w) : "l"(q));
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r0.x), "=f"(r0.y), "=f"(r0.z), "=f"(r0.w), "=f"(r1.x), "=f"(r1.y), "=f"(r1.z), "=f"(r1.w) : "l"(q + 2));
    float tl, tr;
    const bool hl = box_hit(l0, l1, idir, ood, tbest, tl);
    const bool hr = box_hit(r0, r1, idir, ood, tbest, tr);
    const uint32_t ll = f2u(l0.w), lr = f2u(r0.w);
    const bool left_first = tl <= tr;
    const uint32_t near = (hl && (!hr || left_first)) ? ll : lr;
    const uint32_t far = left_first ? lr : ll;
    uint32_t next = near;
    asm volatile(
        "{\n\t"
        ".reg .pred both, none;\n\t"
        ".reg .u32 n;\n\t"
        "setp.ne.u32 both, %4, 0;\n\t"
        "setp.ne.u32 none, %5, 0;\n\t"
        "@both st.local.u32 [%1], %2;\n\t"
        "selp.u32 %2, %3, %2, both;\n\t"
        "selp.u32 n, 4, 0, both;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "@none mov.u32 %0, %2;\n\t"
        "@none ld.local.u32 %2, [%1+-4];\n\t"
        "@none add.u32 %1, %1, -4;\n\t"
        "}"
        : "+r"(next), "+r"(top), "+r"(tos)
        : "r"(far), "r"((uint32_t)(hl && hr)), "r"((uint32_t)(!hl && !hr))
        : "memory");
    return next;
}
// The same step over the QUANTISED pair (lbvh.cu::k_quantize_pairs): one 256-bit load instead of two.  A plane is q * scale + root_lo, so its
// slab parameter is q * qa + qb with qa = scale * idir and qb = (root_lo - o) * idir, formed once per ray.  The boxes only got larger (one
// quantum of slack beyond the outward rounding covers the decode's own rounding), so the traversal stays conservative and the closest hit is
// the one closest_hit() finds; the ORDER of the two children can differ where their entry distances are within a quantum, which changes
// the steps a ray takes, not what it hits.
__device__ __forceinline__ uint32_t pair_node_step_q_dev(const uint4* __restrict__ q, uint32_t cur, f3 qa, f3 qb, float tbest, uint32_t& top, uint32_t& tos) {
    uint32_t w0, w1, w2, w3, w4, w5, ll, lr;
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(ll), "=r"(lr) : "l"(q + cur));
    // left: lo = (w0.lo, w0.hi, w1.lo), hi = (w1.hi, w2.lo, w2.hi); right: lo = (w3.lo, w3.hi, w4.lo), hi = (w4.hi, w5.lo, w5.hi)
    const float l0x = fmaf((float)(w0 & 0xFFFFu), qa.x, qb.x), l0y = fmaf((float)(w0 >> 16), qa.y, qb.y), l0z = fmaf((float)(w1 & 0xFFFFu), qa.z, qb.z);
    const float l1x = fmaf((float)(w1 >> 16), qa.x, qb.x), l1y = fmaf((float)(w2 & 0xFFFFu), qa.y, qb.y), l1z = fmaf((float)(w2 >> 16), qa.z, qb.z);
    const float r0x = fmaf((float)(w3 & 0xFFFFu), qa.x, qb.x), r0y = fmaf((float)(w3 >> 16), qa.y, qb.y), r0z = fmaf((float)(w4 & 0xFFFFu), qa.z, qb.z);
    const float r1x = fmaf((float)(w4 >> 16), qa.x, qb.x), r1y = fmaf((float)(w5 & 0xFFFFu), qa.y, qb.y), r1z = fmaf((float)(w5 >> 16), qa.z, qb.z);
    const float tl = fmaxf(fmaxf(fminf(l0x, l1x), fminf(l0y, l1y)), fmaxf(fminf(l0z, l1z), 0.0f));
    const float fl = fminf(fminf(fmaxf(l0x, l1x), fmaxf(l0y, l1y)), fminf(fmaxf(l0z, l1z), tbest));
    const float tr = fmaxf(fmaxf(fminf(r0x, r1x), fminf(r0y, r1y)), fmaxf(fminf(r0z, r1z), 0.0f));
    const float fr = fminf(fminf(fmaxf(r0x, r1x), fmaxf(r0y, r1y)), fminf(fmaxf(r0z, r1z), tbest));
    const bool hl = tl <= fl, hr = tr <= fr;
    const bool left_first = tl <= tr;
    const uint32_t near = (hl && (!hr || left_first)) ? ll : lr;
    const uint32_t far = left_first ? lr : ll;
    uint32_t next = near;
    asm volatile(
        "{\n\t"
        ".reg .pred both, none;\n\t"
        ".reg .u32 n;\n\t"
        "setp.ne.u32 both, %4, 0;\n\t"
        "setp.ne.u32 none, %5, 0;\n\t"
        "@both st.local.u32 [%1], %2;\n\t"
        "selp.u32 %2, %3, %2, both;\n\t"
        "selp.u32 n, 4, 0, both;\n\t"
        "add.u32 %1, %1, n;\n\t"
        "@none mov.u32 %0, %2;\n\t"
        "@none ld.local.u32 %2, [%1+-4];\n\t"
        "@none add.u32 %1, %1, -4;\n\t"
        "}"
        : "+r"(next), "+r"(top), "+r"(tos)
        : "r"(far), "r"((uint32_t)(hl && hr)), "r"((uint32_t)(!hl && !hr))
        : "memory");
    return next;
}
__device__ __forceinline__ uint32_t stack_pop32_dev(uint32_t& top, uint32_t& tos) {
    const uint32_t next = tos;
    asm volatile(
        "ld.local.u32 %0, [%1+-4];\n\t"
        "add.u32 %1, %1, -4;"
        : "=r"(tos), "+r"(top) : : "memory");
    return next;
}
#endif
// the sphere tests of one leaf reached through a 16-bit link
template <bool kCount>
VN_HD void leaf_test16(const node_f4* __restrict__ geom, uint32_t cur, f3 o, f3 d, float a, float inv_a, float& tbest, int& prim, TraceCounters& cnt) {
    const uint32_t first = (cur & 0x7FFFu) >> 3;
    const uint32_t count = (cur & 7u) + 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (uint32_t k = 0; k < count; k++) {
        const node_f4 g = geom[first + k];
        if (kCount) cnt.spheres += 1;
        const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
        if (t >= 0.0f) { tbest = t; prim = (int)(first + k); }
    }
}
// the sphere tests of one leaf (RayTracer.cu:229-270 on each of its <= 8 spheres)
template <bool kCount>
VN_HD void leaf_test(const node_f4* __restrict__ geom, uint32_t cur, f3 o, f3 d, float a, float inv_a, float& tbest, int& prim, TraceCounters& cnt, bool gate = false) {
    const uint32_t first = (cur & 0x7FFFFFFFu) >> 3;
    const uint32_t count = (cur & 7u) + 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (uint32_t k = 0; k < count; k++) {
        const node_f4 g = geom[first + k];
        if (kCount) cnt.spheres += 1;
        const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
        if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, g.x, g.y, g.z, g.w))) { tbest = t; prim = (int)(first + k); }
    }
}
// One leaf: the reference's ray/sphere test on its (<= 8) spheres; returns the next link.
template <bool kCount>
VN_HD uint32_t leaf_step(const node_f4* __restrict__ geom, uint32_t cur, f3 o, f3 d, float a, float inv_a, float& tbest, int& prim,
                         uint32_t* stack, int& sp, TraceCounters& cnt, bool gate = false) {
    const uint32_t first = (cur & 0x7FFFFFFFu) >> 3;
    const uint32_t count = (cur & 7u) + 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1          // leaves of SAH-built small scenes hold one sphere: keep the loop small instead of unrolled
#endif
    for (uint32_t k = 0; k < count; k++) {
        const node_f4 g = geom[first + k];
        if (kCount) cnt.spheres += 1;
        const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
        if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, g.x, g.y, g.z, g.w))) { tbest = t; prim = (int)(first + k); }
    }
    return sp ? stack[--sp] : kEmptyScene;
}
template <bool kCount>
VN_HD void closest_hit_wide(const node_f4* __restrict__ wnodes, uint32_t oct_stride, const node_f4* __restrict__ geom, uint32_t root_link,
                            f3 o, f3 d, float& t_out, int& prim_out, TraceCounters& cnt, float tbest0 = kTMax, int prim0 = -1) {
    float tbest = tbest0;              // closest hit among the huge spheres tested before the traversal (lbvh_core.cuh::HugeList)
    int prim = prim0;
#if defined(__CUDA_ARCH__)
    {
        const f3 idir = slab_idir(d);
        const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        const float a = dot(d, d);
        const float inv_a = rcp(a);
        uint32_t stack[kStackSize];
        const uint32_t base = (uint32_t)__cvta_generic_to_local(stack);
        uint32_t top = base;
        uint32_t cur = root_link;
        SlabScale ss = slab_scale(idir, ood, tbest);
        const WideBase wb = wide_base(wnodes, ray_octant(d), oct_stride);
        for (;;) {
            while (!(cur & kLeafFlag)) {
                if (kCount) cnt.nodes += 1;
                cur = wide_node_step_dev(wb, cur, ss, top, base);
            }
            if (cur == kEmptyScene) break;
            const float t_before = tbest;
            leaf_test<kCount>(geom, cur, o, d, a, inv_a, tbest, prim, cnt);
            cur = stack_pop_dev(top, base);
            if (tbest != t_before) ss = slab_scale(idir, ood, tbest);
        }
    }
#elif !defined(__CUDACC__)
    {
        const f3 idir = slab_idir(d);
        const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        const float a = dot(d, d);
        const float inv_a = rcp(a);
        uint32_t stack[kStackSize];
        int sp = 0;
        uint32_t cur = root_link;
        SlabScale ss = slab_scale(idir, ood, tbest);
        const WideBase wb = wide_base(wnodes, ray_octant(d), oct_stride);
        for (;;) {
            while (!(cur & kLeafFlag)) {
                if (kCount) cnt.nodes += 1;
                cur = wide_node_step_sat(wb, cur, ss, stack, sp);
            }
            if (cur == kEmptyScene) break;
            const float t_before = tbest;
            cur = leaf_step<kCount>(geom, cur, o, d, a, inv_a, tbest, prim, stack, sp, cnt);
            if (tbest != t_before) ss = slab_scale(idir, ood, tbest);
        }
    }
#endif
    t_out = tbest;
    prim_out = prim;
}

// ---- one full path, iterative: the recursion of RayTracer.cu:190-202 -> closest-hit -> optixTrace -> ... -> miss.
// The albedo product is accumulated forward (throughput), the reference multiplies on recursion unwind
// (RayTracer.cu:313,360): same factors, different association (<= depth * 2^-24 relative).
struct SceneView {
    const node_f4* nodes;
    const node_f4* geom;     // {cx, cy, cz, r}, sorted order
    const node_f4* mat;      // {albedo.xyz, fuzz} or, for glass, {ir, 1/ir, r0 front, r0 back} (DielectricConsts)
    const uint8_t* type;
    uint32_t root_link;
};

struct PathState {
    f3 o, d, thr;
    uint32_t seed;
    int depth;
};

// Closest hit over the CANONICAL 4-wide nodes in global memory (scenes too large for shared memory): child c of node i =
// wide[8i + 2c] = {lo.xyz, link}, wide[8i + 2c + 1] = {hi.xyz, count} -- one 128-byte line per step instead of two dependent
// 64-byte pair fetches.  No octant copies here (8 x 128 B x nodes would not stay in L2), so the step sorts the hit children
// by entry distance: keys = distance bits with the child index in the two low mantissa bits, 5 compare-exchanges.
VN_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
VN_HD uint32_t umax32(uint32_t a, uint32_t b) { return a < b ? b : a; }
VN_HD uint32_t wide_global_step(const node_f4* __restrict__ wide, uint32_t cur, f3 idir, f3 ood, float tbest, uint32_t* stack, int& sp) {
    const node_f4* __restrict__ p = wide + 8ull * cur;
    uint32_t key[4], link[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 4; c++) {
        const node_f4 lo = p[2 * c], hi = p[2 * c + 1];
        float tn;
        link[c] = f2u(lo.w);
        // an empty slot's inverted box only fails the near/far form of the octant copies; min/max would turn it inside out
        const bool h = box_hit(lo, hi, idir, ood, tbest, tn) && link[c] != kEmptyScene;
        key[c] = h ? ((f2u(tn) & ~3u) | (uint32_t)c) : 0xFFFFFFFFu;    // tn >= 0: float bits order like unsigned integers
    }
    // sorting network for 4 keys, ascending: (0,1)(2,3)(0,2)(1,3)(1,2)
    uint32_t a0 = umin32(key[0], key[1]), a1 = umax32(key[0], key[1]), a2 = umin32(key[2], key[3]), a3 = umax32(key[2], key[3]);
    const uint32_t b0 = umin32(a0, a2), b2 = umax32(a0, a2), b1 = umin32(a1, a3), b3 = umax32(a1, a3);
    const uint32_t c1 = umin32(b1, b2), c2 = umax32(b1, b2);
    // push far to near, descend into the nearest
    const uint32_t order[3] = {b3, c2, c1};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; i++) {
        const uint32_t k = order[i];
        const uint32_t idx = k & 3u;
        const uint32_t l = idx == 0u ? link[0] : (idx == 1u ? link[1] : (idx == 2u ? link[2] : link[3]));
        if (k != 0xFFFFFFFFu) stack[sp++] = l;
    }
    if (b0 == 0xFFFFFFFFu) return sp ? stack[--sp] : kEmptyScene;
    const uint32_t idx = b0 & 3u;
    return idx == 0u ? link[0] : (idx == 1u ? link[1] : (idx == 2u ? link[2] : link[3]));
}
template <bool kCount>
VN_HD void closest_hit_wide_global(const node_f4* __restrict__ wide, const node_f4* __restrict__ geom, uint32_t root_link,
                                   f3 o, f3 d, float& t_out, int& prim_out, TraceCounters& cnt, float tbest0 = kTMax, int prim0 = -1, bool gate = false) {
    float tbest = tbest0;
    int prim = prim0;
    {
        const f3 idir = slab_idir(d);
        const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
        const float a = dot(d, d);
        const float inv_a = rcp(a);
        uint32_t stack[kWideStackSize];
        int sp = 0;
        uint32_t cur = root_link;
        for (;;) {
            while (!(cur & kLeafFlag)) {
                if (kCount) cnt.nodes += 1;
                cur = wide_global_step(wide, cur, idir, ood, tbest, stack, sp);
            }
            if (cur == kEmptyScene) break;
            cur = leaf_step<kCount>(geom, cur, o, d, a, inv_a, tbest, prim, stack, sp, cnt, gate);
        }
    }
    t_out = tbest;
    prim_out = prim;
}

// ---- shading programs, one per class of hit, on the hit POINT p (= o + d*t, RayTracer.cu:256).  shade_segment() below
// composes them.
VN_HD f3 hit_point(f3 o, f3 d, float t) { return o + d * t; }
// normal + face-forwarding of hit_frame() for a known hit point (RayTracer.cu:257-258, 219-224)
VN_HD void hit_normal(f3 p, f3 d, const node_f4& g, f3& n, bool& front) {
    float inv = rcp(g.w);
    f3 normal = (p - mk3(g.x, g.y, g.z)) * inv;
    front = dot(d, normal) < 0.0f;
    n = front ? normal : -normal;
}
// __miss__ms (RayTracer.cu:442-450): radiance of a path that left the scene
VN_HD f3 shade_miss(f3 thr, f3 unit_direction) { return thr * sky(unit_direction); }
// __closesthit__lambertian / __closesthit__metal (RayTracer.cu:272-366) after the depth check.  `unit_direction` is only
// read for metal.  Returns false when the ray is absorbed.
VN_HD bool shade_opaque(uint32_t type, const node_f4& m, f3 unit_direction, f3 n, PathState& st) {
    const f3 s = random_in_unit_sphere(st.seed);          // both programs start with the same rejection loop
    f3 dir;
    bool ok = true;
    if (type == 0u) dir = scatter_lambertian(n, s);
    else ok = scatter_metal(unit_direction, n, m.w, s, dir);
    if (!ok) return false;
    st.d = dir;
    st.thr = st.thr * mk3(m.x, m.y, m.z);
    return true;
}

// Shades the closest hit (or miss) of one segment.  Returns true when the path continues (st updated), false when
// it ended with radiance `result`.
VN_HD bool shade_segment(const SceneView& sc, PathState& st, float t, int prim, f3& result) {
    // normalize(direction) is needed by miss (RayTracer.cu:444), metal (:338) and dielectric (:404); computing it for
    // every lane keeps the warp converged (a Lambertian lane simply does not use it).
    const f3 unit_direction = normalize(st.d);
    if (prim < 0) { result = shade_miss(st.thr, unit_direction); return false; }
    if (!(st.depth > 0)) { result = mk3(0.0f); return false; }      // RayTracer.cu:275,324,384: depth budget exhausted
    const node_f4 g = sc.geom[prim];
    const node_f4 m = sc.mat[prim];
    const uint32_t type = sc.type[prim];
    const f3 p = hit_point(st.o, st.d, t);
    f3 n;
    bool front;
    hit_normal(p, st.d, g, n, front);
    if (type != 2u) {
        if (!shade_opaque(type, m, unit_direction, n, st)) { result = mk3(0.0f); return false; }
    } else {
        DielectricConsts dc;
        dc.ir = m.x; dc.inv_ir = m.y; dc.r0_front = m.z; dc.r0_back = m.w;    // written by the builder (lbvh.cu::k_gather)
        st.d = scatter_dielectric(unit_direction, n, front, dc, st.seed);
    }
    st.o = p;
    st.depth -= 1;
    return true;
}

}  // namespace vn
