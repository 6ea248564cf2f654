/*
 * oracle.cpp -- CPU restatement of the reference's per-pixel Monte-Carlo loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared.
 * -ffp-contract=off matters: every float operation below is a single IEEE-754 binary32 operation in
 * the order the reference's source writes it (no FMA contraction), so the result is compiler-independent.
 *
 * Citations are file:line into /root/reference/Core/.
 */
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------ float3 (vec_math.h)
struct V3 { float x, y, z; };
inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 mk(float s) { return V3{s, s, s}; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }        // vec_math.h:410
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }        // vec_math.h:453
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }           // vec_math.h:457
inline V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }           // vec_math.h:461
inline V3 operator/(V3 a, float s) { float inv = 1.0f / s; return a * inv; }            // vec_math.h:483-487 (reciprocal-multiply)
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }              // vec_math.h:527-530
inline V3 normalize(V3 v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; } // vec_math.h:545-549
inline V3 lerp(V3 a, V3 b, float t) { return a + t * (b - a); }                         // vec_math.h:500-503
inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }        // vec_math.h:119-122
inline V3 clamp(V3 v, float a, float b) { return mk(clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b)); }
inline V3 reflect(V3 i, V3 n) { return i - 2.0f * n * dot(n, i); }                      // vec_math.h:558-561
inline V3 refract(V3 uv, V3 n, float etai_over_etat) {                                  // vec_math.h:564-570
    float cos_theta = fminf(dot(-uv, n), 1.0f);
    V3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    V3 r_out_parallel = -sqrtf(fabsf(1.0f - dot(r_out_perp, r_out_perp))) * n;
    return r_out_perp + r_out_parallel;
}

// ------------------------------------------------------------------ RNG (random.cuh)
inline uint32_t tea(uint32_t N, uint32_t val0, uint32_t val1) {                         // random.cuh:31-46
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < N; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
inline uint32_t lcg(uint32_t& prev) {                                                   // random.cuh:49-55
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00FFFFFFu;
}
inline float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }       // random.cuh:64-67

// RayTracer.cu:87-97
inline float random_float(uint32_t& seed) { return rnd(seed); }
inline float random_float(uint32_t& seed, float mn, float mx) { return mn + (mx - mn) * random_float(seed); }

struct Ctx {
    const orc_scene* scene;
    const orc_params* P;
    bool zyx;      // draw order of the three components of a vector (contract is x,y,z)
    bool forward;  // attenuation multiplication order
    bool bvh;
    bool gate = false;   // ORC_CLOSEST_GATE: the hit-point gate (see hit_gate)
    orc_stats st{};
};

// RayTracer.cu:108-115 -- the three draws of one vector; the reference leaves their order to the compiler.
inline V3 random_vec(uint32_t& seed, float mn, float mx, bool zyx) {
    V3 p;
    if (!zyx) { p.x = random_float(seed, mn, mx); p.y = random_float(seed, mn, mx); p.z = random_float(seed, mn, mx); }
    else      { p.z = random_float(seed, mn, mx); p.y = random_float(seed, mn, mx); p.x = random_float(seed, mn, mx); }
    return p;
}
inline V3 random_in_unit_sphere(uint32_t& seed, bool zyx) {                             // RayTracer.cu:117-125
    while (true) {
        V3 p = random_vec(seed, -1, 1, zyx);
        if (dot(p, p) >= 1) continue;
        return p;
    }
}
inline V3 random_unit_vector(uint32_t& seed, bool zyx) { return normalize(random_in_unit_sphere(seed, zyx)); } // :127-130
inline V3 random_in_unit_disk(uint32_t& seed, bool zyx) {                               // RayTracer.cu:141-149
    while (true) {
        V3 p;
        if (!zyx) { p.x = random_float(seed, -1, 1); p.y = random_float(seed, -1, 1); }
        else      { p.y = random_float(seed, -1, 1); p.x = random_float(seed, -1, 1); }
        p.z = 0;
        if (dot(p, p) >= 1) continue;
        return p;
    }
}

inline bool near_zero(V3 e) {                                                           // RayTracer.cu:8-13 (double compare)
    const double s = 1e-8;
    return ((double)fabsf(e.x) < s) && ((double)fabsf(e.y) < s) && ((double)fabsf(e.z) < s);
}

// ------------------------------------------------------------------ tonemap (RayTracer.cu:16-47)
inline V3 toSRGB(V3 c) {
    float invGamma = 1.0f / 2.4f;
    V3 powed = mk(powf(c.x, invGamma), powf(c.y, invGamma), powf(c.z, invGamma));
    return mk(c.x < 0.0031308f ? 12.92f * c.x : 1.055f * powed.x - 0.055f,
              c.y < 0.0031308f ? 12.92f * c.y : 1.055f * powed.y - 0.055f,
              c.z < 0.0031308f ? 12.92f * c.z : 1.055f * powed.z - 0.055f);
}
inline uint8_t quantizeUnsigned8Bits(float x) {
    x = clampf(x, 0.0f, 1.0f);
    unsigned v = (unsigned)(x * 256.0f);
    return (uint8_t)std::min(v, 255u);
}
inline void make_color(V3 c, uint8_t* out) {
    V3 srgb = toSRGB(clamp(c, 0.0f, 1.0f));
    out[0] = quantizeUnsigned8Bits(srgb.x);
    out[1] = quantizeUnsigned8Bits(srgb.y);
    out[2] = quantizeUnsigned8Bits(srgb.z);
    out[3] = 255u;
}

inline float reflectance(float cosine, float ref_idx) {                                 // RayTracer.cu:373-379
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * powf((1 - cosine), 5);   // reference: __powf (approximate on device)
}

}  // namespace

// ------------------------------------------------------------------ scene + CPU BVH
struct BNode {
    float lo[3], hi[3];
    int32_t left, right;   // internal: children; leaf: left = -1
    int32_t first, count;  // leaf range into order[]
};

struct orc_scene {
    std::vector<orc_sphere> s;
    std::vector<BNode> nodes;
    std::vector<int32_t> order;
};

namespace {

struct Hit { float t; int32_t prim; V3 p, n; bool front; };

// RayTracer.cu:229-270 (+ set_face_normal :219-224).  Returns true when the sphere reports a hit in [tmin,tmax].
// The hit-point gate (ORC_CLOSEST_GATE).  In the reference the intersection program only runs for rays that
// reach the sphere's AABB (custom primitives: Renderer.h:187-200 feeds OptiX the boxes of sphere.h:17-28), and a reported hit lies on
// the sphere.  Far from the origin the float quadratic below is dominated by rounding noise (its discriminant carries an absolute
// error of ~1e-7 |o - c|^2) and reports hits for rays that miss the sphere by a good fraction of its radius; whether OptiX would run
// the program for such a ray is decided by its closed-source box test.  The gate makes the outcome a function of (ray, sphere) alone
// -- independent of any BVH, so brute force, the oracle's BVH and the product's LBVH agree bit for bit: a root only counts when the hit
// point the reference computes (RayTracer.cu:256) lies inside the sphere's box grown by 0.5 % of the radius plus 2^-19 of the
// coordinates' magnitude.  True hits (|p - c| = r up to rounding) always pass; on the RTIOW scenes the gate never fires (tests).
inline bool hit_gate(const orc_sphere& s, V3 p) {
    const float g = fabsf(s.r) * 1.005f + (fabsf(s.cx) + fabsf(s.cy) + fabsf(s.cz) + fabsf(s.r)) * 1.9073486328125e-06f;
    return fabsf(p.x - s.cx) <= g && fabsf(p.y - s.cy) <= g && fabsf(p.z - s.cz) <= g;
}

template <bool kGate = false>
inline bool hit_sphere(const orc_sphere& s, V3 origin, V3 direction, float t_min, float t_max, Hit& h) {
    V3 center = mk(s.cx, s.cy, s.cz);
    V3 oc = origin - center;
    float a = dot(direction, direction);
    float half_b = dot(oc, direction);
    float c = dot(oc, oc) - s.r * s.r;
    float discriminant = half_b * half_b - a * c;
    if (discriminant < 0) return false;
    float sqrtd = sqrtf(discriminant);
    float root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    h.t = root;
    h.p = origin + direction * root;
    if (kGate && !hit_gate(s, h.p)) return false;
    V3 normal = (h.p - center) / s.r;
    h.front = dot(direction, normal) < 0;
    h.n = h.front ? normal : -normal;
    return true;
}

void build_bvh(orc_scene& sc) {
    const size_t n = sc.s.size();
    sc.order.resize(n);
    for (size_t i = 0; i < n; i++) sc.order[i] = (int32_t)i;
    sc.nodes.clear();
    if (n == 0) return;
    sc.nodes.reserve(2 * n);
    struct Job { int32_t node, first, count; };
    std::vector<Job> stack;
    sc.nodes.push_back(BNode{});
    stack.push_back({0, 0, (int32_t)n});
    while (!stack.empty()) {
        Job j = stack.back(); stack.pop_back();
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int32_t k = j.first; k < j.first + j.count; k++) {
            const orc_sphere& s = sc.s[sc.order[k]];
            const float c[3] = {s.cx, s.cy, s.cz};
            // conservative bound: |r| plus a relative + absolute pad so that float rounding in the
            // slab test can never cull a sphere the exact quadratic would accept.
            // (it also has to contain the hit-point gate's box with room for the slab test's own rounding: 2 % + 2^-16 of the coordinates)
            const float pad = fabsf(s.r) * 1.02f + 1e-4f + (fabsf(s.cx) + fabsf(s.cy) + fabsf(s.cz) + fabsf(s.r)) * 1.52587890625e-05f;
            for (int a = 0; a < 3; a++) {
                lo[a] = std::min(lo[a], c[a] - pad); hi[a] = std::max(hi[a], c[a] + pad);
                clo[a] = std::min(clo[a], c[a]);     chi[a] = std::max(chi[a], c[a]);
            }
        }
        BNode nd{};
        for (int a = 0; a < 3; a++) { nd.lo[a] = lo[a]; nd.hi[a] = hi[a]; }
        int axis = 0;
        if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
        if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
        if (j.count <= 2 || !(chi[axis] > clo[axis])) {
            nd.left = -1; nd.right = -1; nd.first = j.first; nd.count = j.count;
            sc.nodes[j.node] = nd;
            continue;
        }
        int32_t mid = j.first + j.count / 2;
        auto key = [&](int32_t idx) { const orc_sphere& s = sc.s[idx]; return axis == 0 ? s.cx : axis == 1 ? s.cy : s.cz; };
        std::nth_element(sc.order.begin() + j.first, sc.order.begin() + mid, sc.order.begin() + j.first + j.count,
                         [&](int32_t a, int32_t b) { return key(a) < key(b); });
        nd.left = (int32_t)sc.nodes.size();  sc.nodes.push_back(BNode{});
        nd.right = (int32_t)sc.nodes.size(); sc.nodes.push_back(BNode{});
        nd.first = 0; nd.count = 0;
        sc.nodes[j.node] = nd;
        stack.push_back({nd.left, j.first, mid - j.first});
        stack.push_back({nd.right, mid, j.first + j.count - mid});
    }
}

// Conservative slab test against a padded box; returns the entry distance in tn.
inline bool slab(const BNode& b, V3 o, V3 inv, float tbest, float& tn_out) {
    float t0x = (b.lo[0] - o.x) * inv.x, t1x = (b.hi[0] - o.x) * inv.x;
    float t0y = (b.lo[1] - o.y) * inv.y, t1y = (b.hi[1] - o.y) * inv.y;
    float t0z = (b.lo[2] - o.z) * inv.z, t1z = (b.hi[2] - o.z) * inv.z;
    // plain compare-selects (minss/maxss), not libm fminf/fmaxf calls: 10x faster.  A NaN (0*inf: origin exactly on a
    // padded face plane, direction parallel to it) then culls the box, which is right: such a ray cannot touch a
    // sphere that lies strictly inside the padded box.
    auto mn = [](float a, float b) { return a < b ? a : b; };
    auto mx = [](float a, float b) { return a > b ? a : b; };
    float tn = mx(mx(mn(t0x, t1x), mn(t0y, t1y)), mn(t0z, t1z));
    float tf = mn(mn(mx(t0x, t1x), mx(t0y, t1y)), mx(t0z, t1z));
    tn_out = tn;
    return !(tn > tf * 1.00001f) && !(tf < 0.0f) && !(tn > tbest);
}

// Closest hit = optixTrace(tmin=0.001f, tmax=1e16f) (RayTracer.cu:190-202): OptiX shrinks tmax to the closest
// accepted intersection, so every sphere is tested against the current best t.
template <bool kGate = false>
inline bool closest_brute(const orc_scene& sc, V3 o, V3 d, Hit& best, orc_stats& st) {
    float tmax = 1e16f;
    bool any = false;
    Hit h;
    for (size_t i = 0; i < sc.s.size(); i++) {
        if (hit_sphere<kGate>(sc.s[i], o, d, 0.001f, tmax, h)) { h.prim = (int32_t)i; best = h; tmax = h.t; any = true; }
    }
    st.sphere_tests += sc.s.size();
    return any;
}

// The oracle's own BVH traversal (near child first).  Same closest-hit semantics as brute force; only used because
// brute force is too slow for full frames, and validated against it in tests/test_oracle.py.
template <bool kGate = false>
inline bool closest_bvh(const orc_scene& sc, V3 o, V3 d, Hit& best, orc_stats& st) {
    if (sc.nodes.empty()) return false;
    float tmax = 1e16f;
    bool any = false;
    V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int32_t stack[128];
    int sp = 0;
    float tn;
    if (!slab(sc.nodes[0], o, inv, tmax, tn)) return false;
    stack[sp++] = 0;
    Hit h;
    while (sp) {
        const BNode& nd = sc.nodes[stack[--sp]];
        st.node_visits++;
        if (nd.left < 0) {
            for (int32_t k = nd.first; k < nd.first + nd.count; k++) {
                int32_t idx = sc.order[k];
                st.sphere_tests++;
                if (hit_sphere<kGate>(sc.s[idx], o, d, 0.001f, tmax, h)) { h.prim = idx; best = h; tmax = h.t; any = true; }
            }
        } else {
            float tl, tr;
            const bool hl = slab(sc.nodes[nd.left], o, inv, tmax, tl);
            const bool hr = slab(sc.nodes[nd.right], o, inv, tmax, tr);
            if (hl && hr) {
                if (tl <= tr) { stack[sp++] = nd.right; stack[sp++] = nd.left; }
                else          { stack[sp++] = nd.left;  stack[sp++] = nd.right; }
            } else if (hl) stack[sp++] = nd.left;
            else if (hr) stack[sp++] = nd.right;
        }
    }
    return any;
}

inline bool closest(Ctx& c, V3 o, V3 d, Hit& h) {
    c.st.segments++;
    if (c.gate) return c.bvh ? closest_bvh<true>(*c.scene, o, d, h, c.st) : closest_brute<true>(*c.scene, o, d, h, c.st);
    return c.bvh ? closest_bvh(*c.scene, o, d, h, c.st) : closest_brute(*c.scene, o, d, h, c.st);
}

// One path = the recursion optixTrace -> closest-hit -> optixTrace ... -> miss, restated iteratively.
// The reference multiplies the albedos while the recursion unwinds (RayTracer.cu:313,360):
//   atten = ((sky * a_k) * a_{k-1}) ... * a_1.   ORC_ATTEN_FORWARD gives (((1*a_1)*a_2)...*a_k)*sky instead,
// which is what an iterative tracer computes; the two differ by float re-association only.
V3 trace_path(Ctx& c, V3 origin, V3 direction, uint32_t& seed) {
    const std::vector<orc_sphere>& S = c.scene->s;
    int depth = (int)c.P->max_depth - 1;                       // RayTracer.cu:172,184 (max_depth is 4 there)
    V3 chain_stack[64];
    std::vector<V3> chain_heap;
    V3* chain = chain_stack;
    if (c.P->max_depth > 64) { chain_heap.resize(c.P->max_depth); chain = chain_heap.data(); }
    int nchain = 0;
    V3 throughput = mk(1.0f);
    V3 atten;
    bool absorbed = false;
    uint64_t nseg = 0;
    while (true) {
        Hit h;
        nseg++;
        if (!closest(c, origin, direction, h)) {
            // __miss__ms, RayTracer.cu:442-450
            V3 unit_direction = normalize(direction);
            double t = 0.5 * ((double)unit_direction.y + 1.0);
            atten = lerp(mk(1.0f), mk((float)0.5, (float)0.7, (float)1.0), (float)t);
            depth -= 1;
            break;
        }
        const orc_sphere& s = S[h.prim];
        if (!(depth > 0)) { atten = mk(0.0f); absorbed = true; break; }   // :275,315-318 / :324,367-370 / :384,436-439
        if (s.type == 0) {
            // __closesthit__lambertian, RayTracer.cu:272-319
            V3 scatter_direction = h.n + random_unit_vector(seed, c.zyx);
            if (near_zero(scatter_direction)) scatter_direction = h.n;
            origin = h.p;
            direction = scatter_direction;
            depth -= 1;
            chain[nchain++] = mk(s.ax, s.ay, s.az);
            throughput = throughput * mk(s.ax, s.ay, s.az);
        } else if (s.type == 1) {
            // __closesthit__metal, RayTracer.cu:321-371
            V3 reflected = reflect(normalize(direction), h.n);
            origin = h.p;
            direction = reflected + s.fuzz_or_ir * random_in_unit_sphere(seed, c.zyx);
            if (dot(direction, h.n) > 0) {
                depth -= 1;
                chain[nchain++] = mk(s.ax, s.ay, s.az);
                throughput = throughput * mk(s.ax, s.ay, s.az);
            } else {
                atten = mk(0.0f); absorbed = true;
                break;
            }
        } else {
            // __closesthit__dielectric, RayTracer.cu:381-440
            float refraction_ratio = s.fuzz_or_ir;
            if (h.front) refraction_ratio = (1.0f / s.fuzz_or_ir);
            V3 unit_direction = normalize(direction);
            double cos_theta = fminf(dot(-unit_direction, h.n), 1.0);
            double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
            bool cannot_refract = refraction_ratio * sin_theta > 1.0;
            V3 dir;
            if (cannot_refract || reflectance((float)cos_theta, refraction_ratio) > random_float(seed))
                dir = reflect(unit_direction, h.n);
            else
                dir = refract(unit_direction, h.n, refraction_ratio);
            origin = h.p;
            direction = dir;
            depth -= 1;
        }
    }
    c.st.paths++;
    c.st.max_segments_in_path = std::max(c.st.max_segments_in_path, nseg);
    if (c.forward) return absorbed ? mk(0.0f) : throughput * atten;
    for (int k = nchain - 1; k >= 0; k--) atten = atten * chain[k];   // prd->attenuation *= albedo on unwind
    return atten;
}

// get_ray, RayTracer.cu:151-161
inline void get_ray(Ctx& c, float s, float t, V3& origin, V3& direction, uint32_t& seed) {
    const orc_params& P = *c.P;
    V3 pu = mk(P.u[0], P.u[1], P.u[2]), pv = mk(P.v[0], P.v[1], P.v[2]), pw = mk(P.w[0], P.w[1], P.w[2]);
    V3 rd = P.lens_radius * random_in_unit_disk(seed, c.zyx);
    V3 u = normalize(pu);
    V3 v = normalize(pv);
    V3 offset = u * rd.x + v * rd.y;
    origin = mk(P.origin[0], P.origin[1], P.origin[2]) + offset;
    direction = pw + s * pu * 0.5f + t * pv * 0.5f - offset;
}

// __raygen__rg up to the mean, RayTracer.cu:163-206
void render_pixel(Ctx& c, uint32_t image_index, float* mean_rgba, float* per_sample) {
    const orc_params& P = *c.P;
    const uint32_t x = image_index % P.width, y = image_index / P.width;
    V3 pixel_color = mk(0.0f);
    uint32_t seed = tea(4, image_index, P.subframe_index);
    for (uint32_t sidx = 0; sidx < P.samples_per_pixel; ++sidx) {
        float u = 2 * float(x + random_float(seed)) / (P.width - 1) - 1;
        float v = 2 * float(y + random_float(seed)) / (P.height - 1) - 1;
        V3 origin, direction;
        get_ray(c, u, v, origin, direction, seed);
        uint32_t prd_seed = seed;                               // prd.seed = seed is a COPY (:183); raygen's seed is not updated
        V3 a = trace_path(c, origin, direction, prd_seed);
        if (per_sample) { per_sample[3 * sidx + 0] = a.x; per_sample[3 * sidx + 1] = a.y; per_sample[3 * sidx + 2] = a.z; }
        pixel_color = pixel_color + a;
    }
    V3 accum_color = pixel_color / static_cast<float>(P.samples_per_pixel);
    mean_rgba[0] = accum_color.x; mean_rgba[1] = accum_color.y; mean_rgba[2] = accum_color.z; mean_rgba[3] = 1.0f;
}

}  // namespace

// ------------------------------------------------------------------ C interface
extern "C" {

uint32_t orc_tea(uint32_t rounds, uint32_t v0, uint32_t v1) { return tea(rounds, v0, v1); }
uint32_t orc_lcg(uint32_t* state) { return lcg(*state); }
float orc_rnd(uint32_t* state) { return rnd(*state); }

void orc_normalize(const float* v, float* out) { V3 r = normalize(mk(v[0], v[1], v[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_reflect(const float* i, const float* n, float* out) { V3 r = reflect(mk(i[0], i[1], i[2]), mk(n[0], n[1], n[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_refract(const float* uv, const float* n, float eta, float* out) { V3 r = refract(mk(uv[0], uv[1], uv[2]), mk(n[0], n[1], n[2]), eta); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_lerp(const float* a, const float* b, float t, float* out) { V3 r = lerp(mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]), t); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_make_color(const float* rgb, uint8_t* out4) { make_color(mk(rgb[0], rgb[1], rgb[2]), out4); }
float orc_reflectance(float cosine, float ref_idx) { return reflectance(cosine, ref_idx); }

// Scene.h:13-80.  The reference draws several random numbers inside single argument lists (Scene.h:25,33,42,
// 106-112), whose order C++ leaves unspecified, through std::uniform_real_distribution<float>, whose algorithm is
// implementation-defined.  Pinned here: left-to-right, and libstdc++'s generate_canonical<float,24> for a 32-bit
// engine, i.e. float(g()) / 2^32 with the result clamped below 1.  std::mt19937 itself is fully specified.
uint32_t orc_scene_rtiow_final(orc_sphere* out, uint32_t cap) {
    std::mt19937 generator;   // default seed 5489 (Scene.h:89)
    auto random_float = [&]() -> float {
        float r = (float)generator() / 4294967296.0f;
        if (r >= 1.0f) r = std::nextafter(1.0f, 0.0f);
        return r;
    };
    auto random_float_mm = [&](float mn, float mx) -> float { return mn + (mx - mn) * random_float(); };  // Scene.h:93-96
    std::vector<orc_sphere> v;
    auto push = [&](float cx, float cy, float cz, float r, uint32_t type, float ax, float ay, float az, float fuzz, float ir) {
        orc_sphere s{cx, cy, cz, r, ax, ay, az, type == 2 ? ir : fuzz, type};
        v.push_back(s);
    };
    push(0.0f, -1000.0f, 0.0f, 1000.0f, 0, 0.5f, 0.5f, 0.5f, 0, 0);                      // Scene.h:15-20
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            float choose_mat = random_float();                                          // Scene.h:24
            float cx = (float)(a + 0.9 * random_float());                               // Scene.h:25 (double arithmetic, then float)
            float cy = (float)0.2;
            float cz = (float)(b + 0.9 * random_float());
            float dx = cx - 4.0f, dy = cy - (float)0.2, dz = cz - 0.0f;                 // glm::length(center - vec3(4,0.2,0))
            float len = sqrtf(dx * dx + dy * dy + dz * dz);
            if (len > 0.9) {                                                            // Scene.h:27 (float vs double literal)
                if (choose_mat < 0.8) {                                                 // Scene.h:31-39
                    float r1x = random_float(), r1y = random_float(), r1z = random_float();
                    float r2x = random_float(), r2y = random_float(), r2z = random_float();
                    push(cx, cy, cz, 0.2f, 0, r1x * r2x, r1y * r2y, r1z * r2z, 0, 0);
                } else if (choose_mat < 0.95) {                                         // Scene.h:40-49
                    float ax = random_float_mm(0.5f, 1.0f), ay = random_float_mm(0.5f, 1.0f), az = random_float_mm(0.5f, 1.0f);
                    float fuzz = random_float_mm(0.0f, 0.5f);
                    push(cx, cy, cz, 0.2f, 1, ax, ay, az, fuzz, 0);
                } else {                                                                // Scene.h:50-57
                    push(cx, cy, cz, 0.2f, 2, 0, 0, 0, 0, 1.5f);
                }
            }
        }
    }
    push(0, 1, 0, 1.0f, 2, 0, 0, 0, 0, 1.5f);                                           // Scene.h:62-66
    push(-4, 1, 0, 1.0f, 0, (float)0.4, (float)0.2, (float)0.1, 0, 0);                  // Scene.h:68-72
    push(4, 1, 0, 1.0f, 1, (float)0.7, (float)0.6, (float)0.5, 0.0f, 0);                // Scene.h:74-78
    uint32_t n = (uint32_t)v.size();
    if (out) for (uint32_t i = 0; i < n && i < cap; i++) out[i] = v[i];
    return n;
}

void orc_scene_random(orc_sphere* out, uint64_t n, uint32_t seed, float S, uint32_t mix) {
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s = tea(4, (uint32_t)i, seed);
        orc_sphere o{};
        float rx = rnd(s); o.cx = S * (2.0f * rx - 1.0f);
        float ry = rnd(s); o.cy = S * (2.0f * ry - 1.0f);
        float rz = rnd(s); o.cz = S * (2.0f * rz - 1.0f);
        float rr = rnd(s); o.r = 0.1f + 0.2f * rr;
        float m = rnd(s);
        uint32_t type;
        if (mix == 0) type = m < 0.80f ? 0u : (m < 0.95f ? 1u : 2u);
        else          type = m < 0.50f ? 2u : (m < 0.90f ? 0u : 1u);
        o.type = type;
        if (type == 0) {
            float a0 = rnd(s), a1 = rnd(s); o.ax = a0 * a1;
            float b0 = rnd(s), b1 = rnd(s); o.ay = b0 * b1;
            float c0 = rnd(s), c1 = rnd(s); o.az = c0 * c1;
            o.fuzz_or_ir = 0.0f;
        } else if (type == 1) {
            float a0 = rnd(s); o.ax = 0.5f + 0.5f * a0;
            float a1 = rnd(s); o.ay = 0.5f + 0.5f * a1;
            float a2 = rnd(s); o.az = 0.5f + 0.5f * a2;
            float f = rnd(s);  o.fuzz_or_ir = 0.5f * f;
        } else {
            o.ax = o.ay = o.az = 0.0f;
            o.fuzz_or_ir = 1.5f;
        }
        out[i] = o;
    }
}

// Camera::SetForward (camera.h:24) + Camera::UpdateUVW (Camera.cpp:24-37) with glm's float formulas written out:
// glm::normalize(v) = v * (1/sqrt(dot(v,v))), glm::cross, glm::radians(d) = d * 0.017453292519943295f.
void orc_camera(const float* lookfrom, const float* fwd, float vfov_deg, float aspect, float aperture, float focal,
                float* origin, float* u, float* v, float* w, float* lens_radius) {
    auto gnorm = [](V3 a) { float inv = 1.0f / sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); return mk(a.x * inv, a.y * inv, a.z * inv); };
    auto gcross = [](V3 a, V3 b) { return mk(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); };
    V3 forward = gnorm(mk(fwd[0], fwd[1], fwd[2]));
    V3 up = mk(0.0f, 1.0f, 0.0f);
    V3 m_w = forward * focal;
    V3 m_u = gnorm(gcross(m_w, up));
    V3 m_v = gnorm(gcross(m_u, m_w));
    float theta = vfov_deg * 0.01745329251994329576923690768489f;
    float h = tanf(theta * 0.5f);
    float viewportHeight = 2.0f * h;
    float viewportWidth = aspect * viewportHeight;
    m_u = m_u * (focal * viewportWidth);
    m_v = m_v * (focal * viewportHeight);
    origin[0] = lookfrom[0]; origin[1] = lookfrom[1]; origin[2] = lookfrom[2];
    u[0] = m_u.x; u[1] = m_u.y; u[2] = m_u.z;
    v[0] = m_v.x; v[1] = m_v.y; v[2] = m_v.z;
    w[0] = m_w.x; w[1] = m_w.y; w[2] = m_w.z;
    *lens_radius = aperture * 0.5f;                                                     // camera.h:20
}

orc_scene* orc_scene_create(const orc_sphere* spheres, uint64_t n) {
    orc_scene* sc = new orc_scene();
    sc->s.assign(spheres, spheres + n);
    build_bvh(*sc);
    return sc;
}
void orc_scene_destroy(orc_scene* sc) { delete sc; }

void orc_render_mean(const orc_scene* sc, const orc_params* P, const uint32_t* pixels, uint64_t n_pixels,
                     float* mean_rgba, float* per_sample, orc_stats* stats) {
    const uint64_t total = pixels ? n_pixels : (uint64_t)P->width * P->height;
    unsigned nthreads = P->threads ? P->threads : std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> next{0};
    const uint64_t chunk = 64;
    std::vector<orc_stats> per_thread(nthreads);
    auto work = [&](unsigned tid) {
        Ctx c;
        c.scene = sc; c.P = P;
        c.zyx = P->draw_order == ORC_DRAW_ZYX;
        c.forward = P->atten_order == ORC_ATTEN_FORWARD;
        c.bvh = (P->closest & ORC_CLOSEST_BVH) != 0;
        c.gate = (P->closest & ORC_CLOSEST_GATE) != 0;
        while (true) {
            uint64_t b = next.fetch_add(chunk);
            if (b >= total) break;
            uint64_t e = std::min(total, b + chunk);
            for (uint64_t k = b; k < e; k++) {
                uint32_t px = pixels ? pixels[k] : (uint32_t)k;
                render_pixel(c, px, mean_rgba + 4ull * px,
                             per_sample ? per_sample + 3ull * P->samples_per_pixel * k : nullptr);
            }
        }
        per_thread[tid] = c.st;
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    if (stats) {
        orc_stats s{};
        for (auto& p : per_thread) {
            s.segments += p.segments; s.paths += p.paths; s.node_visits += p.node_visits; s.sphere_tests += p.sphere_tests;
            s.max_segments_in_path = std::max(s.max_segments_in_path, p.max_segments_in_path);
        }
        *stats = s;
    }
}

void orc_accumulate_tonemap(const float* prev_rgba, const float* mean_rgba, int blend, float a,
                            float* out_rgba, uint8_t* out_image, uint64_t n_pixels) {
    for (uint64_t i = 0; i < n_pixels; i++) {
        V3 accum_color = mk(mean_rgba[4 * i], mean_rgba[4 * i + 1], mean_rgba[4 * i + 2]);
        if (blend) {                                                                    // RayTracer.cu:208-213
            V3 prev = mk(prev_rgba[4 * i], prev_rgba[4 * i + 1], prev_rgba[4 * i + 2]);
            accum_color = lerp(prev, accum_color, a);
        }
        if (out_rgba) { out_rgba[4 * i] = accum_color.x; out_rgba[4 * i + 1] = accum_color.y; out_rgba[4 * i + 2] = accum_color.z; out_rgba[4 * i + 3] = 1.0f; }
        if (out_image) make_color(accum_color, out_image + 4 * i);                      // RayTracer.cu:216
    }
}

void orc_closest_hit(const orc_scene* sc, int use_bvh, const float* origins, const float* dirs, uint64_t n,
                     float* t_out, int32_t* prim_out) {
    orc_stats st{};
    for (uint64_t i = 0; i < n; i++) {
        V3 o = mk(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        V3 d = mk(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        Hit h;
        const bool bvh = (use_bvh & ORC_CLOSEST_BVH) != 0, gate = (use_bvh & ORC_CLOSEST_GATE) != 0;
        bool any = gate ? (bvh ? closest_bvh<true>(*sc, o, d, h, st) : closest_brute<true>(*sc, o, d, h, st))
                        : (bvh ? closest_bvh(*sc, o, d, h, st) : closest_brute(*sc, o, d, h, st));
        t_out[i] = any ? h.t : -1.0f;
        prim_out[i] = any ? h.prim : -1;
    }
}

}  // extern "C"
