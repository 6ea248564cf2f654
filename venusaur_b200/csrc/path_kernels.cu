// path_kernels.cu -- the persistent-thread path kernel, the accumulate/tonemap kernels and the unit-test kernels.
// Compiled twice (see kernels.h): -DVN_EXACT=1 -fmad=false -> namespace vn::exact, -DVN_EXACT=0 -> vn::fast.
//
// k_render_persistent replaces the whole OptiX launch of Renderer::Draw (Renderer.h:75): __raygen__rg,
// optixTrace + __intersection__hit_sphere, the three closest-hit programs and __miss__ms (RayTracer.cu:163-450).
//
// Design (B200: 148 SMs, 227 KB smem, no RT cores):
//   * one resident CTA wave (grid = SMs x occupancy); every lane owns one pixel at a time and walks its spp samples
//     in order (so the per-pixel RNG chain of RayTracer.cu:169-183 and the float summation order are the
//     reference's); when its pixel is done the lane takes the next pixel from a global ticket (warp-aggregated
//     atomic), so lanes never idle on finished paths (path regeneration).
//   * the whole scene (32-byte BVH nodes, sphere geometry, materials) is staged once per CTA into shared memory when
//     it fits (RTIOW: ~35 KB), so traversal never leaves the SM: LDS.128 node fetches, no L2/HBM traffic at all;
//     larger scenes are traversed straight from L2/HBM through the same code (128-bit __ldg loads).
//   * the accumulate step (RayTracer.cu:206-215) is fused at the end of each pixel; sRGB + quantise (:216) is the
//     coalesced k_tonemap launched right behind (one pixel per thread, 128-byte uchar4 stores per warp).
//   * k_render_async (the default for scenes whose 4-wide nodes fit in shared memory) is the same per-lane machinery with the
//     round structure relaxed: traversal state survives the shading code, a burst ends when enough lanes are finished, warps
//     own whole 8x4 tiles, and vn_api.cu hands the tiles out most expensive first.
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include <algorithm>
#include <type_traits>

#include "kernels.h"
#include "lbvh_core.cuh"

namespace cg = cooperative_groups;

#if VN_EXACT
#define VN_NS exact
#else
#define VN_NS fast
#endif

namespace vn {

#if VN_EXACT
size_t scene_smem_bytes(uint32_t num_nodes, uint32_t num_spheres, uint32_t node_copies) {
    return (size_t)num_nodes * 32 * node_copies + (size_t)num_spheres * 32 + (((size_t)num_spheres + 15) & ~(size_t)15);
}
size_t grid_smem_bytes(uint32_t n_cells, uint32_t n_refs, uint32_t num_spheres) {
    const size_t idx = (((size_t)n_cells + 1 + n_refs) * 2 + 15) & ~(size_t)15;
    return idx + (size_t)num_spheres * 32 + (((size_t)num_spheres + 15) & ~(size_t)15);
}
size_t wide_smem_bytes(uint32_t num_wide, uint32_t num_spheres) {
    return (size_t)num_wide * 16 * kWideNodeF4 * 8 + (size_t)num_spheres * 32 + (((size_t)num_spheres + 15) & ~(size_t)15);
}
#endif

namespace VN_NS {

namespace {

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t fetch_work(uint32_t* counter) {
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, g.size());
    base = g.shfl(base, 0);
    return base + g.thread_rank();
}

__device__ __forceinline__ void finish_pixel(const RenderLaunch& p, uint32_t pix, f3 sum) {
    f3 mean = sum * p.inv_spp;                                   // RayTracer.cu:206 (vec_math.h:483-487)
    if (p.blend_mode == kBlendLerp) {                            // RayTracer.cu:208-213
        const float4 prev = p.accum[pix];
        mean = lerp3(mk3(prev.x, prev.y, prev.z), mean, p.blend_a);
    } else if (p.blend_mode == kBlendSum) {
        const float4 prev = p.accum[pix];
        mean = mk3(prev.x, prev.y, prev.z) + mean;
    }
    p.accum[pix] = make_float4(mean.x, mean.y, mean.z, 1.0f);    // RayTracer.cu:215
    // RayTracer.cu:216 (make_color) runs afterwards in k_tonemap: three powf per pixel executed by the one or two lanes
    // that happen to finish a pixel in this iteration would cost every other lane of the warp the same issue slots.
}

// Several subframes in ONE launch (k_render_lean<kMulti>, vn_render_subframes): the tickets are (subframe, tile), subframe-major, so the launch
// drains once instead of once per subframe.  The subframes of a pixel must still be blended in order (RayTracer.cu:208-213 is a running
// mean): the accumulation buffer's .w -- 1.0f by contract (RayTracer.cu:215) -- carries "subframe j of this launch is in" while the launch
// runs (kMultiTag + j; the last subframe writes the 1.0f back).  A pixel's float4 is read and written whole (one 16-byte L2 access each),
// so whoever holds subframe j of a pixel sees either the value it needs, tag included, or not yet: then it keeps the finished sum and
// asks again in its warp's next iteration.  No fences, no counters; a lane waits only where a pixel of the subframe before is still
// bouncing (the launch's tickets hand a tile's subframes out a whole frame apart).
// pxy of those launches: px in bits 0-12, py in bits 16-28, the subframe (0-63) in bits 13-15 and 29-31.
constexpr uint32_t kMultiTag = 0x40000000u;
__device__ __forceinline__ uint32_t multi_px(uint32_t pxy) { return pxy & 0x1FFFu; }
__device__ __forceinline__ uint32_t multi_py(uint32_t pxy) { return (pxy >> 16) & 0x1FFFu; }
__device__ __forceinline__ uint32_t multi_sub(uint32_t pxy) { return ((pxy >> 13) & 7u) | ((pxy >> 26) & 0x38u); }
__device__ __forceinline__ bool finish_pixel_multi(const RenderLaunch& p, uint32_t pxy, f3 sum) {
    const uint32_t j = multi_sub(pxy);
    float4* a = p.accum + (multi_py(pxy) * p.width + multi_px(pxy));
    f3 mean = sum * p.inv_spp;                                   // RayTracer.cu:206 (vec_math.h:483-487)
    const uint32_t have = p.accum_count + j;                     // subframes the running mean holds when this one is blended
    if (p.blend_mode == kBlendSum || have > 0u) {
        const float4 prev = __ldcg(a);                           // (from L2: another SM wrote it, moments or a frame ago)
        if (j > 0u && __float_as_uint(prev.w) != kMultiTag + j - 1u) return false;
        if (p.blend_mode == kBlendSum) mean = mk3(prev.x, prev.y, prev.z) + mean;
        else mean = lerp3(mk3(prev.x, prev.y, prev.z), mean, __frcp_rn((float)(have + 1u)));     // RayTracer.cu:208-213; a = 1 / (subframes + 1), as fill_launch forms it
    }
    __stcg(a, make_float4(mean.x, mean.y, mean.z, j + 1u == p.n_sub ? 1.0f : __uint_as_float(kMultiTag + j)));   // RayTracer.cu:215
    return true;
}

// closest_hit_wide() with a warp vote instead of the while-while phases: every iteration the converged lanes of the warp
// either all take a node step or all test a leaf sphere set -- the leaf turn comes when `leaf_vote` lanes wait at a leaf (or
// nobody stands on a node).  With 4-wide nodes a ray only takes ~5 node steps between ~2 leaves, so in the while-while
// form a lane that reached its leaf idled until the slowest lane of the warp had found one too (9.5 of 32 lanes active in
// the node step, ncu); tools/simt_model + the SIMT simulator predicted 16 of 32 for a threshold of 8.  Same steps per
// ray, same order, same result.
template <bool kCount>
__device__ __forceinline__ void closest_hit_wide_vote(const float4* __restrict__ wnodes, uint32_t oct_stride, const float4* __restrict__ geom,
                                                      uint32_t root_link, uint32_t leaf_vote, f3 o, f3 d, float& t_out, int& prim_out,
                                                      TraceCounters& cnt, float tbest0, int prim0) {
    float tbest = tbest0;              // closest hit among the huge spheres (tested by the caller, all lanes together)
    int prim = prim0;
    const f3 idir = slab_idir(d);
    const float4* __restrict__ wn = wnodes + ray_octant(d) * oct_stride;
    const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const float a = dot(d, d);
    const float inv_a = rcp(a);
    uint32_t stack[kStackSize];
    int sp = 0;
    uint32_t cur = root_link;
    while (cur != kEmptyScene) {
        const bool at_leaf = (cur & kLeafFlag) != 0u;
        const unsigned act = __activemask();
        const unsigned lm = __ballot_sync(act, at_leaf);
        if (lm == act || (uint32_t)__popc(lm) >= leaf_vote) {
            if (at_leaf) cur = leaf_step<kCount>(geom, cur, o, d, a, inv_a, tbest, prim, stack, sp, cnt);
        } else if (!at_leaf) {
            if (kCount) cnt.nodes += 1;
            cur = wide_node_step(wn, cur, idir, ood, tbest, stack, sp);
        }
    }
    t_out = tbest;
    prim_out = prim;
}

// Closest hit through the uniform grid + oversize list (grid_core.cuh) with the same kind of vote: the oversize spheres are
// tested by all lanes together; then every iteration is either a sphere turn (lanes with untested references in their current
// cell test ONE sphere) or an advance turn (lanes whose cell is exhausted step the DDA to the next cell).
template <bool kCount>
__device__ __forceinline__ void closest_hit_grid_vote(const GridHeader& g, const uint16_t* __restrict__ start, const uint16_t* __restrict__ refs,
                                                      const float4* __restrict__ geom, uint32_t sphere_vote, f3 o, f3 d, float& t_out, int& prim_out,
                                                      TraceCounters& cnt, bool gate) {
    float tbest = kTMax;
    int prim = -1;
    const float a = dot(d, d);
    const float inv_a = rcp(a);
    for (uint32_t i = 0; i < g.n_big; i++) {
        const uint32_t s = g.big[i];
        const float4 sp = geom[s];
        if (kCount) cnt.spheres += 1;
        const float t = sphere_root(o, d, a, inv_a, sp.x, sp.y, sp.z, sp.w, kTMin, tbest);
        if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, sp.x, sp.y, sp.z, sp.w))) { tbest = t; prim = (int)s; }
    }
    GridRay r;
    grid_ray_setup(g, start, o, d, tbest, r);
    if (kCount && r.alive) cnt.nodes += 1;
    if (sphere_vote == 0u) {
        // while-while form: walk to the next cell that holds a sphere, then test one sphere
        while (r.alive) {
            while (r.alive && r.q >= r.q_end) {
                grid_ray_advance(g, start, tbest, r);
                if (kCount && r.alive) cnt.nodes += 1;
            }
            if (!r.alive) break;
            const uint32_t s = refs[r.q++];
            const float4 sp = geom[s];
            if (kCount) cnt.spheres += 1;
            const float t = sphere_root(o, d, a, inv_a, sp.x, sp.y, sp.z, sp.w, kTMin, tbest);
            if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, sp.x, sp.y, sp.z, sp.w))) { tbest = t; prim = (int)s; }
        }
        t_out = tbest;
        prim_out = prim;
        return;
    }
    while (r.alive) {
        const bool at_sphere = r.q < r.q_end;
        const unsigned act = __activemask();
        const unsigned sm = __ballot_sync(act, at_sphere);
        if (sm == act || (uint32_t)__popc(sm) >= sphere_vote) {
            if (at_sphere) {
                const uint32_t s = refs[r.q++];
                const float4 sp = geom[s];
                if (kCount) cnt.spheres += 1;
                const float t = sphere_root(o, d, a, inv_a, sp.x, sp.y, sp.z, sp.w, kTMin, tbest);
                if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, sp.x, sp.y, sp.z, sp.w))) { tbest = t; prim = (int)s; }
            }
        } else if (!at_sphere) {
            grid_ray_advance(g, start, tbest, r);
            if (kCount && r.alive) cnt.nodes += 1;
        }
    }
    t_out = tbest;
    prim_out = prim;
}

// Scenes too large for shared memory: canonical 128-byte wide nodes from L2/HBM (vn_math.cuh::wide_global_step), same vote.
template <bool kCount>
__device__ __forceinline__ void closest_hit_wide_global_vote(const float4* __restrict__ wide, const float4* __restrict__ geom, uint32_t root_link,
                                                             uint32_t leaf_vote, f3 o, f3 d, float& t_out, int& prim_out, TraceCounters& cnt,
                                                             float tbest0, int prim0, bool gate) {
    float tbest = tbest0;
    int prim = prim0;
    const f3 idir = slab_idir(d);
    const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const float a = dot(d, d);
    const float inv_a = rcp(a);
    uint32_t stack[kWideStackSize];
    int sp = 0;
    uint32_t cur = root_link;
    while (cur != kEmptyScene) {
        const bool at_leaf = (cur & kLeafFlag) != 0u;
        const unsigned act = __activemask();
        const unsigned lm = __ballot_sync(act, at_leaf);
        if (lm == act || (uint32_t)__popc(lm) >= leaf_vote) {
            if (at_leaf) cur = leaf_step<kCount>(geom, cur, o, d, a, inv_a, tbest, prim, stack, sp, cnt, gate);
        } else if (!at_leaf) {
            if (kCount) cnt.nodes += 1;
            cur = wide_global_step(wide, cur, idir, ood, tbest, stack, sp);
        }
    }
    t_out = tbest;
    prim_out = prim;
}

template <bool kSmem, bool kCount, bool kOct, int kMaxThreads, bool kWide = false, bool kGrid = false>
__global__ void __launch_bounds__(kMaxThreads) k_render_persistent(const __grid_constant__ RenderLaunch p) {
    extern __shared__ float4 s_scene[];
    SceneView sc;
    const uint32_t node_f4s = kWide ? kWideNodeF4 * p.num_wide : 2 * p.num_nodes;
    const uint16_t* g_start = nullptr;
    const uint16_t* g_refs = nullptr;
    if (kGrid) {
        // stage cell starts | references | geom | mat | type (RTIOW: 4.6 + 3.4 + 15.6 KB): the grid needs no node array at all
        uint16_t* s_idx = reinterpret_cast<uint16_t*>(s_scene);
        const uint32_t n_idx = p.grid.n_cells + 1u + p.grid.n_refs;
        float4* s_geom = s_scene + (((size_t)n_idx * 2 + 15) / 16);
        float4* s_mat = s_geom + p.num_spheres;
        uint8_t* s_type = reinterpret_cast<uint8_t*>(s_mat + p.num_spheres);
        for (uint32_t i = threadIdx.x; i <= p.grid.n_cells; i += blockDim.x) s_idx[i] = p.grid_start[i];
        for (uint32_t i = threadIdx.x; i < p.grid.n_refs; i += blockDim.x) s_idx[p.grid.n_cells + 1u + i] = p.grid_refs[i];
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) { s_geom[i] = p.geom[i]; s_mat[i] = p.mat[i]; s_type[i] = p.type[i]; }
        __syncthreads();
        g_start = s_idx; g_refs = s_idx + p.grid.n_cells + 1u;
        sc.nodes = nullptr; sc.geom = s_geom; sc.mat = s_mat; sc.type = s_type;
    } else if (kSmem) {
        // stage nodes | geom | mat | type into shared memory with 128-bit copies
        float4* s_nodes = s_scene;
        float4* s_geom = s_nodes + (size_t)node_f4s * ((kOct || kWide) ? 8 : 1);
        float4* s_mat = s_geom + p.num_spheres;
        uint8_t* s_type = reinterpret_cast<uint8_t*>(s_mat + p.num_spheres);
        if (kWide) {
            // 8 octant-specialised copies of the 4-wide nodes: children sorted front to back for that octant, planes in
            // near/far form (lbvh_core.cuh::wide_octant_node); one thread per (octant, node)
            for (uint32_t i = threadIdx.x; i < 8u * p.num_wide; i += blockDim.x) {
                const uint32_t k = i / p.num_wide, j = i - k * p.num_wide;
                float4 canon[8], out[kWideNodeF4];
#pragma unroll
                for (int q = 0; q < 8; q++) canon[q] = p.wide[8ull * j + q];
                wide_octant_node(canon, k, out);
#pragma unroll
                for (int q = 0; q < (int)kWideNodeF4; q++) s_nodes[(size_t)k * node_f4s + kWideNodeF4 * j + q] = out[q];
            }
        } else if (kOct) {
            // 8 copies of the node array, one per ray-direction octant, in near/far-plane form: for octant k the first
            // float4 of a node holds the planes a ray of that octant enters through (hi where the direction component
            // is negative), the second the planes it leaves through.  227 KB of shared memory buys the removal of the
            // six per-axis min/max from every box test (box_hit_oct).
            for (uint32_t i = threadIdx.x; i < 8u * p.num_nodes; i += blockDim.x) {
                const uint32_t k = i / p.num_nodes, j = i - k * p.num_nodes;
                const float4 lo = p.nodes[2 * j], hi = p.nodes[2 * j + 1];
                float4 nr = lo, fr = hi;
                if (k & 1u) { nr.x = hi.x; fr.x = lo.x; }
                if (k & 2u) { nr.y = hi.y; fr.y = lo.y; }
                if (k & 4u) { nr.z = hi.z; fr.z = lo.z; }
                s_nodes[(size_t)k * node_f4s + 2 * j] = nr;
                s_nodes[(size_t)k * node_f4s + 2 * j + 1] = fr;
            }
        } else {
            for (uint32_t i = threadIdx.x; i < node_f4s; i += blockDim.x) s_nodes[i] = p.nodes[i];
        }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) { s_geom[i] = p.geom[i]; s_mat[i] = p.mat[i]; }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) s_type[i] = p.type[i];
        __syncthreads();
        sc.nodes = s_nodes; sc.geom = s_geom; sc.mat = s_mat; sc.type = s_type;
    } else {
        sc.nodes = p.nodes; sc.geom = p.geom; sc.mat = p.mat; sc.type = p.type;
    }
    sc.root_link = p.root_link;

    uint32_t pix = 0, px = 0, py = 0, cam_seed = 0, s_left = 0;
    bool has_pixel = false, active = false;
    f3 sum = mk3(0.0f);
    PathState st;
    st.o = st.d = st.thr = mk3(0.0f); st.seed = 0; st.depth = 0;
    uint32_t n_seg = 0, n_path = 0, seg0 = 0;    // seg0: n_seg when the current pixel started (tile cost feedback)
    TraceCounters cnt{0u, 0u};
    unsigned long long n_nodes = 0, n_sph = 0;

    for (;;) {
        if (!active) {
            if (has_pixel && s_left == 0u) {
                finish_pixel(p, pix, sum);
                if (p.tile_cost) { uint32_t* tc = p.tile_cost + ((py - p.row_begin) >> 2) * p.tiles_x + (px >> 3); atomicAdd(tc, n_seg - seg0); atomicMax(tc + p.tile_cost_stride, n_seg - seg0); }
                has_pixel = false;
            }
            if (!has_pixel) {
                bool got = false;
                for (;;) {
                    const uint32_t w = fetch_work(p.work_counter);
                    if (w >= p.total_work) break;
                    uint32_t tile = w >> 5;
                    const uint32_t in = w & 31u;
                    if (p.tile_order) tile = __ldg(p.tile_order + tile);
                    uint32_t ty, tx;
                        tile_row_col(tile, p.tiles_x, p.tiles_x_inv, ty, tx);
                    px = tx * 8u + (in & 7u);
                    py = p.row_begin + ty * 4u + (in >> 3);
                    if (px < p.width && py < p.row_end) { got = true; break; }
                }
                if (!got) break;                                  // no pixels left: this lane retires
                pix = py * p.width + px;
                cam_seed = tea4(pix, p.subframe_index);           // RayTracer.cu:169
                seg0 = n_seg;
                sum = mk3(0.0f);
                s_left = p.spp;
                has_pixel = true;
            }
            camera_ray(p.cam, px, py, cam_seed, st.o, st.d);      // RayTracer.cu:173-177
            st.thr = mk3(1.0f);
            st.seed = cam_seed;                                   // prd.seed = seed: a copy (RayTracer.cu:183)
            st.depth = (int)p.max_depth - 1;                      // RayTracer.cu:184
            s_left -= 1u;
            active = true;
            n_path += 1u;
        }
        float t;
        int prim;
        if (kGrid) closest_hit_grid_vote<kCount>(p.grid, g_start, g_refs, sc.geom, p.grid_vote, st.o, st.d, t, prim, cnt, p.gate != 0u);
        else if (kWide) {
            // the huge spheres first (RTIOW: the ground): one convergent sphere test instead of a leaf turn per ray
            float t0 = kTMax;
            int prim0 = -1;
            if (p.huge.n) {
                const float a = dot(st.d, st.d), inv_a = rcp(a);
                for (uint32_t i = 0; i < p.huge.n; i++) {
                    const uint32_t hs = p.huge.idx[i];
                    const float4 g = sc.geom[hs];
                    if (kCount) cnt.spheres += 1;
                    const float th = sphere_root(st.o, st.d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, t0);
                    if (th >= 0.0f) { t0 = th; prim0 = (int)hs; }
                }
            }
            if (!kSmem) {
                if (p.leaf_vote) closest_hit_wide_global_vote<kCount>(p.wide, sc.geom, p.wide_root, p.leaf_vote, st.o, st.d, t, prim, cnt, t0, prim0, p.gate != 0u);
                else closest_hit_wide_global<kCount>(p.wide, sc.geom, p.wide_root, st.o, st.d, t, prim, cnt, t0, prim0, p.gate != 0u);
            }
            else if (p.leaf_vote) closest_hit_wide_vote<kCount>(sc.nodes, node_f4s, sc.geom, p.wide_root, p.leaf_vote, st.o, st.d, t, prim, cnt, t0, prim0);
            else closest_hit_wide<kCount>(sc.nodes, node_f4s, sc.geom, p.wide_root, st.o, st.d, t, prim, cnt, t0, prim0);
        }
        else closest_hit<kCount, kOct>(sc.nodes, sc.geom, sc.root_link, st.o, st.d, t, prim, cnt, node_f4s, p.gate != 0u);
        n_seg += 1u;
        if (kCount) { n_nodes += cnt.nodes; n_sph += cnt.spheres; cnt.nodes = 0; cnt.spheres = 0; }
        f3 result;
        if (!shade_segment(sc, st, t, prim, result)) {
            sum = sum + result;                                   // pixel_color += prd.attenuation (RayTracer.cu:203)
            active = false;
        }
    }
    // per-lane totals -> global counters
    {
        unsigned long long seg = n_seg, path = n_path;
        cg::coalesced_group g = cg::coalesced_threads();
        seg = cg::reduce(g, seg, cg::plus<unsigned long long>());
        path = cg::reduce(g, path, cg::plus<unsigned long long>());
        if (kCount) {
            n_nodes = cg::reduce(g, n_nodes, cg::plus<unsigned long long>());
            n_sph = cg::reduce(g, n_sph, cg::plus<unsigned long long>());
        }
        if (g.thread_rank() == 0) {
            atomicAdd(&p.counters[0], seg);
            atomicAdd(&p.counters[1], path);
            if (kCount) { atomicAdd(&p.counters[2], n_nodes); atomicAdd(&p.counters[3], n_sph); }
        }
    }
}

// k_render_async: the wide-node path kernel with ASYNCHRONOUS shading.  In k_render_persistent every lane of a warp traces one
// segment per round and the round lasts as long as its slowest ray (RTIOW: 4.8 node steps per ray, but 13.7 node turns per round);
// here a lane keeps its traversal state (current link, stack, closest hit) across the shading code, so the warp can leave the
// traversal as soon as `async_done` lanes have finished, shade those and give them their next ray while the stragglers simply
// continue in the next burst.  Inside the burst the turns are voted like closest_hit_wide_vote: a leaf turn when `async_leaf`
// lanes wait at a leaf (or nobody stands on a node), else node steps for as long as `async_node` lanes stand on nodes.  Every lane
// still walks its own pixel's samples in order and runs exactly the per-ray steps of closest_hit_wide(), so the result is
// bit-identical to k_render_persistent; only the order in which a warp's lanes take their turns differs (tools/simt_sim_async.cpp
// is the model the thresholds came from).  Lanes that run out of pixels stay in the loop as zombies until the whole warp is done,
// which keeps every vote a full-mask __ballot_sync.

template <bool kCount, int kMaxThreads, bool kPhase = false, bool kWarpTile = false>
__global__ void __launch_bounds__(kMaxThreads) k_render_async(const __grid_constant__ RenderLaunch p) {
    extern __shared__ float4 s_scene[];
    SceneView sc;
    const uint32_t node_f4s = kWideNodeF4 * p.num_wide;
    {
        float4* s_nodes = s_scene;
        float4* s_geom = s_nodes + (size_t)node_f4s * 8;
        float4* s_mat = s_geom + p.num_spheres;
        uint8_t* s_type = reinterpret_cast<uint8_t*>(s_mat + p.num_spheres);
        for (uint32_t i = threadIdx.x; i < 8u * p.num_wide; i += blockDim.x) {
            const uint32_t k = i / p.num_wide, j = i - k * p.num_wide;
            float4 canon[8], out[kWideNodeF4];
#pragma unroll
            for (int q = 0; q < 8; q++) canon[q] = p.wide[8ull * j + q];
            wide_octant_node(canon, k, out);
#pragma unroll
            for (int q = 0; q < (int)kWideNodeF4; q++) s_nodes[(size_t)k * node_f4s + kWideNodeF4 * j + q] = out[q];
        }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) { s_geom[i] = p.geom[i]; s_mat[i] = p.mat[i]; }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) s_type[i] = p.type[i];
        __syncthreads();
        sc.nodes = s_nodes; sc.geom = s_geom; sc.mat = s_mat; sc.type = s_type;
    }
    sc.root_link = p.root_link;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    // instrumented variant: launch timeline in counters[8..11) (the words the slot kernel uses for its scheduler statistics, read with
    // vn_read_sched_counters): ~min start, ~min time at which a lane found the tickets exhausted, max warp end -- tools/tail_probe.py
    // turns them into the share of a launch spent draining (lanes idle behind the last pixels)
    if (kCount && threadIdx.x == 0) atomicMax(&p.counters[8], ~global_ns());

    uint32_t pix = 0, px = 0, py = 0, cam_seed = 0, s_left = 0;
    bool has_pixel = false, active = false, retired = false;
    f3 sum = mk3(0.0f);
    PathState st;
    st.o = st.d = st.thr = mk3(0.0f); st.seed = 0; st.depth = 0;
    uint32_t n_seg = 0, n_path = 0, seg0 = 0;    // seg0: n_seg when the current pixel started (tile cost feedback)
    TraceCounters cnt{0u, 0u};
    // traversal state, alive across the shading of other lanes
    uint32_t stack[kStackSize];
    const uint32_t base = (uint32_t)__cvta_generic_to_local(stack);
    uint32_t top = base;
    uint32_t cur = kEmptyScene;
    float tbest = kTMax;
    int prim = -1;
    SlabScale ss;
    ss.sdir = ss.nsood = mk3(0.0f);
    WideBase wb = wide_base(sc.nodes, 0u, node_f4s);
    const uint32_t t_node = p.async_node, t_leaf = p.async_leaf;
    uint32_t w_tile = 0u, w_cursor = 32u;          // kWarpTile: the warp's current tile and the next pixel of it to hand out (warp-uniform)

    for (;;) {
        if (cur == kEmptyScene && !retired) {
            if (active) {                                         // a finished traversal: shade it
                n_seg += 1u;
                f3 result;
                if (!shade_segment(sc, st, tbest, prim, result)) {
                    sum = sum + result;                           // pixel_color += prd.attenuation (RayTracer.cu:203)
                    active = false;
                }
            }
            if (!active) {
                if (has_pixel && s_left == 0u) {
                    finish_pixel(p, pix, sum);
                    if (p.tile_cost) { uint32_t* tc = p.tile_cost + ((py - p.row_begin) >> 2) * p.tiles_x + (px >> 3); atomicAdd(tc, n_seg - seg0); atomicMax(tc + p.tile_cost_stride, n_seg - seg0); }
                    has_pixel = false;
                }
                if (!kWarpTile && !has_pixel) {
                    bool got = false;
                    for (;;) {
                        const uint32_t w = fetch_work(p.work_counter);
                        if (w >= p.total_work) break;
                        uint32_t tile = w >> 5;
                        const uint32_t in = w & 31u;
                        if (p.tile_order) tile = __ldg(p.tile_order + tile);
                        uint32_t ty, tx;
                        tile_row_col(tile, p.tiles_x, p.tiles_x_inv, ty, tx);
                        px = tx * 8u + (in & 7u);
                        py = p.row_begin + ty * 4u + (in >> 3);
                        if (px < p.width && py < p.row_end) { got = true; break; }
                    }
                    if (got) {
                        pix = py * p.width + px;
                        cam_seed = tea4(pix, p.subframe_index);   // RayTracer.cu:169
                        seg0 = n_seg;
                        sum = mk3(0.0f);
                        s_left = p.spp;
                        has_pixel = true;
                    } else { retired = true; if (kCount) atomicMax(&p.counters[9], ~global_ns()); }   // no pixels left: this lane only votes from now on
                }
                if (!kWarpTile && !retired) {
                    camera_ray(p.cam, px, py, cam_seed, st.o, st.d);   // RayTracer.cu:173-177
                    st.thr = mk3(1.0f);
                    st.seed = cam_seed;                           // prd.seed = seed: a copy (RayTracer.cu:183)
                    st.depth = (int)p.max_depth - 1;              // RayTracer.cu:184
                    s_left -= 1u;
                    active = true;
                    n_path += 1u;
                }
            }
            if (!kWarpTile && !retired) {
                // start the next segment: the huge spheres first (lbvh_core.cuh::HugeList), then the traversal constants
                tbest = kTMax;
                prim = -1;
                if (p.huge.n) {
                    const float a = dot(st.d, st.d), inv_a = rcp(a);
                    for (uint32_t i = 0; i < p.huge.n; i++) {
                        const uint32_t hs = p.huge.idx[i];
                        const float4 g = sc.geom[hs];
                        if (kCount) cnt.spheres += 1;
                        const float th = sphere_root(st.o, st.d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                        if (th >= 0.0f) { tbest = th; prim = (int)hs; }
                    }
                }
                const f3 idir = slab_idir(st.d);
                wb = wide_base(sc.nodes, ray_octant(st.d), node_f4s);
                ss = slab_scale(idir, mk3(st.o.x * idir.x, st.o.y * idir.y, st.o.z * idir.z), tbest);
                top = base;
                cur = p.wide_root;
            }
        }
        if (kWarpTile) {
            // Warp-owned tiles: the warp takes a whole 8x4 tile with ONE global ticket and hands its 32 pixels to its own lanes, in
            // order, as they become free -- the lanes of a warp then always hold pixels of the same one or two tiles (coherent
            // primary rays, similar materials) instead of whatever tickets happened to be next when each lane asked
            // (tools/simt_sim_async.cpp, WARPTILE=1: 49.2 -> 46.1 model warp-instructions per segment on top of the cost order).
            const bool wants_ray = cur == kEmptyScene && !retired;      // shaded above (or fresh): needs its next segment
            bool need = wants_ray && !active && !has_pixel;
            unsigned m = __ballot_sync(kFull, need);
            while (m) {
                if (w_cursor >= 32u) {                                  // warp-uniform: the current tile is handed out
                    uint32_t t = 0u;
                    if ((threadIdx.x & 31u) == 0u) t = atomicAdd(p.work_counter, 1u);
                    t = __shfl_sync(kFull, t, 0);
                    if (t >= (p.total_work >> 5)) {                     // no tiles left: the asking lanes only vote from now on
                        if (need) { retired = true; if (kCount) atomicMax(&p.counters[9], ~global_ns()); }
                        break;
                    }
                    w_tile = p.tile_order ? __ldg(p.tile_order + t) : t;
                    w_cursor = 0u;
                }
                const uint32_t in = w_cursor + (uint32_t)__popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
                if (need && in < 32u) {
                    uint32_t ty, tx;
                    tile_row_col(w_tile, p.tiles_x, p.tiles_x_inv, ty, tx);
                    px = tx * 8u + (in & 7u);
                    py = p.row_begin + ty * 4u + (in >> 3);
                    if (px < p.width && py < p.row_end) {
                        need = false;
                        pix = py * p.width + px;
                        cam_seed = tea4(pix, p.subframe_index);   // RayTracer.cu:169
                        seg0 = n_seg;
                        sum = mk3(0.0f);
                        s_left = p.spp;
                        has_pixel = true;
                    }
                }
                w_cursor = min(32u, w_cursor + (uint32_t)__popc(m));
                m = __ballot_sync(kFull, need);
            }
            if (wants_ray && !retired) {
                if (!active) {
                    camera_ray(p.cam, px, py, cam_seed, st.o, st.d);   // RayTracer.cu:173-177
                    st.thr = mk3(1.0f);
                    st.seed = cam_seed;                           // prd.seed = seed: a copy (RayTracer.cu:183)
                    st.depth = (int)p.max_depth - 1;              // RayTracer.cu:184
                    s_left -= 1u;
                    active = true;
                    n_path += 1u;
                }
                tbest = kTMax;
                prim = -1;
                if (p.huge.n) {
                    const float a = dot(st.d, st.d), inv_a = rcp(a);
                    for (uint32_t i = 0; i < p.huge.n; i++) {
                        const uint32_t hs = p.huge.idx[i];
                        const float4 g = sc.geom[hs];
                        if (kCount) cnt.spheres += 1;
                        const float th = sphere_root(st.o, st.d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                        if (th >= 0.0f) { tbest = th; prim = (int)hs; }
                    }
                }
                const f3 idir = slab_idir(st.d);
                wb = wide_base(sc.nodes, ray_octant(st.d), node_f4s);
                ss = slab_scale(idir, mk3(st.o.x * idir.x, st.o.y * idir.y, st.o.z * idir.z), tbest);
                top = base;
                cur = p.wide_root;
            }
        }
        const unsigned live = __ballot_sync(kFull, !retired);
        if (live == 0u) break;
        const uint32_t n_live = (uint32_t)__popc(live);
        const uint32_t t_done = p.async_done < n_live ? p.async_done : n_live;
        if (kPhase) {
            // phase form: the while-while phases of closest_hit_wide() without any vote inside them -- every lane descends until it
            // holds a leaf (divergent loop, no ballot), the lanes at a leaf test it -- and ONE vote per phase: the burst ends as soon
            // as t_done lanes are finished.  A round of the persistent kernel spends 5.8 of its 13.2 node turns and 1.7 of its 2.7
            // leaf turns in second and later phases that only ~4 lanes take part in; here those stragglers ride along with the next
            // burst of everybody else.
            for (;;) {
                while ((cur & kLeafFlag) == 0u) {
                    if (kCount) cnt.nodes += 1;
                    cur = wide_node_step_dev(wb, cur, ss, top, base);
                }
                if (cur != kEmptyScene) {
                    const float a = dot(st.d, st.d), t_before = tbest;
                    leaf_test<kCount>(sc.geom, cur, st.o, st.d, a, rcp(a), tbest, prim, cnt);
                    cur = stack_pop_dev(top, base);
                    if (tbest != t_before) {
                        const float g = t_before * rcp_approx(tbest);
                        ss.sdir = ss.sdir * g; ss.nsood = ss.nsood * g;
                    }
                }
                if ((uint32_t)__popc(__ballot_sync(kFull, cur == kEmptyScene && !retired)) >= t_done) break;
            }
            continue;
        }
        // traversal burst: voted node / leaf turns until t_done lanes hold a finished ray
        for (;;) {
            const bool done = cur == kEmptyScene;
            const bool at_leaf = !done && (cur & kLeafFlag) != 0u;
            const uint32_t nd = (uint32_t)__popc(__ballot_sync(kFull, done && !retired));
            const uint32_t nl = (uint32_t)__popc(__ballot_sync(kFull, at_leaf));
            if (nd >= t_done) break;
            if (nl >= t_leaf || nd + nl == n_live) {
                if (at_leaf) {
                    const float a = dot(st.d, st.d), t_before = tbest;
                    leaf_test<kCount>(sc.geom, cur, st.o, st.d, a, rcp(a), tbest, prim, cnt);
                    cur = stack_pop_dev(top, base);
                    if (tbest != t_before) {                      // the slab test works in units of tbest: rescale
                        const float g = t_before * rcp_approx(tbest);
                        ss.sdir = ss.sdir * g; ss.nsood = ss.nsood * g;
                    }
                }
            } else {
                for (;;) {
                    const bool at_node = (cur & kLeafFlag) == 0u;
                    if (at_node) {
                        if (kCount) cnt.nodes += 1;
                        cur = wide_node_step_dev(wb, cur, ss, top, base);
                    }
                    if ((uint32_t)__popc(__ballot_sync(kFull, (cur & kLeafFlag) == 0u)) < t_node) break;
                }
            }
        }
    }
    if (kCount && (threadIdx.x & 31u) == 0u) atomicMax(&p.counters[10], global_ns());
    {
        unsigned long long seg = n_seg, path = n_path, nn = cnt.nodes, ns = cnt.spheres;
        cg::coalesced_group g = cg::coalesced_threads();
        seg = cg::reduce(g, seg, cg::plus<unsigned long long>());
        path = cg::reduce(g, path, cg::plus<unsigned long long>());
        if (kCount) {
            nn = cg::reduce(g, nn, cg::plus<unsigned long long>());
            ns = cg::reduce(g, ns, cg::plus<unsigned long long>());
        }
        if (g.thread_rank() == 0) {
            atomicAdd(&p.counters[0], seg);
            atomicAdd(&p.counters[1], path);
            if (kCount) { atomicAdd(&p.counters[2], nn); atomicAdd(&p.counters[3], ns); }
        }
    }
}

// k_render_lean: k_render_async in its default form (while-while phases, ONE vote per phase, warps own whole 8x4 tiles) rebuilt around
// what ncu said the warps of that kernel wait for (profiles/r01s3_path_kernel_ncu.txt: long_scoreboard 2.7 warps per issue cycle): LOCAL
// MEMORY.  At 1024 threads a lane has 64 registers; k_render_async keeps 38 words of path / pixel / traversal state alive across a node
// step that itself needs 34, so six words (the pixel's sum, its sample counter, the camera seed, the segment counter) were spilled, and
// the 64-word traversal stack lives in local memory as well -- 32 warps x (6 + ~10 stack levels) x 128-byte lines against the 28 KB of
// L1 that 227 KB of shared memory leave: a quarter of those loads went to L2.  Here
//   * the per-lane statistics are gone: segments and paths are counted per warp (popc of the vote that already exists), the per-pixel
//     segment count for the cost-ordered tile schedule only exists in the kCost variant (the first launch of a view);
//   * pixel coordinates travel as one word, the lane's scheduling state as one word;
//   * links are 16 bits (vn_math.cuh::link16): two stack entries per word, a sentinel instead of the emptiness test, and the next node
//     is chosen in registers (wide_node_step16_dev) -- the stack is only read when no child was hit.
// Every lane walks its pixel's samples in order and every ray takes exactly the steps of closest_hit_wide(): the accumulation buffer
// is bit-identical to k_render_persistent / k_render_async (tests/test_gpu_parity.py).  Needs spp, max_depth, width, height < 65536 and
// < 4096 spheres (vn_api.cu falls back to k_render_async otherwise).
// Drain (sample stealing).  Once the tile tickets are exhausted a warp's lanes run dry one by one while the warp keeps paying whole
// iterations for the few lanes that still hold pixels -- measured (tools/tail_probe.py profile=1): the last 0.3 ms of a 5.4 ms launch, 5 % of
// its lane-time, the very end being single pixels whose 16 paths bounce through glass for 150 segments.  From then on a lane without work
// takes over ONE SAMPLE of a pixel another lane of its warp still holds: the camera seed chain (RayTracer.cu:169-183: prd.seed is a copy, the
// path's draws never feed back into it) is replayed without tracing up to that sample and the path is traced as usual.  Samples are taken
// from the END of the owner's range (most samples left first), so a split pixel is a prefix the owner sums in its registers, in order,
// plus a suffix of single samples.  Nobody waits for anybody: the helper writes its sample's radiance into the owner's scratch slot
// (RenderLaunch::steal_scratch, one slot of spp + 1 float4 per lane of the grid -- a lane owns at most one pixel after the tickets ran out), the
// owner writes its prefix sum there when it has started all the samples it kept, and whoever arrives last (one atomic counter per slot) adds
// the suffix to the prefix in sample order -- pixel_color += prd.attenuation (RayTracer.cu:203) sees the samples in the reference's order, so the
// accumulation buffer keeps its bits.  s_left: [15:0] samples the lane still has to start itself, [25:16] helper: its sample index / owner:
// samples given away, [30:26] helper: the owner's lane, [31] helper.
enum LeanState : uint32_t { kLaneNoPixel = 0u, kLaneIdle = 1u, kLaneActive = 2u, kLaneRetired = 3u };
// the draws of one get_ray (camera_ray: jitter u, v and the lens rejection loop) without the ray
__device__ __forceinline__ void camera_skip(uint32_t& seed) {
    (void)rnd(seed); (void)rnd(seed);
    float dx, dy;
    random_in_unit_disk(seed, dx, dy);
}

// the warp's statistics at the end of k_render_lean (or of its drain)
template <bool kCount>
__device__ __forceinline__ void lean_epilogue(const RenderLaunch& p, const TraceCounters& cnt, uint32_t w_seg, uint32_t w_path) {
    if (kCount && (threadIdx.x & 31u) == 0u) atomicMax(&p.counters[10], global_ns());
    if (kCount) {
        unsigned long long nn = cnt.nodes, ns = cnt.spheres;
        cg::coalesced_group g = cg::coalesced_threads();
        nn = cg::reduce(g, nn, cg::plus<unsigned long long>());
        ns = cg::reduce(g, ns, cg::plus<unsigned long long>());
        if (g.thread_rank() == 0) { atomicAdd(&p.counters[2], nn); atomicAdd(&p.counters[3], ns); }
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&p.counters[0], (unsigned long long)w_seg);
        atomicAdd(&p.counters[1], (unsigned long long)w_path);
    }
}
// The drain of k_render_lean: what a warp does once it has found the tile tickets exhausted (see the comment above LeanState).  First the
// traversals in flight are finished, then the warp works in rounds -- shade, hand pixels in, take samples over, start paths, trace every
// active lane's segment to its end -- so that no lane carries traversal state (stack pointer, node, slab constants: 12 registers) across the
// code that redistributes the samples.  With that code inside the first loop, or behind a call, the compiler took the first loop's
// warp-uniform counters out of the uniform registers and spilled the pixel's sum (measured: -1.7 % on the whole launch); the rounds cost the
// drain a little lane utilisation instead, where most lanes idle anyway.
// Sample-range units (k_render_lean<kGlobal>).  A pixel of a million-sphere scene costs a lane 16 samples x 23 segments x a dozen dependent
// L2 fetches each: ~20 ms, one tenth of the launch -- whatever the tile order, the launch ends with whole pixels started late (measured, 1 M
// spheres: tickets exhausted after 76 % of the launch; 86 % with the sample-stealing drain).  So the work items of those scenes are
// (tile, unit): a tile's spp samples are cut into 2, 4, 8 or 16 ranges, handed out as separate tickets -- all first ranges, then all second ranges, ...
// A unit hands the pixel on through memory: {sum so far, camera seed} in `carry`, and a per-pixel flag that says how many units are done; the
// warp that draws (tile, u) starts it once the flags of the tile's 32 pixels say u (it keeps the ticket and asks again in its next iteration
// otherwise: it never spins, the unit it waits for may be in flight in its own lanes).  The samples of a pixel are still summed in sample
// order by whoever holds the pixel, from the carried sum on: the accumulation buffer keeps its bits.  The lane's unit (0..15) rides in the
// free bits of pxy (frames are < 16384 pixels wide and high): bits 14-15 and 30-31.
__device__ __forceinline__ uint32_t unit_px(uint32_t pxy) { return pxy & 0x3FFFu; }
__device__ __forceinline__ uint32_t unit_py(uint32_t pxy) { return (pxy >> 16) & 0x3FFFu; }
__device__ __forceinline__ uint32_t unit_of(uint32_t pxy) { return ((pxy >> 14) & 3u) | ((pxy >> 28) & 12u); }
__device__ __forceinline__ uint32_t unit_samples(const RenderLaunch& p, uint32_t) { return p.spp >> p.units_log2; }
__device__ __forceinline__ void finish_unit(const RenderLaunch& p, uint32_t pxy, f3 sum, uint32_t seed_after) {
    const uint32_t px = unit_px(pxy), py = unit_py(pxy), u = unit_of(pxy);
    if (u + 1u == (1u << p.units_log2)) { finish_pixel(p, py * p.width + px, sum); return; }
    const uint32_t ry = py - p.row_begin;
    const uint32_t ci = ((ry >> 2) * p.tiles_x + (px >> 3)) * 32u + (ry & 3u) * 8u + (px & 7u);
    // One 32-byte store (STG.256: one L2 sector, written whole) carries the sum, the seed and the tag that says which unit they belong to:
    // the reader needs no fence and the writer none either.  With a flag word behind a __threadfence() every finished unit invalidated the
    // SM's L1 (CCTL.IVALL after the MEMBAR) -- where the upper levels of the node tree live: +60 % launch time on a 2 000-sphere scene.
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %4, %4};" ::"l"(p.carry + 2u * ci), "r"(__float_as_uint(sum.x)), "r"(__float_as_uint(sum.y)),
                 "r"(__float_as_uint(sum.z)), "r"(0u), "r"(seed_after), "r"(p.unit_epoch + u + 1u) : "memory");
}

template <bool kCount, bool kGlobal>
__device__ __forceinline__ void lean_finish_traversal(const RenderLaunch& p, const SceneView& sc, const PathState& st, uint32_t& cur, uint32_t& top, uint32_t& tos,
                                                      float& tbest, int& prim, SlabScale& ss, const WideBase& wb, TraceCounters& cnt) {
    if (kGlobal) {
        while (cur != kEmptyScene) {
            if ((cur & kLeafFlag) == 0u) {
                if (kCount) cnt.nodes += 1;
                cur = p.qnodes ? pair_node_step_q_dev(p.qnodes, cur, ss.sdir, ss.nsood, tbest, top, tos) : pair_node_step_dev(sc.nodes, cur, ss.sdir, ss.nsood, tbest, top, tos);
            } else {
                const float a = dot(st.d, st.d);
                leaf_test<kCount>(sc.geom, cur, st.o, st.d, a, rcp(a), tbest, prim, cnt, p.gate != 0u);
                cur = stack_pop32_dev(top, tos);
            }
        }
    } else {
        for (;;) {
            while ((cur & kLeaf16) == 0u) {
                if (kCount) cnt.nodes += 1;
                cur = wide_node_step16_dev(wb, cur, ss, top, tos);
            }
            if (cur == kDone16) break;
            const float a = dot(st.d, st.d), t_before = tbest;
            leaf_test16<kCount>(sc.geom, cur, st.o, st.d, a, rcp(a), tbest, prim, cnt);
            cur = stack_pop16_dev(top, tos);
            if (tbest != t_before) {                              // the slab test works in units of tbest: rescale
                const float g = t_before * rcp_approx(tbest);
                ss.sdir = ss.sdir * g; ss.nsood = ss.nsood * g;
            }
        }
    }
}
template <bool kCount, bool kCost, bool kGlobal>
__device__ __forceinline__ void lean_drain(const RenderLaunch& p, const SceneView& sc, uint32_t node_f4s, uint32_t base, uint32_t& pxy, uint32_t& cam_seed, uint32_t& s_left,
                                           uint32_t& lane_state, f3& sum, PathState& st, uint32_t& cur, uint32_t& top, uint32_t& tos, float& tbest, int& prim, SlabScale& ss,
                                           WideBase& wb, uint32_t& px_seg, uint32_t& last_seg, unsigned long long t_start) {
    uint32_t w_seg = 0u, w_path = 0u;                             // (the first loop has added its counts to the launch's statistics)
    TraceCounters cnt{0u, 0u};
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr uint32_t kDone = kGlobal ? kEmptyScene : kDone16;
    constexpr uint32_t kWordBytes = kGlobal ? 4u : 2u;
    // (the cost-collecting launch attributes ray segments to the lane's own pixel: no stealing there)
    const bool steal_on = !kCost && p.steal != 0u && p.steal_scratch != nullptr;
    constexpr bool kUnits = kGlobal;                               // sample-range units (see finish_unit): the lane's unit rides in pxy
    volatile uint32_t park[6];
    lean_finish_traversal<kCount, kGlobal>(p, sc, st, cur, top, tos, tbest, prim, ss, wb, cnt);
    for (;;) {
        const bool shading = lane_state == kLaneActive;           // every active lane holds a finished traversal
        {
            const uint32_t n_shade = (uint32_t)__popc(__ballot_sync(kFull, shading));
            w_seg += n_shade;
            if (kCount && p.timeline && n_shade && (threadIdx.x & 31u) == 0u)
                atomicAdd(p.timeline + 3072u + (uint32_t)min((unsigned long long)1023u, (global_ns() - t_start) >> 13), n_shade);
        }
        if (shading) {
            if (kCost || kCount) px_seg += 1u;
            f3 result;
            if (!shade_segment(sc, st, tbest, prim, result)) {
                sum = sum + result;                               // pixel_color += prd.attenuation (RayTracer.cu:203)
                lane_state = kLaneIdle;
            }
        }
        bool retire = false;
        if (lane_state == kLaneIdle && s_left == 0u) {            // a whole pixel (or a whole unit of it)
            const uint32_t px = kUnits ? unit_px(pxy) : (pxy & 0xFFFFu), py = kUnits ? unit_py(pxy) : (pxy >> 16);
            if (kUnits) finish_unit(p, pxy, sum, cam_seed); else finish_pixel(p, py * p.width + px, sum);
            if (kCost) { uint32_t* tc = p.tile_cost + ((py - p.row_begin) >> 2) * p.tiles_x + (px >> 3); atomicAdd(tc, px_seg); atomicMax(tc + p.tile_cost_stride, px_seg); }
            if (kCount) last_seg = px_seg;
            retire = true;
        } else if (lane_state == kLaneIdle && (s_left & 0xFFFFu) == 0u) {
            // a part of a split pixel: hand it in, the last one to arrive sums
            const uint32_t lane0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
            const bool helper = (s_left >> 31) != 0u;
            const uint32_t g = lane0 + (helper ? ((s_left >> 26) & 31u) : (threadIdx.x & 31u));
            float4* slot = p.steal_scratch + (size_t)g * (p.spp + 1u);
            const uint32_t m_s = kUnits ? unit_samples(p, pxy) : p.spp;   // samples of the pixel's unit (indices below are relative to its first)
            uint32_t first = 0xFFFFFFFFu;                         // != 0xFFFFFFFF: this lane arrived last; first sample of the suffix
            if (helper) {
                // (.w: the camera seed behind this sample -- the last sample's is what the pixel's next unit starts from)
                __stcg(slot + 1u + ((s_left >> 16) & 0x3FFu), make_float4(sum.x, sum.y, sum.z, __uint_as_float(cam_seed)));
                __threadfence();
                const uint32_t v = atomicAdd(p.steal_count + g, 1u);
                if (v & 0x80000000u) {                            // the owner has handed its prefix in: slot[0].w = first sample of the suffix
                    __threadfence();
                    const uint32_t j_end = __float_as_uint(__ldcg(slot).w);
                    if ((v & 0x7FFFFFFFu) + 1u == m_s - j_end) first = j_end;
                }
            } else {
                const uint32_t n_out = (s_left >> 16) & 0x3FFu;
                __stcg(slot, make_float4(sum.x, sum.y, sum.z, __uint_as_float(m_s - n_out)));
                __threadfence();
                if (atomicAdd(p.steal_count + g, 0x80000000u) == n_out) first = m_s - n_out;
            }
            if (first != 0xFFFFFFFFu) {
                __threadfence();
                const float4 b = __ldcg(slot);
                f3 total = mk3(b.x, b.y, b.z);
                uint32_t seed_after = 0u;
                for (uint32_t k = first; k < m_s; k++) { const float4 r = __ldcg(slot + 1u + k); total = total + mk3(r.x, r.y, r.z); seed_after = __float_as_uint(r.w); }   // RayTracer.cu:203, in order
                if (kUnits) finish_unit(p, pxy, total, seed_after); else finish_pixel(p, (pxy >> 16) * p.width + (pxy & 0xFFFFu), total);
                p.steal_count[g] = 0u;                            // (a lane owns one pixel in the drain: the slot is not used again in this launch)
            }
            retire = true;
        }
        if (retire) {
            lane_state = kLaneRetired;
            if (kCount && p.timeline) {
                const uint32_t bin = (uint32_t)min((unsigned long long)1023u, (global_ns() - t_start) >> 13);
                atomicAdd(p.timeline + bin, 1u);
                atomicAdd(p.timeline + 1024u + bin, last_seg);
            }
        }
        if (steal_on) {
            unsigned thieves = __ballot_sync(kFull, lane_state == kLaneRetired);
            if (thieves) {
                // samples a lane can give away: the ones it has not started, less the one an idle lane starts itself right below
                uint32_t avail = 0u;
                if ((lane_state == kLaneIdle || lane_state == kLaneActive) && (s_left >> 31) == 0u) avail = (s_left & 0xFFFFu) - (lane_state == kLaneIdle ? 1u : 0u);
                while (thieves) {
                    const uint32_t best = __reduce_max_sync(kFull, (avail << 5) | (threadIdx.x & 31u));
                    if ((best >> 5) < p.steal) break;             // (p.steal >= 1: the fewest samples a lane must have left to give one away)
                    const int d = (int)(best & 31u);
                    const uint32_t n = min((uint32_t)__popc(thieves), best >> 5);
                    const uint32_t d_s = __shfl_sync(kFull, s_left, d), d_seed = __shfl_sync(kFull, cam_seed, d), d_pxy = __shfl_sync(kFull, pxy, d);
                    const uint32_t d_out = (d_s >> 16) & 0x3FFu, d_own = d_s & 0xFFFFu;
                    const uint32_t rank = (uint32_t)__popc(thieves & ((1u << (threadIdx.x & 31u)) - 1u));
                    const uint32_t m_d = kUnits ? unit_samples(p, d_pxy) : p.spp;   // samples of the owner's unit
                    if (lane_state == kLaneRetired && rank < n) {
                        const uint32_t k = m_d - d_out - 1u - rank;       // from the end of the owner's range
                        uint32_t sd = d_seed;                              // the owner's seed stands before sample m_d - d_out - d_own
                        for (uint32_t j = m_d - d_out - d_own; j < k; j++) camera_skip(sd);
                        cam_seed = sd;
                        pxy = d_pxy;
                        sum = mk3(0.0f);
                        s_left = 0x80000000u | ((uint32_t)d << 26) | (k << 16) | 1u;
                        lane_state = kLaneIdle;
                    }
                    if ((int)(threadIdx.x & 31u) == d) { s_left = s_left - n + (n << 16); avail -= n; }
                    thieves = __ballot_sync(kFull, lane_state == kLaneRetired);
                }
            }
        }
        const bool launching = lane_state == kLaneIdle;          // (a lane that just took a sample, or whose path just ended)
        w_path += (uint32_t)__popc(__ballot_sync(kFull, launching));
        if (launching) {
            camera_ray(p.cam, kUnits ? unit_px(pxy) : (pxy & 0xFFFFu), kUnits ? unit_py(pxy) : (pxy >> 16), cam_seed, st.o, st.d);   // RayTracer.cu:173-177
            st.thr = mk3(1.0f);
            st.seed = cam_seed;                                   // prd.seed = seed: a copy (RayTracer.cu:183)
            st.depth = (int)p.max_depth - 1;                      // RayTracer.cu:184
            s_left -= 1u;
            lane_state = kLaneActive;
        }
        if (__ballot_sync(kFull, lane_state == kLaneActive) == 0u) break;
        // the pixel's state waits in local memory while the segment is traced (explicitly: left to the register allocator, the drain's
        // copy of the traversal made it spill the pixel's sum and the warp's counters in the FIRST loop as well)
        park[0] = pxy; park[1] = cam_seed; park[2] = s_left; park[3] = __float_as_uint(sum.x); park[4] = __float_as_uint(sum.y); park[5] = __float_as_uint(sum.z);
        if (lane_state == kLaneActive) {
            // the segment, traced to its end: the huge spheres first (lbvh_core.cuh::HugeList), then the traversal
            tbest = kTMax;
            prim = -1;
            const f3 idir = slab_idir(st.d);
            if (kGlobal) {
                if (p.qnodes) {                                   // quantised pairs: slab parameter of plane q = q * sdir + nsood
                    ss.sdir = mk3(p.q_scale[0] * idir.x, p.q_scale[1] * idir.y, p.q_scale[2] * idir.z);
                    ss.nsood = mk3((p.q_lo[0] - st.o.x) * idir.x, (p.q_lo[1] - st.o.y) * idir.y, (p.q_lo[2] - st.o.z) * idir.z);
                } else {
                    ss.sdir = idir;
                    ss.nsood = mk3(st.o.x * idir.x, st.o.y * idir.y, st.o.z * idir.z);
                }
                cur = p.root_link;
            } else {
                if (p.huge.n) {
                    const float a = dot(st.d, st.d), inv_a = rcp(a);
                    for (uint32_t i = 0; i < p.huge.n; i++) {
                        const uint32_t hs = p.huge.idx[i];
                        const float4 g = sc.geom[hs];
                        if (kCount) cnt.spheres += 1;
                        const float th = sphere_root(st.o, st.d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                        if (th >= 0.0f) { tbest = th; prim = (int)hs; }
                    }
                }
                wb = wide_base(sc.nodes, ray_octant(st.d), node_f4s);
                ss = slab_scale(idir, mk3(st.o.x * idir.x, st.o.y * idir.y, st.o.z * idir.z), tbest);
                cur = p.wide_root;
            }
            top = base + kWordBytes;
            tos = kDone;
            lean_finish_traversal<kCount, kGlobal>(p, sc, st, cur, top, tos, tbest, prim, ss, wb, cnt);
        }
        pxy = park[0]; cam_seed = park[1]; s_left = park[2]; sum = mk3(__uint_as_float(park[3]), __uint_as_float(park[4]), __uint_as_float(park[5]));
    }
    lean_epilogue<kCount>(p, cnt, w_seg, w_path);
}

template <bool kCount, bool kCost, int kMaxThreads, bool kGlobal = false, int kMinBlocks = 1, bool kDrain = false, bool kMulti = false>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) k_render_lean(const __grid_constant__ RenderLaunch p) {
    extern __shared__ float4 s_scene[];
    SceneView sc;
    const uint32_t node_f4s = kWideNodeF4 * p.num_wide;
    if (kGlobal) {
        sc.nodes = p.nodes; sc.geom = p.geom; sc.mat = p.mat; sc.type = p.type;
    } else {
        float4* s_nodes = s_scene;
        float4* s_geom = s_nodes + (size_t)node_f4s * 8;
        float4* s_mat = s_geom + p.num_spheres;
        uint8_t* s_type = reinterpret_cast<uint8_t*>(s_mat + p.num_spheres);
        for (uint32_t i = threadIdx.x; i < 8u * p.num_wide; i += blockDim.x) {
            const uint32_t k = i / p.num_wide, j = i - k * p.num_wide;
            float4 canon[8], out[kWideNodeF4];
#pragma unroll
            for (int q = 0; q < 8; q++) canon[q] = p.wide[8ull * j + q];
            wide_octant_node(canon, k, out);
            out[6] = make_float4(u2f(link16(f2u(out[6].x))), u2f(link16(f2u(out[6].y))), u2f(link16(f2u(out[6].z))), u2f(link16(f2u(out[6].w))));
#pragma unroll
            for (int q = 0; q < (int)kWideNodeF4; q++) s_nodes[(size_t)k * node_f4s + kWideNodeF4 * j + q] = out[q];
        }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) { s_geom[i] = p.geom[i]; s_mat[i] = p.mat[i]; }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) s_type[i] = p.type[i];
        __syncthreads();
        sc.nodes = s_nodes; sc.geom = s_geom; sc.mat = s_mat; sc.type = s_type;
    }
    sc.root_link = p.root_link;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    if (kCount && threadIdx.x == 0) atomicMax(&p.counters[8], ~global_ns());     // launch timeline, see k_render_async

    uint32_t pxy = 0u;                 // py << 16 | px of the lane's pixel
    uint32_t cam_seed = 0u;
    uint32_t sd = 0u;                  // samples of the pixel the lane still has to start << 16 | depth budget left of its path (prd.depth): one register
    uint32_t lane_state = kLaneNoPixel;
    f3 sum = mk3(0.0f);
    PathState st;
    st.o = st.d = st.thr = mk3(0.0f); st.seed = 0; st.depth = 0;
    uint32_t w_seg = 0u, w_path = 0u;  // warp-uniform: segments shaded / camera rays started by this warp
    uint32_t px_seg = 0u;              // kCost: ray segments of the current pixel (tile cost feedback)
    uint32_t last_seg = 0u;            // kCount: ray segments of the pixel this lane finished last (launch timeline)
    const unsigned long long t_start = kCount ? global_ns() : 0ull;
    TraceCounters cnt{0u, 0u};
    // traversal state, alive across the shading of other lanes.  kGlobal: pair nodes from L2 / HBM, 32-bit links (kEmptyScene = done).
    typedef typename std::conditional<kGlobal, uint32_t, uint16_t>::type StackWord;
    constexpr uint32_t kDone = kGlobal ? kEmptyScene : kDone16;
    constexpr uint32_t kLeafBit = kGlobal ? kLeafFlag : kLeaf16;
    constexpr uint32_t kWordBytes = (uint32_t)sizeof(StackWord);
    StackWord stack[kStackSize + 2];
    const uint32_t base = (uint32_t)__cvta_generic_to_local(stack);
    stack[0] = (StackWord)kDone;       // never a live entry: the word a pop of the last entry refills `tos` from
    uint32_t top = base + kWordBytes;
    uint32_t tos = kDone;              // the newest stack entry (vn_math.cuh::wide_node_step16_dev); kDone = the sentinel that ends the traversal
    uint32_t cur = kDone;
    float tbest = kTMax;
    int prim = -1;
    SlabScale ss;                      // shared-memory wide nodes: idir / tbest, -(o * idir) / tbest; kGlobal: idir, o * idir
    ss.sdir = ss.nsood = mk3(0.0f);
    WideBase wb = wide_base(sc.nodes, 0u, node_f4s);
    uint32_t w_ox = 0u, w_oy = 0u, w_cursor = 32u; // the warp's current tile (its first pixel) and the next pixel of it to hand out (warp-uniform)
    constexpr bool kUnits = kGlobal;               // sample-range units (see finish_unit above); runtime switch: p.units_log2
    uint32_t w_item = 0u;                          // kUnits: the warp's current work item (tile | unit << 24) ...
    bool w_pending = false;                        // ... drawn but not started: the tile's previous unit is not complete yet
    uint32_t t_seed = 0u;                          // camera seed of pixel (lane) of that tile
    uint32_t w_sub = 0u;                           // kMulti: the subframe (of this launch) the warp's current tile belongs to

    for (;;) {
        const bool fin = cur == kDone && lane_state != kLaneRetired;       // holds a finished traversal, or no ray at all
        const bool shading = fin && lane_state == kLaneActive;
        {
            const uint32_t n_shade = (uint32_t)__popc(__ballot_sync(kFull, shading));
            w_seg += n_shade;
            if (kCount && p.timeline && n_shade && (threadIdx.x & 31u) == 0u)       // instrumented launches: throughput over time
                atomicAdd(p.timeline + 3072u + (uint32_t)min((unsigned long long)1023u, (global_ns() - t_start) >> 13), n_shade);
        }
        if (shading) {
            if (kCost || kCount) px_seg += 1u;
            f3 result;
            st.depth = (int)(sd & 0xFFFFu);
            if (!shade_segment(sc, st, tbest, prim, result)) {
                sum = sum + result;                               // pixel_color += prd.attenuation (RayTracer.cu:203)
                lane_state = kLaneIdle;
            } else {
                sd -= 1u;                                         // (shade_segment's prd.depth - 1 of a path that goes on, RayTracer.cu:314,361,417)
            }
        }
        if (kMulti) {
            // (a lane whose pixel still waits for its subframe before stays idle with its sum, starts nothing, and asks again next time)
            if (fin && lane_state == kLaneIdle && sd < 0x10000u && finish_pixel_multi(p, pxy, sum)) lane_state = kLaneNoPixel;
        } else
        if (fin && lane_state == kLaneIdle && sd < 0x10000u) {
            const uint32_t px = kUnits ? unit_px(pxy) : (pxy & 0xFFFFu), py = kUnits ? unit_py(pxy) : (pxy >> 16);
            if (kUnits) finish_unit(p, pxy, sum, cam_seed); else finish_pixel(p, py * p.width + px, sum);
            if (kCost) { uint32_t* tc = p.tile_cost + ((py - p.row_begin) >> 2) * p.tiles_x + (px >> 3); atomicAdd(tc, px_seg); atomicMax(tc + p.tile_cost_stride, px_seg); }
            if (kCount) last_seg = px_seg;
            lane_state = kLaneNoPixel;
        }
        // warp-owned tiles (see k_render_async): one global ticket per 8x4 tile, pixels handed to the asking lanes in order
        {
            bool need = fin && lane_state == kLaneNoPixel;
            unsigned m = __ballot_sync(kFull, need);
            while (m) {
                if (w_cursor >= 32u) {
                    uint32_t t = 0u;
                    if (!(kUnits && w_pending)) {
                    if ((threadIdx.x & 31u) == 0u) t = atomicAdd(p.work_counter, 1u);
                    t = __shfl_sync(kFull, t, 0);
                    }
                    if (!(kUnits && w_pending) && t >= (p.total_work >> 5)) {   // no tiles left: the asking lanes only vote from now on (or help, see above)
                        if (kDrain) w_cursor = 33u;               // (> 32: this warp has seen the tickets exhausted and leaves the loop below)
                        if (need) {
                            lane_state = kLaneRetired;
                            if (kCount) {
                                const unsigned long long now = global_ns();
                                atomicMax(&p.counters[9], ~now);
                                if (p.timeline) {
                                    const uint32_t bin = (uint32_t)min((unsigned long long)1023u, (now - t_start) >> 13);
                                    atomicAdd(p.timeline + bin, 1u);
                                    atomicAdd(p.timeline + 1024u + bin, last_seg);
                                }
                            }
                        }
                        break;
                    }
                    if (kMulti) { w_sub = t / p.tiles_per_sub; t -= w_sub * p.tiles_per_sub; }      // tickets are subframe-major
                    if (!(kUnits && w_pending)) w_item = p.tile_order ? __ldg(p.tile_order + t) : t;
                    const uint32_t tile = kUnits ? (w_item & 0x00FFFFFFu) : w_item;
                    // every lane forms the camera seed of "its" pixel of the new tile (pixel i of the tile by lane i) while the warp is
                    // converged; the lanes that take a pixel pick its seed up with one shuffle.  Formed by the taker, the 70 instructions
                    // of tea<4> ran at 1.6 active lanes 1.3 M times per launch (ncu: 2.9 % of all warp instructions), now once per tile.
                    uint32_t ty, tx;
                    tile_row_col(tile, p.tiles_x, p.tiles_x_inv, ty, tx);
                    w_ox = tx * 8u; w_oy = p.row_begin + ty * 4u;
                    if (kUnits && (w_item >> 24) != 0u) {
                        // a later unit of the tile: its pixels' sums and seeds come from the unit before -- once ALL of them are there
                        const uint32_t lpx = w_ox + (threadIdx.x & 7u), lpy = w_oy + ((threadIdx.x & 31u) >> 3);
                        const uint32_t ci = tile * 32u + (threadIdx.x & 31u);
                        bool ready = true;
                        uint32_t c_seed = 0u;
                        if (lpx < p.width && lpy < p.row_end) {
                            uint32_t f;                            // {seed, tag}: the second half of the pixel's 32-byte hand-over (finish_unit), from L2
                            asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(c_seed), "=r"(f) : "l"(p.carry + 2u * ci + 1u) : "memory");
                            ready = (int32_t)(f - (p.unit_epoch + (w_item >> 24))) >= 0;
                        }
                        w_pending = __all_sync(kFull, ready) == 0;
                        if (w_pending) break;                      // (the askers stay without a pixel and ask again in the warp's next iteration)
                        t_seed = c_seed;
                    } else {
                        t_seed = tea4((w_oy + ((threadIdx.x & 31u) >> 3)) * p.width + w_ox + (threadIdx.x & 7u), p.subframe_index + (kMulti ? w_sub * p.sub_stride : 0u));   // RayTracer.cu:169
                    }
                    w_cursor = 0u;
                    if (kCount && p.timeline && (threadIdx.x & 31u) == 0u)
                        atomicAdd(p.timeline + 2048u + (uint32_t)min((unsigned long long)1023u, (global_ns() - t_start) >> 13), 1u);
                }
                const uint32_t in = w_cursor + (uint32_t)__popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
                const uint32_t in_seed = __shfl_sync(kFull, t_seed, (int)(in & 31u));
                if (need && in < 32u) {
                    const uint32_t px = w_ox + (in & 7u), py = w_oy + (in >> 3);
                    if (px < p.width && py < p.row_end) {
                        need = false;
                        pxy = (py << 16) | px;
                        if (kMulti) pxy |= ((w_sub & 7u) << 13) | ((w_sub >> 3) << 29);
                        cam_seed = in_seed;
                        sum = mk3(0.0f);
                        sd = p.spp << 16;
                        if (kUnits && p.units_log2) {
                            const uint32_t u = w_item >> 24;
                            pxy |= ((u & 3u) << 14) | ((u & 12u) << 28);
                            sd = (p.spp >> p.units_log2) << 16;
                            if (u) { const float4 c = __ldcg(p.carry + 2u * ((w_item & 0x00FFFFFFu) * 32u + in)); sum = mk3(c.x, c.y, c.z); }
                        }
                        if (kCost || kCount) px_seg = 0u;
                        lane_state = kLaneIdle;
                    }
                }
                w_cursor = min(32u, w_cursor + (uint32_t)__popc(m));
                m = __ballot_sync(kFull, need);
            }
        }
        const bool launching = fin && lane_state == kLaneIdle && !(kMulti && sd < 0x10000u);   // (a lane that just got a pixel, or whose path just ended)
        // (kUnits: the lanes of a warp whose next unit is not ready yet hold no pixel -- they start nothing and do not count in the votes below)
        const bool starting = fin && lane_state != kLaneRetired && !(kUnits && lane_state == kLaneNoPixel) && !(kMulti && lane_state == kLaneIdle && !launching);
        w_path += (uint32_t)__popc(__ballot_sync(kFull, launching));
        if (starting) {
            if (launching) {
                camera_ray(p.cam, kMulti ? multi_px(pxy) : kUnits ? unit_px(pxy) : (pxy & 0xFFFFu), kMulti ? multi_py(pxy) : kUnits ? unit_py(pxy) : (pxy >> 16), cam_seed, st.o, st.d);   // RayTracer.cu:173-177
                st.thr = mk3(1.0f);
                st.seed = cam_seed;                               // prd.seed = seed: a copy (RayTracer.cu:183)
                sd = ((sd - 0x10000u) & 0xFFFF0000u) | (p.max_depth - 1u);   // one sample less to start; prd.depth = max_depth - 1 (RayTracer.cu:184)
                lane_state = kLaneActive;
            }
            // start the next segment: the huge spheres first (lbvh_core.cuh::HugeList), then the traversal constants
            tbest = kTMax;
            prim = -1;
            const f3 idir = slab_idir(st.d);
            if (kGlobal) {
                if (p.qnodes) {                                   // quantised pairs: slab parameter of plane q = q * sdir + nsood
                    ss.sdir = mk3(p.q_scale[0] * idir.x, p.q_scale[1] * idir.y, p.q_scale[2] * idir.z);
                    ss.nsood = mk3((p.q_lo[0] - st.o.x) * idir.x, (p.q_lo[1] - st.o.y) * idir.y, (p.q_lo[2] - st.o.z) * idir.z);
                } else {
                    ss.sdir = idir;
                    ss.nsood = mk3(st.o.x * idir.x, st.o.y * idir.y, st.o.z * idir.z);
                }
                cur = p.root_link;
            } else {
                if (p.huge.n) {
                    const float a = dot(st.d, st.d), inv_a = rcp(a);
                    for (uint32_t i = 0; i < p.huge.n; i++) {
                        const uint32_t hs = p.huge.idx[i];
                        const float4 g = sc.geom[hs];
                        if (kCount) cnt.spheres += 1;
                        const float th = sphere_root(st.o, st.d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                        if (th >= 0.0f) { tbest = th; prim = (int)hs; }
                    }
                }
                wb = wide_base(sc.nodes, ray_octant(st.d), node_f4s);
                ss = slab_scale(idir, mk3(st.o.x * idir.x, st.o.y * idir.y, st.o.z * idir.z), tbest);
                cur = p.wide_root;
            }
            top = base + kWordBytes;
            tos = kDone;
        }
        const unsigned live = __ballot_sync(kFull, lane_state != kLaneRetired);
        if (live == 0u) break;
        const uint32_t n_live = kUnits ? (uint32_t)__popc(__ballot_sync(kFull, lane_state == kLaneActive)) : (uint32_t)__popc(live);
        const uint32_t t_done = p.async_done < n_live ? p.async_done : n_live;
        if (kGlobal) {
            // Nodes from L2 / HBM: every node step is a dependent fetch of ~300-800 cycles and a ray takes 20 steps to its first leaf in a
            // scene of a million spheres, so the turns are VOTED: all lanes standing on a node step together; the lanes holding a leaf
            // wait until async_leaf of them do (or nobody stands on a node), then test their spheres together; the burst ends when
            // t_done lanes hold a finished ray.  A lane is never idle for longer than the others need to reach their next leaf -- in
            // k_render_persistent every lane of a warp waited for the warp's slowest RAY (ncu on 1 M spheres: 4.8 of 32 lanes active).
            const uint32_t t_leaf = p.async_leaf;
            for (;;) {
                if ((cur & kLeafBit) == 0u) {
                    if (kCount) cnt.nodes += 1;
                    cur = p.qnodes ? pair_node_step_q_dev(p.qnodes, cur, ss.sdir, ss.nsood, tbest, top, tos) : pair_node_step_dev(sc.nodes, cur, ss.sdir, ss.nsood, tbest, top, tos);
                }
                const bool at_leaf = (cur & kLeafBit) != 0u && cur != kDone;
                const unsigned lm = __ballot_sync(kFull, at_leaf);
                const unsigned nm = __ballot_sync(kFull, (cur & kLeafBit) == 0u);
                if (nm == 0u || (uint32_t)__popc(lm) >= t_leaf) {
                    if (at_leaf) {
                        const float a = dot(st.d, st.d);
                        leaf_test<kCount>(sc.geom, cur, st.o, st.d, a, rcp(a), tbest, prim, cnt, p.gate != 0u);
                        cur = stack_pop32_dev(top, tos);
                    }
                }
                if ((uint32_t)__popc(__ballot_sync(kFull, cur == kDone && lane_state == kLaneActive)) >= t_done) break;
            }
            if (kDrain && w_cursor > 32u) break;                           // (see the end of the loop)
            continue;
        }
        // one traversal burst: while-while phases, one vote per phase; it ends as soon as t_done lanes hold a finished ray
        for (;;) {
            while ((cur & kLeaf16) == 0u) {
                if (kCount) cnt.nodes += 1;
                cur = wide_node_step16_dev(wb, cur, ss, top, tos);
            }
            if (cur != kDone16) {
                const float a = dot(st.d, st.d), t_before = tbest;
                leaf_test16<kCount>(sc.geom, cur, st.o, st.d, a, rcp(a), tbest, prim, cnt);
                cur = stack_pop16_dev(top, tos);
                if (tbest != t_before) {                          // the slab test works in units of tbest: rescale
                    const float g = t_before * rcp_approx(tbest);
                    ss.sdir = ss.sdir * g; ss.nsood = ss.nsood * g;
                }
            }
            if ((uint32_t)__popc(__ballot_sync(kFull, cur == kDone16 && lane_state != kLaneRetired)) >= t_done) break;
        }
        if (kDrain && w_cursor > 32u) break;                               // the tickets are exhausted: the rest of the launch is lean_drain()
    }
    lean_epilogue<kCount>(p, cnt, w_seg, w_path);
    uint32_t s_left = sd >> 16;
    st.depth = (int)(sd & 0xFFFFu);
    if (kDrain && w_cursor > 32u) lean_drain<kCount, kCost, kGlobal>(p, sc, node_f4s, base, pxy, cam_seed, s_left, lane_state, sum, st, cur, top, tos, tbest, prim, ss, wb, px_seg, last_seg, t_start);
}

// Sort keys of the cost-ordered tile schedule: most urgent tile first = smallest key.  A tile's urgency is its total cost (ray segments
// of its 32 pixels in one launch, mode 1), its most expensive pixel (mode 2), or the larger of the total / 8 and the most expensive
// pixel (mode 3: a cheap tile that holds one glass pixel must not be handed out last).  24 key bits are plenty.
// Mode 4 (the default): by total cost, but tiles in which NO path ever hit anything (their most expensive pixel cost exactly spp segments: sky)
// go last, whatever their order among themselves.  Their cost is deterministic -- spp segments per pixel, every launch -- whereas a tile that
// was cheap in the collecting launch but touches a glass silhouette can hold a 100-segment pixel in the next one; handed out among the last,
// such pixels were the thin tail that kept a launch alive for its last 0.19 ms (tools/tail_probe.py: 0.6 % of the lanes, 3.6 % of the time).
// In every mode the tiles in which no path hit anything end up last, and *n_all_miss counts them: a split frame (vn_api.cu, "split_tail") hands
// exactly those to its second launch.
__global__ void __launch_bounds__(256) k_tile_keys(const uint32_t* __restrict__ cost, uint32_t stride, uint32_t n, uint32_t mode, uint32_t spp, uint32_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals, uint32_t* __restrict__ n_all_miss) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool all_miss = i < n && cost[stride + i] <= spp;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, all_miss);
    if ((threadIdx.x & 31u) == 0u && m) atomicAdd(n_all_miss, (uint32_t)__popc(m));
    // ... and n_all_miss[2..3] (a 64-bit word) adds the costs up: ray segments of the collecting launch, the view's mean path length for the host
    {
        const uint32_t w_sum = __reduce_add_sync(0xFFFFFFFFu, i < n ? cost[i] : 0u);
        if ((threadIdx.x & 31u) == 0u && w_sum) atomicAdd(reinterpret_cast<unsigned long long*>(n_all_miss + 2), (unsigned long long)w_sum);
    }
    if (i >= n) return;
    const uint32_t sum = cost[i], mx = cost[stride + i];
    uint32_t c = mode == 2u ? mx : (mode == 3u ? max(sum >> 3, mx) : sum);
    if (mode == 4u) c = sum + 1u;
    if (all_miss) c = 0u;
    c = c < 0x00FFFFFFu ? c : 0x00FFFFFFu;
    keys[i] = 0x00FFFFFFu - c;
    vals[i] = i;
}

// Work items of a launch with sample-range units (finish_unit above): all first units in tile order, then all second units, ...
__global__ void __launch_bounds__(256) k_unit_items(const uint32_t* __restrict__ order, uint32_t n, uint32_t units_log2, uint32_t* __restrict__ items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n << units_log2)) return;
    const uint32_t u = i / n, j = i - u * n;
    items[i] = (order ? order[j] : j) | (u << 24);
}

// Kernel (3) of the north star when used stand-alone: image = make_color(accum * scale).  One pixel per thread:
// a warp reads 512 contiguous bytes of float4 and writes 128 contiguous bytes of uchar4.
__global__ void __launch_bounds__(256) k_tonemap(const float4* __restrict__ accum, float scale, uint32_t* __restrict__ image, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 a = accum[i];
        f3 c = mk3(a.x, a.y, a.z);
        if (scale != 1.0f) c = c * scale;
        image[i] = make_color_u32(c);
    }
}

struct PeerList { const float4* p[kMaxPeers]; };
struct FlagList { const uint32_t* f[kMaxPeers]; };

// ---- device-side ordering between GPUs that belong to DIFFERENT processes (one process per GPU, buffers mapped through CUDA IPC): a
// 32-bit epoch flag per rank in device memory.  k_signal publishes everything the stream has written so far (fence, then a release
// store); a waiting kernel on another GPU polls the flag with acquire loads.  This replaces the two host barriers (NCCL all-reduces,
// ~0.4 ms each on 8 GPUs) that used to bracket the once-per-frame reduce.  A poll gives up after kFlagTimeoutNs and reports through
// *error, so that a peer that died cannot hang the others' GPUs.
constexpr unsigned long long kFlagTimeoutNs = 4000000000ull;
__global__ void k_signal(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__device__ __forceinline__ bool wait_flags(const FlagList& flags, uint32_t n, uint32_t value, uint32_t* error) {
    const unsigned long long t0 = global_ns();
    for (uint32_t r = 0; r < n; r++) {
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags.f[r]) : "memory");
            if ((int32_t)(v - value) >= 0) break;                  // epochs only grow (wrap-safe compare)
            if (global_ns() - t0 > kFlagTimeoutNs) { if (error) atomicExch(error, 1u + r); return false; }
            __nanosleep(200);
        }
    }
    return true;
}
__global__ void k_wait_flags(const FlagList flags, uint32_t n, uint32_t value, uint32_t* error) {
    wait_flags(flags, n, value, error);
}

// Fused cross-GPU reduce + tonemap: each GPU owns a slice of pixels, loads that slice from every peer's partial-sum
// buffer over NVLink (peer-mapped pointers), adds them in rank order (deterministic), writes the reduced float4
// into its own accum and the uchar4 pixels into the image.  Replaces ncclReduce + a separate tonemap launch.
__global__ void __launch_bounds__(256) k_reduce_tonemap_peers(const PeerList peers, uint32_t n_peers, float scale, uint64_t begin, uint64_t end,
                                                             float4* __restrict__ accum_out, uint32_t* __restrict__ image,
                                                             const FlagList flags, uint32_t wait_value, uint32_t* error) {
    if (wait_value) {
        // every peer's partial sum must be complete before it is read: one thread per CTA polls the peers' epoch flags
        if (threadIdx.x == 0) wait_flags(flags, n_peers, wait_value, error);
        __syncthreads();
    }
    for (uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (uint64_t)gridDim.x * blockDim.x) {
        f3 s = mk3(0.0f);
        for (uint32_t r = 0; r < n_peers; r++) {
            const float4 a = __ldcg(&peers.p[r][i]);
            s = s + mk3(a.x, a.y, a.z);
        }
        if (accum_out) accum_out[i] = make_float4(s.x, s.y, s.z, 1.0f);
        if (image) image[i] = make_color_u32(scale != 1.0f ? s * scale : s);
    }
}

// ---- unit-test kernels
__global__ void k_test_rng(const uint32_t* __restrict__ v0, const uint32_t* __restrict__ v1, uint64_t n, uint32_t n_draws,
                           uint32_t* __restrict__ seeds, uint32_t* __restrict__ lcg_out, float* __restrict__ rnd_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = tea4(v0[i], v1[i]);
    seeds[i] = s;
    uint32_t s2 = s;
    for (uint32_t k = 0; k < n_draws; k++) {
        lcg_out[i * n_draws + k] = lcg(s);
        rnd_out[i * n_draws + k] = rnd(s2);
    }
}

__global__ void k_trace_rays(const float4* __restrict__ nodes, const float4* __restrict__ geom, uint32_t root_link, const float* __restrict__ o,
                             const float* __restrict__ d, uint64_t n, float* __restrict__ t_out, int32_t* __restrict__ prim_out,
                             const uint32_t* __restrict__ orig, bool gate) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t;
    int prim;
    TraceCounters cnt{0u, 0u};
    closest_hit<false>(nodes, geom, root_link, mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), t, prim, cnt, 0u, gate);
    t_out[i] = prim >= 0 ? t : -1.0f;
    prim_out[i] = prim >= 0 ? (int32_t)orig[prim] : -1;
}

__global__ void k_trace_rays_grid(const GridHeader g, const uint16_t* __restrict__ start, const uint16_t* __restrict__ refs, const float4* __restrict__ geom,
                                  const float* __restrict__ o, const float* __restrict__ d, uint64_t n, float* __restrict__ t_out,
                                  int32_t* __restrict__ prim_out, const uint32_t* __restrict__ orig, bool gate) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t;
    int prim;
    TraceCounters cnt{0u, 0u};
    closest_hit_grid<false>(g, start, refs, geom, mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), t, prim, cnt, gate);
    t_out[i] = prim >= 0 ? t : -1.0f;
    prim_out[i] = prim >= 0 ? (int32_t)orig[prim] : -1;
}

__global__ void k_make_color(const float* __restrict__ rgb, uint64_t n, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_color_u32(mk3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
}

__global__ void k_scatter(uint32_t type, float4 mat, const float* __restrict__ dirs, const float* __restrict__ normals,
                          const uint8_t* __restrict__ front, const uint32_t* __restrict__ seeds, uint64_t n, float* __restrict__ dirs_out,
                          uint8_t* __restrict__ scattered, uint32_t* __restrict__ seeds_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    const f3 nrm = mk3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
    uint32_t seed = seeds[i];
    f3 out = mk3(0.0f);
    bool ok = true;
    if (type == 0u) out = scatter_lambertian(nrm, random_in_unit_sphere(seed));
    else if (type == 1u) ok = scatter_metal(normalize(d), nrm, mat.w, random_in_unit_sphere(seed), out);
    else out = scatter_dielectric(normalize(d), nrm, front[i] != 0, mat.x, seed);
    dirs_out[3 * i] = out.x; dirs_out[3 * i + 1] = out.y; dirs_out[3 * i + 2] = out.z;
    scattered[i] = ok ? 1 : 0;
    seeds_out[i] = seed;
}

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

inline uint32_t grid_for(uint64_t n, int threads) { return (uint32_t)((n + threads - 1) / threads); }

}  // namespace

namespace {
typedef void (*PathKernel)(const RenderLaunch);
PathKernel pick_kernel(bool scene_in_smem, bool count, bool octant, bool wide = false, int threads = 1024, bool grid = false, bool async = false, bool phase = false, bool warp_tiles = false,
                       bool lean = false, bool cost = false, int global_ctas = 4, bool drain = false, bool multi = false) {
    if (lean && !scene_in_smem && !wide && !grid) {        // pair nodes from L2 / HBM, asynchronous (k_render_lean<kGlobal>); 4, 5 or 6 CTAs of 256 threads per SM (64 / 48 / 40 registers)
        if (drain && !cost) {                              // with the sample-stealing drain (lean_drain)
            if (global_ctas >= 6) return count ? k_render_lean<true, false, 256, true, 6, true> : k_render_lean<false, false, 256, true, 6, true>;
            if (global_ctas == 5) return count ? k_render_lean<true, false, 256, true, 5, true> : k_render_lean<false, false, 256, true, 5, true>;
            return count ? k_render_lean<true, false, 256, true, 4, true> : k_render_lean<false, false, 256, true, 4, true>;
        }
        if (global_ctas >= 6) return count ? (cost ? k_render_lean<true, true, 256, true, 6> : k_render_lean<true, false, 256, true, 6>) : (cost ? k_render_lean<false, true, 256, true, 6> : k_render_lean<false, false, 256, true, 6>);
        if (global_ctas == 5) return count ? (cost ? k_render_lean<true, true, 256, true, 5> : k_render_lean<true, false, 256, true, 5>) : (cost ? k_render_lean<false, true, 256, true, 5> : k_render_lean<false, false, 256, true, 5>);
        return count ? (cost ? k_render_lean<true, true, 256, true, 4> : k_render_lean<true, false, 256, true, 4>) : (cost ? k_render_lean<false, true, 256, true, 4> : k_render_lean<false, false, 256, true, 4>);
    }
    if (lean && async && phase && warp_tiles && wide && scene_in_smem && !grid) {
        if (multi && !cost && !count) return threads <= 768 ? k_render_lean<false, false, 768, false, 1, false, true> : k_render_lean<false, false, 1024, false, 1, false, true>;
        if (drain && !cost) {
            if (threads <= 768) return count ? k_render_lean<true, false, 768, false, 1, true> : k_render_lean<false, false, 768, false, 1, true>;
            return count ? k_render_lean<true, false, 1024, false, 1, true> : k_render_lean<false, false, 1024, false, 1, true>;
        }
        if (threads <= 768) return count ? (cost ? k_render_lean<true, true, 768> : k_render_lean<true, false, 768>) : (cost ? k_render_lean<false, true, 768> : k_render_lean<false, false, 768>);
        return count ? (cost ? k_render_lean<true, true, 1024> : k_render_lean<true, false, 1024>) : (cost ? k_render_lean<false, true, 1024> : k_render_lean<false, false, 1024>);
    }
    if (async && wide && scene_in_smem && !grid) {
        if (phase && warp_tiles) {
            if (threads <= 768) return count ? k_render_async<true, 768, true, true> : k_render_async<false, 768, true, true>;
            return count ? k_render_async<true, 1024, true, true> : k_render_async<false, 1024, true, true>;
        }
        if (phase) {
            if (threads <= 768) return count ? k_render_async<true, 768, true> : k_render_async<false, 768, true>;
            return count ? k_render_async<true, 1024, true> : k_render_async<false, 1024, true>;
        }
        if (threads <= 512) return count ? k_render_async<true, 512> : k_render_async<false, 512>;
        if (threads <= 768) return count ? k_render_async<true, 768> : k_render_async<false, 768>;
        return count ? k_render_async<true, 1024> : k_render_async<false, 1024>;
    }
    if (grid && threads <= 512) return count ? k_render_persistent<true, true, false, 512, false, true> : k_render_persistent<true, false, false, 512, false, true>;
    if (grid && threads <= 768) return count ? k_render_persistent<true, true, false, 768, false, true> : k_render_persistent<true, false, false, 768, false, true>;
    if (grid) return count ? k_render_persistent<true, true, false, 1024, false, true> : k_render_persistent<true, false, false, 1024, false, true>;
    if (wide && !scene_in_smem) return count ? k_render_persistent<false, true, false, 256, true> : k_render_persistent<false, false, false, 256, true>;
    if (wide && threads <= 512) return count ? k_render_persistent<true, true, false, 512, true> : k_render_persistent<true, false, false, 512, true>;
    if (wide && threads <= 768) return count ? k_render_persistent<true, true, false, 768, true> : k_render_persistent<true, false, false, 768, true>;
    if (wide) return count ? k_render_persistent<true, true, false, 1024, true> : k_render_persistent<true, false, false, 1024, true>;
    if (scene_in_smem && octant) return count ? k_render_persistent<true, true, true, 1024> : k_render_persistent<true, false, true, 1024>;
    if (scene_in_smem) return count ? k_render_persistent<true, true, false, 256> : k_render_persistent<true, false, false, 256>;
    return count ? k_render_persistent<false, true, false, 256> : k_render_persistent<false, false, false, 256>;
}
}  // namespace

int max_blocks_per_sm(int threads, size_t smem_bytes, bool scene_in_smem, bool count, bool octant, bool wide, bool grid, bool lean, int global_ctas) {
    int nb = 0;
    PathKernel k = pick_kernel(scene_in_smem, count, octant, wide, threads, grid, false, false, false, lean, false, global_ctas);
    if (smem_bytes > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, threads, smem_bytes);
    return e == cudaSuccess ? nb : -1;
}

cudaError_t launch_render_persistent(const RenderLaunch& p, const KernelConfig& cfg, cudaStream_t stream) {
    PathKernel k = pick_kernel(cfg.scene_in_smem, cfg.count, cfg.octant, cfg.wide, cfg.threads, cfg.grid, cfg.async, cfg.async && p.async_node == 0u, cfg.warp_tiles,
                               cfg.lean, p.tile_cost != nullptr, cfg.global_ctas, p.steal != 0u && p.steal_scratch != nullptr, p.n_sub > 1u);
    if (cfg.smem_bytes > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
    }
    k<<<cfg.blocks, cfg.threads, cfg.smem_bytes, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_tile_keys(const uint32_t* cost, uint32_t stride, uint32_t n, uint32_t mode, uint32_t spp, uint32_t* keys, uint32_t* vals, uint32_t* n_all_miss, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_tile_keys<<<(n + 255u) / 256u, 256, 0, stream>>>(cost, stride, n, mode, spp, keys, vals, n_all_miss);
    return cudaGetLastError();
}

cudaError_t launch_unit_items(const uint32_t* order, uint32_t n, uint32_t units_log2, uint32_t* items, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_unit_items<<<((n << units_log2) + 255u) / 256u, 256, 0, stream>>>(order, n, units_log2, items);
    return cudaGetLastError();
}

cudaError_t launch_tonemap(const float4* accum, float scale, uint32_t* image, uint64_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_tonemap<<<(uint32_t)std::min<uint64_t>(grid_for(n, 256), 148u * 16u), 256, 0, stream>>>(accum, scale, image, n);
    return cudaGetLastError();
}

cudaError_t launch_reduce_tonemap_peers(const float4* const* peers, uint32_t n_peers, float scale, uint64_t begin, uint64_t end,
                                        float4* accum_out, uint32_t* image, const uint32_t* const* peer_flags, uint32_t wait_value, uint32_t* error,
                                        cudaStream_t stream) {
    if (end <= begin) return cudaSuccess;
    if (n_peers > (uint32_t)kMaxPeers) return cudaErrorInvalidValue;
    PeerList pl;
    FlagList fl;
    for (int i = 0; i < kMaxPeers; i++) { pl.p[i] = i < (int)n_peers ? peers[i] : nullptr; fl.f[i] = (peer_flags && i < (int)n_peers) ? peer_flags[i] : nullptr; }
    // a grid that is resident at once (the CTAs may poll flags): at most 8 CTAs of 256 threads per SM
    k_reduce_tonemap_peers<<<(uint32_t)std::min<uint64_t>(grid_for(end - begin, 256), 148u * 8u), 256, 0, stream>>>(pl, n_peers, scale, begin, end, accum_out, image,
                                                                                                                   fl, peer_flags ? wait_value : 0u, error);
    return cudaGetLastError();
}

cudaError_t launch_signal(uint32_t* flag, uint32_t value, cudaStream_t stream) {
    k_signal<<<1, 1, 0, stream>>>(flag, value);
    return cudaGetLastError();
}

cudaError_t launch_wait_flags(const uint32_t* const* flags, uint32_t n, uint32_t value, uint32_t* error, cudaStream_t stream) {
    if (n > (uint32_t)kMaxPeers) return cudaErrorInvalidValue;
    FlagList fl;
    for (int i = 0; i < kMaxPeers; i++) fl.f[i] = i < (int)n ? flags[i] : nullptr;
    k_wait_flags<<<1, 1, 0, stream>>>(fl, n, value, error);
    return cudaGetLastError();
}

cudaError_t launch_test_rng(const uint32_t* v0, const uint32_t* v1, uint64_t n, uint32_t n_draws, uint32_t* seeds, uint32_t* lcg_out,
                            float* rnd_out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_test_rng<<<grid_for(n, 128), 128, 0, stream>>>(v0, v1, n, n_draws, seeds, lcg_out, rnd_out);
    return cudaGetLastError();
}

cudaError_t launch_trace_rays(const RenderLaunch& scene, const float* o, const float* d, uint64_t n, float* t_out, int32_t* prim_out,
                              const uint32_t* orig, bool grid, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (grid) k_trace_rays_grid<<<grid_for(n, 128), 128, 0, stream>>>(scene.grid, scene.grid_start, scene.grid_refs, scene.geom, o, d, n, t_out, prim_out, orig, scene.gate != 0u);
    else k_trace_rays<<<grid_for(n, 128), 128, 0, stream>>>(scene.nodes, scene.geom, scene.root_link, o, d, n, t_out, prim_out, orig, scene.gate != 0u);
    return cudaGetLastError();
}

cudaError_t launch_make_color(const float* rgb, uint64_t n, uint32_t* out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_make_color<<<grid_for(n, 128), 128, 0, stream>>>(rgb, n, out);
    return cudaGetLastError();
}

cudaError_t launch_scatter(uint32_t type, float4 mat, const float* dirs, const float* normals, const uint8_t* front, const uint32_t* seeds,
                           uint64_t n, float* dirs_out, uint8_t* scattered, uint32_t* seeds_out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    k_scatter<<<grid_for(n, 128), 128, 0, stream>>>(type, mat, dirs, normals, front, seeds, n, dirs_out, scattered, seeds_out);
    return cudaGetLastError();
}

}  // namespace VN_NS
}  // namespace vn
