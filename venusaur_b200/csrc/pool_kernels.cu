// pool_kernels.cu -- the warp-pool wavefront path kernel: the wavefront organisation of the north star (SoA ray/hit
// queues, ballot/popc compaction, material-sorted shade queues, persistent threads) held ENTIRELY IN SHARED MEMORY.
// Compiled twice like path_kernels.cu (vn::exact / vn::fast).
//
// Why: ncu on k_render_persistent (profiles/) shows the SM issue slots 82 % busy with only ~10 of 32 lanes active per
// instruction -- a warp's traversal lasts as long as its slowest ray, and sphere tests / shading run for a few lanes at
// a time.  HBM queues (wavefront.cu) fix the lane utilisation but pay ~150 B of queue traffic and grid barriers per
// bounce.  Here the queues live next to the scene in the SM's 227 KB:
//
//   * every WARP owns a pool of S path slots (S = 64..160) in shared memory, SoA, one 4-byte field array per state word
//     (origin, direction, throughput, path seed, camera seed, pixel sum, pixel id, depth/samples-left, hit t, hit prim:
//     18 words = 72 B per slot).  A slot is a pixel "in flight": it walks the pixel's spp samples in order, so the
//     reference's per-pixel RNG chain (RayTracer.cu:169-183) and float summation order (:203) are preserved exactly.
//   * the warp cycles through three phases, synchronised only by __syncwarp/ballots (no block or grid barrier):
//       regen   slots whose path ended: add the radiance to the pixel sum, finish the pixel / take the next pixel from
//               the global ticket when its samples are done, generate the next camera ray      -> ray queue
//       extend  lanes PULL rays from the ray queue as they become free (persistent lanes), so a long traversal no
//               longer stalls 31 finished lanes; node steps and leaf (sphere) steps are issued for the sub-set of lanes
//               that want them, and leaf steps are batched until enough lanes wait           -> hit queues by material
//       shade   one compacted queue per shading program: opaque (Lambertian + metal share their rejection loop) and
//               dielectric; survivors -> ray queue, ended paths -> regen queue.  Misses go straight to regen.
//     Queue appends are warp ballots + popc prefix ranks; queue sizes are warp-uniform registers: no atomics.
//   * the math is the same vn_math.cuh code as everywhere else: in the IEEE build the accumulation buffer is
//     bit-identical to the persistent kernel's and to the oracle's (tests/test_gpu_parity.py).
#include <algorithm>

#include "kernels.h"

#if VN_EXACT
#define VN_NS exact
#else
#define VN_NS fast
#endif

namespace vn {
namespace VN_NS {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kPoolFields = 18;
constexpr uint32_t kMetaHasPixel = 0x80000000u;   // meta = has_pixel<<31 | miss_pending<<30 | s_left<<15 | depth
constexpr uint32_t kMetaMiss = 0x40000000u;

enum LaneState : int { kIdle = 0, kBusy = 1, kDoneRay = 2 };

struct Pool {
    float *ox, *oy, *oz, *dx, *dy, *dz, *tr, *tg, *tb, *sr, *sg, *sb, *hit_t;
    uint32_t *seed, *cam_seed, *pix, *meta;
    int32_t* hit_prim;
    uint16_t *q_ray, *q_opaque, *q_diel, *q_regen;
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() { return (1u << lane_id()) - 1u; }

// Appends `slot` of every lane with flag set to queue q (warp-uniform size n).  All 32 lanes must call.
__device__ __forceinline__ void q_append(uint16_t* q, uint32_t& n, bool flag, uint32_t slot) {
    const unsigned m = __ballot_sync(kFull, flag);
    if (flag) q[n + __popc(m & lanemask_lt())] = (uint16_t)slot;
    n += __popc(m);
}

__device__ __forceinline__ void pool_finish_pixel(const RenderLaunch& p, uint32_t pix, f3 sum) {
    f3 mean = sum * p.inv_spp;                                   // RayTracer.cu:206
    if (p.blend_mode == kBlendLerp) {                            // RayTracer.cu:208-213
        const float4 prev = p.accum[pix];
        mean = lerp3(mk3(prev.x, prev.y, prev.z), mean, p.blend_a);
    } else if (p.blend_mode == kBlendSum) {
        const float4 prev = p.accum[pix];
        mean = mk3(prev.x, prev.y, prev.z) + mean;
    }
    p.accum[pix] = make_float4(mean.x, mean.y, mean.z, 1.0f);    // RayTracer.cu:215; make_color (:216) runs in k_tonemap
}

template <bool kSmem>
__global__ void __launch_bounds__(768) k_render_pool(const __grid_constant__ RenderLaunch p, const uint32_t S, const uint32_t scene_bytes,
                                                     const uint32_t service_threshold, const uint32_t leaf_batch) {
    extern __shared__ float4 s_mem[];
    SceneView sc;
    if (kSmem) {
        float4* s_nodes = s_mem;
        float4* s_geom = s_nodes + 2 * (size_t)p.num_nodes;
        float4* s_mat = s_geom + p.num_spheres;
        uint8_t* s_type = reinterpret_cast<uint8_t*>(s_mat + p.num_spheres);
        for (uint32_t i = threadIdx.x; i < 2 * p.num_nodes; i += blockDim.x) s_nodes[i] = p.nodes[i];
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) { s_geom[i] = p.geom[i]; s_mat[i] = p.mat[i]; }
        for (uint32_t i = threadIdx.x; i < p.num_spheres; i += blockDim.x) s_type[i] = p.type[i];
        __syncthreads();
        sc.nodes = s_nodes; sc.geom = s_geom; sc.mat = s_mat; sc.type = s_type;
    } else {
        sc.nodes = p.nodes; sc.geom = p.geom; sc.mat = p.mat; sc.type = p.type;
    }
    sc.root_link = p.root_link;

    // ---- this warp's pool
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t pool_bytes = S * (kPoolFields * 4u + 4u * 2u);
    uint32_t* base = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(s_mem) + scene_bytes + (size_t)warp * pool_bytes);
    Pool P;
    {
        uint32_t* w = base;
        auto takef = [&]() { float* r = reinterpret_cast<float*>(w); w += S; return r; };
        auto takeu = [&]() { uint32_t* r = w; w += S; return r; };
        P.ox = takef(); P.oy = takef(); P.oz = takef(); P.dx = takef(); P.dy = takef(); P.dz = takef();
        P.tr = takef(); P.tg = takef(); P.tb = takef(); P.sr = takef(); P.sg = takef(); P.sb = takef(); P.hit_t = takef();
        P.seed = takeu(); P.cam_seed = takeu(); P.pix = takeu(); P.meta = takeu();
        P.hit_prim = reinterpret_cast<int32_t*>(takeu());
        uint16_t* q = reinterpret_cast<uint16_t*>(w);
        P.q_ray = q; P.q_opaque = q + S; P.q_diel = q + 2 * S; P.q_regen = q + 3 * S;
    }
    uint32_t n_ray = 0, n_opaque = 0, n_diel = 0, n_regen = 0;      // warp-uniform queue sizes
    for (uint32_t s = lane; s < S; s += 32) { P.meta[s] = 0u; P.q_regen[s] = (uint16_t)s; }
    n_regen = S;
    __syncwarp();

    uint32_t n_seg = 0, n_path = 0;
    bool pixels_left = true;                                           // warp-uniform: the global ticket still has work

    for (;;) {
        // ================= regen: ended paths -> pixel sum -> next sample / next pixel -> camera ray =================
        for (uint32_t c0 = 0; c0 < n_regen; c0 += 32) {
            const uint32_t e = c0 + lane;
            const bool valid = e < n_regen;
            const uint32_t slot = valid ? P.q_regen[e] : 0u;
            uint32_t meta = valid ? P.meta[slot] : 0u;
            bool has_pixel = valid && (meta & kMetaHasPixel);
            uint32_t s_left = (meta >> 15) & 0x7FFFu;
            f3 sum = mk3(0.0f);
            uint32_t pix = 0, cam_seed = 0;
            if (has_pixel) {
                sum = mk3(P.sr[slot], P.sg[slot], P.sb[slot]);
                pix = P.pix[slot];
                cam_seed = P.cam_seed[slot];
                f3 result = mk3(0.0f);                                 // absorbed / depth exhausted: black
                if (meta & kMetaMiss) {                                // __miss__ms, RayTracer.cu:442-450
                    const f3 d = mk3(P.dx[slot], P.dy[slot], P.dz[slot]);
                    result = mk3(P.tr[slot], P.tg[slot], P.tb[slot]) * sky(normalize(d));
                }
                sum = sum + result;                                    // pixel_color += prd.attenuation, RayTracer.cu:203
                if (s_left == 0u) { pool_finish_pixel(p, pix, sum); has_pixel = false; }
            }
            // slots without a pixel take the next one from the global ticket (one atomic per warp and round); a ticket
            // that maps outside the frame (ragged 8x4 tiles) is dropped and the slot asks again
            bool want_pixel = valid && !has_pixel;
            while (pixels_left) {
                const unsigned want_mask = __ballot_sync(kFull, want_pixel);
                if (want_mask == 0u) break;
                uint32_t first = 0;
                if (lane == 0) first = atomicAdd(p.work_counter, (uint32_t)__popc(want_mask));
                first = __shfl_sync(kFull, first, 0);
                if (first + (uint32_t)__popc(want_mask) >= p.total_work) pixels_left = false;
                if (want_pixel) {
                    const uint32_t w = first + __popc(want_mask & lanemask_lt());
                    if (w < p.total_work) {
                        const uint32_t tile = w >> 5, in = w & 31u;
                        const uint32_t ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
                        const uint32_t px = tx * 8u + (in & 7u), py = p.row_begin + ty * 4u + (in >> 3);
                        if (px < p.width && py < p.row_end) {
                            pix = py * p.width + px;
                            cam_seed = tea4(pix, p.subframe_index);    // RayTracer.cu:169
                            sum = mk3(0.0f);
                            s_left = p.spp;
                            has_pixel = true;
                            want_pixel = false;
                        }
                    }
                }
            }
            if (has_pixel) {
                const uint32_t py = pix / p.width, px = pix - py * p.width;
                f3 o, d;
                camera_ray(p.cam, px, py, cam_seed, o, d);             // RayTracer.cu:173-177
                s_left -= 1u;
                P.ox[slot] = o.x; P.oy[slot] = o.y; P.oz[slot] = o.z;
                P.dx[slot] = d.x; P.dy[slot] = d.y; P.dz[slot] = d.z;
                P.tr[slot] = 1.0f; P.tg[slot] = 1.0f; P.tb[slot] = 1.0f;
                P.seed[slot] = cam_seed;                               // prd.seed = seed: a copy (:183)
                P.cam_seed[slot] = cam_seed;
                P.sr[slot] = sum.x; P.sg[slot] = sum.y; P.sb[slot] = sum.z;
                P.pix[slot] = pix;
                P.meta[slot] = kMetaHasPixel | (s_left << 15) | (p.max_depth - 1u);   // depth = max_depth - 1 (:184)
                n_path += 1u;
            } else if (valid) {
                P.meta[slot] = 0u;                                     // the ticket is exhausted: this slot retires
            }
            q_append(P.q_ray, n_ray, has_pixel, slot);
        }
        n_regen = 0;
        __syncwarp();
        if (n_ray == 0u) break;                                        // no live paths and no pixels left

        // ================= extend: persistent lanes pull rays; node / leaf steps issued per sub-set ===================
        {
            uint32_t q_head = 0;
            int state = kIdle;
            uint32_t slot = 0, cur = kEmptyScene;
            int sp = 0, prim = -1;
            float tbest = kTMax, a = 1.0f, inv_a = 1.0f;
            f3 o = mk3(0.0f), d = mk3(0.0f), idir = mk3(0.0f), ood = mk3(0.0f);
            uint32_t stack[kStackSize];
            for (;;) {
                unsigned busy_mask = __ballot_sync(kFull, state == kBusy);
                const bool queue_left = q_head < n_ray;
                if (busy_mask == 0u || (queue_left && (32u - __popc(busy_mask)) >= service_threshold)) {
                    // ---- service: retire finished rays into the hit queues, hand out new rays
                    const bool done = state == kDoneRay;
                    int cls = -1;                                      // 0 miss, 1 opaque, 2 dielectric
                    if (done) {
                        P.hit_t[slot] = tbest;
                        P.hit_prim[slot] = prim;
                        cls = prim < 0 ? 0 : (sc.type[prim] == 2u ? 2 : 1);
                        if (prim < 0) P.meta[slot] |= kMetaMiss;
                        n_seg += 1u;
                    }
                    q_append(P.q_regen, n_regen, cls == 0, slot);
                    q_append(P.q_opaque, n_opaque, cls == 1, slot);
                    q_append(P.q_diel, n_diel, cls == 2, slot);
                    if (done) state = kIdle;
                    const bool idle = state == kIdle;
                    const unsigned want = __ballot_sync(kFull, idle);
                    const uint32_t idx = q_head + __popc(want & lanemask_lt());
                    if (idle && idx < n_ray) {
                        slot = P.q_ray[idx];
                        o = mk3(P.ox[slot], P.oy[slot], P.oz[slot]);
                        d = mk3(P.dx[slot], P.dy[slot], P.dz[slot]);
                        idir = slab_idir(d);
                        ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
                        a = dot(d, d);
                        inv_a = rcp(a);
                        tbest = kTMax;
                        prim = -1;
                        sp = 0;
                        cur = sc.root_link;
                        state = cur == kEmptyScene ? kDoneRay : kBusy;
                    }
                    q_head = min(n_ray, q_head + (uint32_t)__popc(want));
                    busy_mask = __ballot_sync(kFull, state == kBusy);
                    if (busy_mask == 0u) {
                        if (__ballot_sync(kFull, state == kDoneRay) == 0u && q_head >= n_ray) break;
                        continue;
                    }
                }
                // ---- one traversal step for the lanes that want it
                const bool at_leaf = state == kBusy && (cur & kLeafFlag);
                const bool at_node = state == kBusy && !(cur & kLeafFlag);
                const unsigned leaf_mask = __ballot_sync(kFull, at_leaf);
                const unsigned node_mask = busy_mask & ~leaf_mask;
                if (node_mask != 0u && (uint32_t)__popc(leaf_mask) < leaf_batch) {
                    if (at_node) {
                        const float4 l0 = sc.nodes[2 * cur], l1 = sc.nodes[2 * cur + 1];
                        const float4 r0 = sc.nodes[2 * cur + 2], r1 = sc.nodes[2 * cur + 3];
                        float tl, tr;
                        const bool hl = box_hit(l0, l1, idir, ood, tbest, tl);
                        const bool hr = box_hit(r0, r1, idir, ood, tbest, tr);
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
                        if (cur == kEmptyScene) state = kDoneRay;
                    }
                } else {
                    if (at_leaf) {
                        const uint32_t first = (cur & 0x7FFFFFFFu) >> 3;
                        const uint32_t count = (cur & 7u) + 1u;
                        for (uint32_t k = 0; k < count; k++) {
                            const float4 g = sc.geom[first + k];
                            const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                            if (t >= 0.0f) { tbest = t; prim = (int)(first + k); }
                        }
                        cur = sp ? stack[--sp] : kEmptyScene;
                        if (cur == kEmptyScene) state = kDoneRay;
                    }
                }
            }
        }
        n_ray = 0;
        __syncwarp();

        // ================= shade: one compacted queue per shading program ===========================================
#pragma unroll 1
        for (int which = 0; which < 2; which++) {
            const uint16_t* q = which == 0 ? P.q_opaque : P.q_diel;
            const uint32_t n = which == 0 ? n_opaque : n_diel;
            for (uint32_t c0 = 0; c0 < n; c0 += 32) {
                const uint32_t e = c0 + lane;
                const bool valid = e < n;
                const uint32_t slot = valid ? q[e] : 0u;
                bool cont = false;
                if (valid) {
                    PathState st;
                    st.o = mk3(P.ox[slot], P.oy[slot], P.oz[slot]);
                    st.d = mk3(P.dx[slot], P.dy[slot], P.dz[slot]);
                    st.thr = mk3(P.tr[slot], P.tg[slot], P.tb[slot]);
                    st.seed = P.seed[slot];
                    const uint32_t meta = P.meta[slot];
                    st.depth = (int)(meta & 0x7FFFu);
                    f3 result;
                    cont = shade_segment(sc, st, P.hit_t[slot], P.hit_prim[slot], result);
                    if (cont) {
                        P.ox[slot] = st.o.x; P.oy[slot] = st.o.y; P.oz[slot] = st.o.z;
                        P.dx[slot] = st.d.x; P.dy[slot] = st.d.y; P.dz[slot] = st.d.z;
                        P.tr[slot] = st.thr.x; P.tg[slot] = st.thr.y; P.tb[slot] = st.thr.z;
                        P.seed[slot] = st.seed;
                        P.meta[slot] = (meta & ~0x7FFFu) | (uint32_t)st.depth;
                    }
                    // ended here (absorbed or depth budget exhausted): radiance 0, kMetaMiss stays clear
                }
                q_append(P.q_ray, n_ray, valid && cont, slot);
                q_append(P.q_regen, n_regen, valid && !cont, slot);
            }
        }
        n_opaque = 0;
        n_diel = 0;
        __syncwarp();
    }

    {
        unsigned long long seg = n_seg, path = n_path;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            seg += __shfl_xor_sync(kFull, seg, o);
            path += __shfl_xor_sync(kFull, path, o);
        }
        if (lane == 0) { atomicAdd(&p.counters[0], seg); atomicAdd(&p.counters[1], path); }
    }
}

}  // namespace

size_t pool_smem_bytes(uint32_t num_nodes, uint32_t num_spheres, bool scene_in_smem, int warps, uint32_t slots) {
    const size_t scene = scene_in_smem ? ((scene_smem_bytes(num_nodes, num_spheres) + 15) & ~(size_t)15) : 0;
    return scene + (size_t)warps * slots * (kPoolFields * 4u + 4u * 2u);
}

cudaError_t launch_render_pool(const RenderLaunch& p, bool scene_in_smem, int threads, int blocks, uint32_t slots, uint32_t service_threshold,
                               uint32_t leaf_batch, cudaStream_t stream) {
    const uint32_t scene_bytes = scene_in_smem ? (uint32_t)((scene_smem_bytes(p.num_nodes, p.num_spheres) + 15) & ~(size_t)15) : 0u;
    const size_t smem = pool_smem_bytes(p.num_nodes, p.num_spheres, scene_in_smem, threads / 32, slots);
    cudaError_t e;
    if (scene_in_smem) {
        e = cudaFuncSetAttribute(k_render_pool<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_render_pool<true><<<blocks, threads, smem, stream>>>(p, slots, scene_bytes, service_threshold, leaf_batch);
    } else {
        e = cudaFuncSetAttribute(k_render_pool<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_render_pool<false><<<blocks, threads, smem, stream>>>(p, slots, scene_bytes, service_threshold, leaf_batch);
    }
    return cudaGetLastError();
}

int pool_max_blocks_per_sm(bool scene_in_smem, int threads, size_t smem) {
    int nb = 0;
    cudaError_t e;
    if (scene_in_smem) {
        cudaFuncSetAttribute(k_render_pool<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_render_pool<true>, threads, smem);
    } else {
        cudaFuncSetAttribute(k_render_pool<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_render_pool<false>, threads, smem);
    }
    return e == cudaSuccess ? nb : -1;
}

}  // namespace VN_NS
}  // namespace vn
