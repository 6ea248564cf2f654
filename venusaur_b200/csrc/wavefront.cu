// wavefront.cu -- the queue-based wavefront path tracer (north-star kernel 2) and its accumulate/tonemap kernel
// (north-star kernel 3).  Compiled twice like path_kernels.cu (vn::exact / vn::fast).
//
// One PERSISTENT cooperative kernel (grid = SMs x occupancy, resident for the whole frame) runs the bounce loop;
// the phases of one iteration are separated by grid-wide barriers instead of kernel launches:
//   generate   free slots of the ray queue are refilled with new camera paths (path regeneration), one thread per
//              pixel walking its spp samples in order so the per-pixel RNG chain is the reference's
//              (RayTracer.cu:169-183); written sample-major so the stores coalesce.
//   extend     closest hit for every ray of the queue (RayTracer.cu:190-202, 229-270) -> hit queue {t, prim};
//              blocks take 256-ray chunks from an atomic ticket.  Each ray is then binned by what it hit into four
//              material queues (miss / Lambertian / metal / dielectric): per warp a ballot + popc gives the lane's
//              rank, a block-level prefix over the 8 warps gives the block's total, ONE atomicAdd per block and
//              material reserves the span.
//   shade      one material queue at a time, so a warp runs a single closest-hit program (RayTracer.cu:272-450)
//              without divergence; finished paths store their radiance in the per-(sample, pixel) buffer, survivors
//              are compacted into the other ray queue (same ballot/popc + block prefix + one atomic per block).
// Ray and hit queues are SoA arrays of 4-byte fields: every load/store of a warp is one 128-byte line.  The scene
// (32-byte nodes read as 2 x 128-bit, spheres, materials) is staged in shared memory when it fits.
// k_wf_accumulate then sums each pixel's samples IN SAMPLE ORDER (bit-identical to the reference's
// pixel_color += ..., RayTracer.cu:203), blends into accum and writes uchar4 with coalesced stores (:206-216).
#include <cooperative_groups.h>

#include <algorithm>

#include "kernels.h"

namespace cg = cooperative_groups;

#if VN_EXACT
#define VN_NS exact
#else
#define VN_NS fast
#endif

namespace vn {
namespace VN_NS {

namespace {

constexpr int kWfThreads = 256;      // CTA size of the pair-node variants; the wide-node variant (one CTA per SM, scene copies fill its shared memory) runs 1024

__device__ __forceinline__ uint32_t ld_count(const uint32_t* p) { return __ldcg(p); }

// Block-wide stream compaction step: returns this thread's output position for a `flag`ged item, after reserving
// the block's span with one atomicAdd on *global_count.  All threads of the block must call it.
template <int kWfWarps>
__device__ __forceinline__ uint32_t block_reserve(bool flag, uint32_t* global_count, uint32_t* s_warp /*[kWfWarps]*/, uint32_t* s_base) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t mask = __ballot_sync(0xffffffffu, flag);
    const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kWfWarps; w++) { const uint32_t c = s_warp[w]; s_warp[w] = total; total += c; }
        *s_base = total ? atomicAdd(global_count, total) : 0u;
    }
    __syncthreads();
    const uint32_t pos = *s_base + s_warp[warp] + rank;
    __syncthreads();
    return pos;
}

// kWide: the extend phase traverses the 4-wide, octant-sorted nodes of the path kernels (eight copies in shared memory, lbvh_core.cuh::
// wide_octant_node) after the huge-sphere list -- the same closest-hit structure as k_render_lean, so that wavefront against megakernel is a
// comparison of SCHEDULES (round 1 compared this kernel on pair nodes with a megakernel on wide nodes).
template <bool kSmem, bool kWide, int kThreads>
__global__ void __launch_bounds__(kThreads) k_wavefront(const __grid_constant__ RenderLaunch p, const __grid_constant__ WavefrontBuffers wf) {
    constexpr int kWfWarps = kThreads / 32;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float4 s_scene[];
    __shared__ uint32_t s_warp[kWfWarps];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_ticket;

    SceneView sc;
    const uint32_t node_f4s = kWideNodeF4 * p.num_wide;
    if (kWide) {
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
    } else if (kSmem) {
        float4* s_nodes = s_scene;
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

    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gsize = gridDim.x * blockDim.x;
    const uint32_t rows = p.row_end - p.row_begin;
    const uint32_t region_pixels = p.width * rows;
    uint32_t* C = wf.counts;

    uint32_t cur = 0;          // ray queue being traced
    uint32_t live = 0;         // its size          } uniform over the grid: every thread derives them from the
    uint32_t next_pixel = 0;   // pixels started     } same counters after the same barriers
    uint32_t n_seg = 0, n_path = 0;

    for (;;) {
        // ---------------- generate: refill the current queue with new camera paths
        const uint32_t room = (wf.capacity - live) / p.spp;
        const uint32_t n_new = min(room, region_pixels - next_pixel);
        {
            const WfState A = wf.st[cur];
            for (uint32_t j = gtid; j < n_new; j += gsize) {
                const uint32_t local = next_pixel + j;
                const uint32_t ly = local / p.width, px = local - ly * p.width, py = p.row_begin + ly;
                uint32_t cam_seed = tea4(py * p.width + px, p.subframe_index);      // RayTracer.cu:169
                for (uint32_t s = 0; s < p.spp; s++) {
                    f3 o, d;
                    camera_ray(p.cam, px, py, cam_seed, o, d);                      // RayTracer.cu:173-177
                    const uint32_t slot = live + s * n_new + j;
                    A.ox[slot] = o.x; A.oy[slot] = o.y; A.oz[slot] = o.z;
                    A.dx[slot] = d.x; A.dy[slot] = d.y; A.dz[slot] = d.z;
                    A.tr[slot] = 1.0f; A.tg[slot] = 1.0f; A.tb[slot] = 1.0f;
                    A.seed[slot] = cam_seed;                                        // prd.seed = seed, a copy (:183)
                    A.ps[slot] = s * region_pixels + local;
                    A.depth[slot] = (int32_t)p.max_depth - 1;                       // :184
                    n_path += 1u;
                }
            }
        }
        live += n_new * p.spp;
        next_pixel += n_new;
        if (live == 0u) break;
        if (gtid == 0) {
#pragma unroll
            for (int k = 0; k < (int)kWfCountWords; k++) C[k] = 0u;
        }
        grid.sync();

        // ---------------- extend: closest hit + binning into material queues
        {
            const WfState A = wf.st[cur];
            for (;;) {
                if (threadIdx.x == 0) s_ticket = atomicAdd(&C[kWfTicketExtend], (uint32_t)kThreads);
                __syncthreads();
                const uint32_t base = s_ticket;
                __syncthreads();
                if (base >= live) break;
                const uint32_t i = base + threadIdx.x;
                const bool valid = i < live;
                uint32_t cls = 0xFFu;
                if (valid) {
                    float t;
                    int prim;
                    TraceCounters cnt{0u, 0u};
                    const f3 ro = mk3(A.ox[i], A.oy[i], A.oz[i]), rd = mk3(A.dx[i], A.dy[i], A.dz[i]);
                    if (kWide) {
                        float t0 = kTMax;
                        int prim0 = -1;
                        if (p.huge.n) {                           // the huge spheres first (lbvh_core.cuh::HugeList), as in the path kernels
                            const float a = dot(rd, rd), inv_a = rcp(a);
                            for (uint32_t h = 0; h < p.huge.n; h++) {
                                const uint32_t hs = p.huge.idx[h];
                                const float4 g = sc.geom[hs];
                                const float th = sphere_root(ro, rd, a, inv_a, g.x, g.y, g.z, g.w, kTMin, t0);
                                if (th >= 0.0f) { t0 = th; prim0 = (int)hs; }
                            }
                        }
                        closest_hit_wide<false>(sc.nodes, node_f4s, sc.geom, p.wide_root, ro, rd, t, prim, cnt, t0, prim0);
                    } else {
                        closest_hit<false>(sc.nodes, sc.geom, sc.root_link, ro, rd, t, prim, cnt, 0u, p.gate != 0u);
                    }
                    wf.hit_t[i] = t;
                    wf.hit_prim[i] = prim;
                    cls = prim < 0 ? 0u : 1u + (uint32_t)sc.type[prim];
                    n_seg += 1u;
                }
#pragma unroll
                for (uint32_t m = 0; m < 4u; m++) {
                    const bool mine = cls == m;
                    const uint32_t pos = block_reserve<kWfWarps>(mine, &C[kWfCountMat + m], s_warp, &s_base);
                    if (mine) wf.mat_queue[m][pos] = i;
                }
            }
        }
        grid.sync();

        // ---------------- shade: one material queue at a time; survivors compacted into the other ray queue
        {
            const WfState A = wf.st[cur];
            const WfState B = wf.st[cur ^ 1u];
#pragma unroll 1
            for (uint32_t m = 0; m < 4u; m++) {
                const uint32_t n_m = ld_count(&C[kWfCountMat + m]);
                for (;;) {
                    if (threadIdx.x == 0) s_ticket = atomicAdd(&C[kWfTicketShade + m], (uint32_t)kThreads);
                    __syncthreads();
                    const uint32_t base = s_ticket;
                    __syncthreads();
                    if (base >= n_m) break;
                    const uint32_t q = base + threadIdx.x;
                    bool cont = false;
                    PathState st;
                    uint32_t ps = 0;
                    if (q < n_m) {
                        const uint32_t i = wf.mat_queue[m][q];
                        st.o = mk3(A.ox[i], A.oy[i], A.oz[i]);
                        st.d = mk3(A.dx[i], A.dy[i], A.dz[i]);
                        st.thr = mk3(A.tr[i], A.tg[i], A.tb[i]);
                        st.seed = A.seed[i];
                        st.depth = A.depth[i];
                        ps = A.ps[i];
                        f3 result;
                        cont = shade_segment(sc, st, wf.hit_t[i], wf.hit_prim[i], result);
                        if (!cont) {
                            float* out = wf.sample_rgb + 3ull * ps;
                            out[0] = result.x; out[1] = result.y; out[2] = result.z;
                        }
                    }
                    const uint32_t pos = block_reserve<kWfWarps>(cont, &C[kWfCountNext], s_warp, &s_base);
                    if (cont) {
                        B.ox[pos] = st.o.x; B.oy[pos] = st.o.y; B.oz[pos] = st.o.z;
                        B.dx[pos] = st.d.x; B.dy[pos] = st.d.y; B.dz[pos] = st.d.z;
                        B.tr[pos] = st.thr.x; B.tg[pos] = st.thr.y; B.tb[pos] = st.thr.z;
                        B.seed[pos] = st.seed;
                        B.ps[pos] = ps;
                        B.depth[pos] = st.depth;
                    }
                }
            }
        }
        grid.sync();
        live = ld_count(&C[kWfCountNext]);
        cur ^= 1u;
        grid.sync();   // nobody may zero the counters (next iteration) before everyone has read `live`
    }

    {
        unsigned long long seg = n_seg, path = n_path;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            seg += __shfl_xor_sync(0xffffffffu, seg, o);
            path += __shfl_xor_sync(0xffffffffu, path, o);
        }
        if ((threadIdx.x & 31u) == 0u) { atomicAdd(&p.counters[0], seg); atomicAdd(&p.counters[1], path); }
    }
}

// North-star kernel (3): per pixel, sum the spp sample radiances in sample order, mean, blend into accum, sRGB +
// quantise.  One pixel per thread: coalesced 12-byte reads per sample plane, 16-byte accum RMW, 4-byte uchar4 store.
__global__ void __launch_bounds__(256) k_wf_accumulate(const __grid_constant__ RenderLaunch p, const float* __restrict__ sample_rgb) {
    const uint32_t rows = p.row_end - p.row_begin;
    const uint32_t region_pixels = p.width * rows;
    for (uint32_t local = blockIdx.x * blockDim.x + threadIdx.x; local < region_pixels; local += gridDim.x * blockDim.x) {
        f3 sum = mk3(0.0f);
        for (uint32_t s = 0; s < p.spp; s++) {
            const float* in = sample_rgb + 3ull * ((uint64_t)s * region_pixels + local);
            sum = sum + mk3(in[0], in[1], in[2]);                          // RayTracer.cu:203
        }
        const uint32_t pix = p.row_begin * p.width + local;
        f3 mean = sum * p.inv_spp;                                         // RayTracer.cu:206
        if (p.blend_mode == kBlendLerp) {                                  // RayTracer.cu:208-213
            const float4 prev = p.accum[pix];
            mean = lerp3(mk3(prev.x, prev.y, prev.z), mean, p.blend_a);
        } else if (p.blend_mode == kBlendSum) {
            const float4 prev = p.accum[pix];
            mean = mk3(prev.x, prev.y, prev.z) + mean;
        }
        p.accum[pix] = make_float4(mean.x, mean.y, mean.z, 1.0f);          // RayTracer.cu:215
        if (p.image) p.image[pix] = make_color_u32(mean);                  // RayTracer.cu:216
    }
}

template <bool kSmem, bool kWide, int kThreads>
cudaError_t launch_wf(const RenderLaunch& p, const WavefrontBuffers& wf, int num_sms, size_t smem, cudaStream_t stream) {
    cudaError_t e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(k_wavefront<kSmem, kWide, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wavefront<kSmem, kWide, kThreads>, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    dim3 grid((unsigned)(num_sms * per_sm)), block(kThreads);
    void* args[] = {(void*)&p, (void*)&wf};
    return cudaLaunchCooperativeKernel((const void*)k_wavefront<kSmem, kWide, kThreads>, grid, block, args, smem, stream);
}

}  // namespace

cudaError_t launch_wavefront(const RenderLaunch& p, const WavefrontBuffers& wf, int num_sms, cudaStream_t stream, uint32_t* launches) {
    if (wf.capacity < p.spp) return cudaErrorInvalidValue;
    const size_t need = scene_smem_bytes(p.num_nodes, p.num_spheres);
    const bool in_smem = p.num_spheres > 0 && need <= 100 * 1024;
    // the wide nodes of the path kernels when their eight octant copies fit (the caller clears p.wide otherwise): one 1024-thread CTA per SM
    const size_t need_wide = p.wide && p.num_wide ? wide_smem_bytes(p.num_wide, p.num_spheres) : 0;
    cudaError_t e = need_wide ? launch_wf<true, true, 1024>(p, wf, num_sms, need_wide, stream)
                              : (in_smem ? launch_wf<true, false, kWfThreads>(p, wf, num_sms, need, stream) : launch_wf<false, false, kWfThreads>(p, wf, num_sms, 0, stream));
    if (e != cudaSuccess) return e;
    const uint32_t region_pixels = p.width * (p.row_end - p.row_begin);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)region_pixels + 255) / 256, (uint64_t)num_sms * 16);
    k_wf_accumulate<<<blocks, 256, 0, stream>>>(p, wf.sample_rgb);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

// North-star kernel (3) on its own: used behind the slot-scheduled path kernel, which also writes per-sample radiance.
cudaError_t launch_accumulate_samples(const RenderLaunch& p, const float* sample_rgb, cudaStream_t stream) {
    const uint32_t region_pixels = p.width * (p.row_end - p.row_begin);
    if (region_pixels == 0u) return cudaSuccess;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(((uint64_t)region_pixels + 255) / 256, 148ull * 16);
    k_wf_accumulate<<<blocks, 256, 0, stream>>>(p, sample_rgb);
    return cudaGetLastError();
}

}  // namespace VN_NS
}  // namespace vn
