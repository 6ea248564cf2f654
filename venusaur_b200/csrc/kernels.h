// kernels.h -- internal launch interface between the C ABI (vn_api.cu) and the trace / tonemap kernels.
// path_kernels.cu and wavefront.cu are each compiled twice: namespace vn::exact (-DVN_EXACT=1 -fmad=false) and
// namespace vn::fast (-DVN_EXACT=0, FMA + approximate reciprocal/rsqrt).  See vn_math.cuh.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vn_math.cuh"
#include "grid_core.cuh"
#include "lbvh_core.cuh"

namespace vn {

constexpr int kMaxPeers = 8;

enum BlendMode : uint32_t { kBlendOverwrite = 0, kBlendLerp = 1, kBlendSum = 2 };

// Everything one render launch needs; passed by value (__grid_constant__), i.e. in the constant bank like the
// reference's `__constant__ Params params` (RayTracer.cu:49-52).
struct RenderLaunch {
    Camera cam;
    uint32_t width, height, spp, subframe_index, max_depth;
    uint32_t row_begin, row_end;
    uint32_t blend_mode;
    uint32_t n_sub, tiles_per_sub, accum_count, sub_stride;   // k_render_lean<kMulti>: subframes subframe_index + j * sub_stride, j < n_sub, in one launch (tickets = j * tiles_per_sub + tile rank);
                                                  //   accum_count = subframes the running mean holds before the first of them (path_kernels.cu::finish_pixel_multi); n_sub = 1 otherwise
    float blend_a, inv_spp;
    float4* accum;
    uint32_t* image;                 // uchar4 pixels, may be null
    const float4* nodes;
    const float4* geom;
    const float4* mat;
    const uint8_t* type;
    uint32_t root_link, num_nodes, num_spheres;
    const float4* wide;              // canonical 4-wide nodes (8 float4 each) or null; staged per octant by the path kernel
    uint32_t num_wide, wide_root;
    HugeList huge;                   // spheres left out of the wide nodes: tested by every ray before the traversal (lbvh_core.cuh)
    GridHeader grid;                 // uniform grid + oversize list (kGrid kernels); grid_start/grid_refs are device arrays staged in shared memory
    const uint16_t* grid_start;
    const uint16_t* grid_refs;
    uint32_t grid_vote;              // grid traversal: lanes holding an untested sphere that trigger the sphere turn (0 = while-while)
    uint32_t leaf_vote;              // wide traversal: lanes waiting at a leaf that trigger the leaf turn (0 = while-while phases)
    uint32_t async_done, async_node, async_leaf;   // k_render_async: lanes with a finished ray that end a traversal burst; lanes that keep node steps / trigger a leaf turn
    unsigned long long* counters;    // [0] segments, [1] paths, [2] node visits, [3] sphere tests
    uint32_t* work_counter;          // persistent-thread work ticket
    uint32_t total_work, tiles_x;    // work items = 8x4 pixel tiles * 32
    const uint32_t* tile_order;      // ticket >> 5 -> tile index, most expensive tiles first (null = row-major); see vn_api.cu::prepare_tile_order
    uint32_t* tile_cost;             // when non-null the kernel records every finished pixel's ray segments in its tile's entries:
    uint32_t tile_cost_stride;       //   tile_cost[tile] += segments, tile_cost[tile_cost_stride + tile] = max(.., segments)
    uint32_t* timeline;              // instrumented launches of k_render_lean: [0..1024) lanes retired per 8 us bin since the CTA's start, [1024..2048) the ray
                                     // segments of the last pixel those lanes finished (0 outside the cost-collecting launch); null otherwise
    const uint4* qnodes;             // k_render_lean<kGlobal>: the quantised pairs (lbvh.cu::k_quantize_pairs) or null: one 256-bit load per traversal step
    float q_lo[3], q_scale[3];       //   plane = q * q_scale + q_lo (the root box and its extent / 65535)
    uint32_t units_log2;             // k_render_lean<kGlobal>: the work items are (tile, unit) with a tile's spp samples cut into 1 << units_log2 sample ranges; the item
                                     //   list in tile_order carries tile | unit << 24 (path_kernels.cu, "sample-range units"); 0 = whole tiles
    float4* carry;                   //   per pixel (tile-major: tile * 32 + position in the tile) 2 x float4, written by ONE 32-byte store: {sum so far, -} {camera seed,
                                     //   unit_epoch + finished units of this launch, -, -}: what a unit hands to the next, and the tag that says it is there
    uint32_t unit_epoch;             //   (a multiple of 32 that grows with every launch: no memset between launches)
    uint32_t steal;                  // k_render_lean: >= 1 = sample stealing inside a warp during the drain (path_kernels.cu): the fewest samples a lane must have left to give one away
    float4* steal_scratch;           //   one slot of spp + 1 float4 per lane of the grid: [0] the owner's prefix sum (.w = first sample of the suffix), [1 + k] sample k's radiance
    uint32_t* steal_count;           //   per slot: parts handed in (bit 31: the owner's prefix); zero between launches
    uint32_t gate;                   // 1: the pair-node / L2-HBM traversals apply the hit-point gate (vn_math.cuh::hit_gate_ok); the shared-memory wide-node kernels ignore it
    uint32_t tiles_x_inv;            // floor(2^32 / tiles_x): tile / tiles_x = __umulhi(tile, tiles_x_inv) plus at most two corrections (tile_row_col)
};

// tile index -> (row, column) of the 8x4-pixel work tile without the 25-instruction integer division (the quotient estimate is at most
// two short for any 32-bit tile index; exact for tiles_x = 1, where floor(2^32 / 1) does not fit and 0xFFFFFFFF is stored)
__device__ __forceinline__ void tile_row_col(uint32_t tile, uint32_t tiles_x, uint32_t tiles_x_inv, uint32_t& ty, uint32_t& tx) {
    uint32_t q = __umulhi(tile, tiles_x_inv);
    uint32_t r = tile - q * tiles_x;
    if (r >= tiles_x) { q += 1u; r -= tiles_x; }
    if (r >= tiles_x) { q += 1u; r -= tiles_x; }
    ty = q; tx = r;
}

struct KernelConfig {
    int threads;
    int blocks;
    size_t smem_bytes;               // dynamic shared memory (scene staging), 0 when the scene stays in HBM/L2
    bool scene_in_smem;
    bool count;                      // instrumented variant
    bool octant;                     // nodes staged 8x in shared memory, once per ray-direction octant (near/far-plane form)
    bool wide;                       // 4-wide octant-sorted nodes in shared memory (implies scene_in_smem; excludes octant)
    bool grid;                       // uniform grid + oversize list in shared memory (excludes the others)
    bool async = false;              // k_render_async (wide nodes in shared memory only): asynchronous shading, see path_kernels.cu
    bool warp_tiles = false;         // phase form only: warps own whole 8x4 tiles (one ticket per tile) instead of lanes taking single pixels
    int global_ctas = 4;             // k_render_lean<kGlobal>: CTAs of 256 threads per SM the kernel is compiled for (4 = 64 registers, 5 = 48, 6 = 40)
    bool lean = false;               // k_render_lean: the phase form with warp-owned tiles, 16-bit links and no per-lane statistics (path_kernels.cu)
};

size_t scene_smem_bytes(uint32_t num_nodes, uint32_t num_spheres, uint32_t node_copies = 1);
size_t wide_smem_bytes(uint32_t num_wide, uint32_t num_spheres);
size_t grid_smem_bytes(uint32_t n_cells, uint32_t n_refs, uint32_t num_spheres);

// Wavefront state: SoA queues in HBM (L2-resident at the default capacity), owned by the context.
// One path slot = 48 bytes of ray state (o 12, d 12, throughput 12, seed 4, sample slot 4, depth 4).
struct WfState {
    float *ox, *oy, *oz, *dx, *dy, *dz, *tr, *tg, *tb;
    uint32_t* seed;
    uint32_t* ps;        // index into sample_rgb: sample * region_pixels + region_pixel
    int32_t* depth;
};
constexpr int kWfStateArrays = 12;
enum WfCount : uint32_t {      // words of WavefrontBuffers::counts
    kWfCountNext = 0,          // entries appended to the next ray queue so far
    kWfCountMat = 1,           // [1..4] material queue sizes: miss, Lambertian, metal, dielectric
    kWfTicketExtend = 5,
    kWfTicketShade = 6,        // [6..9]
    kWfCountWords = 16
};
struct WavefrontBuffers {
    uint32_t capacity = 0;     // path slots per queue
    void* slab = nullptr;      // one allocation, carved into the arrays below
    WfState st[2];             // ray queues, ping-pong: shade compacts survivors of st[cur] into st[cur^1]
    float* hit_t = nullptr;    // hit queue, indexed like the current ray queue
    int32_t* hit_prim = nullptr;
    uint32_t* mat_queue[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* counts = nullptr;
    float* sample_rgb = nullptr;   // per (sample, pixel) radiance, summed in sample order by k_wf_accumulate
};
constexpr int kWfArraysTotal = 2 * kWfStateArrays + 2 + 4;   // 4-byte arrays of `capacity` entries in the slab

#define VN_DECLARE_KERNEL_API                                                                                          \
    int max_blocks_per_sm(int threads, size_t smem_bytes, bool scene_in_smem, bool count, bool octant, bool wide, bool grid, bool lean, int global_ctas); \
    cudaError_t launch_render_persistent(const RenderLaunch& p, const KernelConfig& cfg, cudaStream_t stream);         \
    cudaError_t launch_tonemap(const float4* accum, float scale, uint32_t* image, uint64_t n, cudaStream_t stream);    \
    cudaError_t launch_unit_items(const uint32_t* order, uint32_t n, uint32_t units_log2, uint32_t* items, cudaStream_t stream); \
    cudaError_t launch_tile_keys(const uint32_t* cost, uint32_t stride, uint32_t n, uint32_t mode, uint32_t spp, uint32_t* keys, uint32_t* vals, uint32_t* n_all_miss, cudaStream_t stream); \
    cudaError_t launch_reduce_tonemap_peers(const float4* const* peers, uint32_t n_peers, float scale, uint64_t begin, \
                                            uint64_t end, float4* accum_out, uint32_t* image, const uint32_t* const* peer_flags, \
                                            uint32_t wait_value, uint32_t* error, cudaStream_t stream);               \
    cudaError_t launch_signal(uint32_t* flag, uint32_t value, cudaStream_t stream);                                    \
    cudaError_t launch_wait_flags(const uint32_t* const* flags, uint32_t n, uint32_t value, uint32_t* error, cudaStream_t stream); \
    cudaError_t launch_test_rng(const uint32_t* v0, const uint32_t* v1, uint64_t n, uint32_t n_draws, uint32_t* seeds, \
                                uint32_t* lcg_out, float* rnd_out, cudaStream_t stream);                               \
    cudaError_t launch_trace_rays(const RenderLaunch& scene, const float* o, const float* d, uint64_t n, float* t_out, \
                                  int32_t* prim_out, const uint32_t* orig, bool grid, cudaStream_t stream);                     \
    cudaError_t launch_make_color(const float* rgb, uint64_t n, uint32_t* out, cudaStream_t stream);                   \
    cudaError_t launch_scatter(uint32_t type, float4 mat, const float* dirs, const float* normals,                     \
                               const uint8_t* front, const uint32_t* seeds, uint64_t n, float* dirs_out,               \
                               uint8_t* scattered, uint32_t* seeds_out, cudaStream_t stream);                          \
    cudaError_t launch_wavefront(const RenderLaunch& p, const WavefrontBuffers& wf, int num_sms, cudaStream_t stream,   \
                                 uint32_t* launches);                                                                  \
    cudaError_t launch_accumulate_samples(const RenderLaunch& p, const float* sample_rgb, cudaStream_t stream);

namespace exact { VN_DECLARE_KERNEL_API }
namespace fast { VN_DECLARE_KERNEL_API }

}  // namespace vn
