// vn_api.cu -- implementation of the C ABI declared in include/venusaur_b200.h.
// Host-side plumbing only: device memory, one stream per handle, CUDA-event timing, launch configuration.
// All compute happens in the kernels of lbvh.cu / path_kernels.cu / wavefront.cu; there is no CPU path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/venusaur/Camera.h"
#include "../../include/venusaur/Scene.h"
#include "kernels.h"
#include "grid.h"
#include "lbvh.h"
#include "lbvh_core.cuh"

using namespace vn;

#include <cstddef>
static_assert(sizeof(vn_sphere) == 36 && sizeof(vn_node32) == 32, "ABI");
static_assert(sizeof(GridHeader) == 104, "ABI (vn_read_grid)");
static_assert(sizeof(vn_params) == 96 && offsetof(vn_params, origin) == 32 && offsetof(vn_params, flags) == 92, "ABI");
static_assert(sizeof(vn_stats) == 64 && sizeof(vn_bvh_info) == 48, "ABI");

constexpr uint32_t kStatSlots = 8;

struct vn_context {
    int device = 0;
    int num_sms = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string last_error;

    vn_sphere* d_spheres = nullptr;
    uint64_t n_spheres = 0;
    bool have_spheres = false, bvh_valid = false;
    LbvhScene scene;
    LbvhWorkspace bvh_ws;

    uint32_t width = 0, height = 0;
    float4* accum_own = nullptr;
    float4* accum = nullptr;          // own or external
    uint32_t* image_tmp = nullptr;    // device staging for VN_IMAGE_HOST
    uint64_t image_tmp_pixels = 0;
    // VN_IMAGE_HOST | VN_ASYNC: frame k's D2H runs on its own stream under frame k+1's kernel (two staging buffers)
    uint32_t* image_pipe[2] = {nullptr, nullptr};
    uint64_t image_pipe_pixels = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t tail_stream = nullptr;         // second launch of a split frame (see "split_tail")
    cudaEvent_t ev_tail[2] = {nullptr, nullptr};
    cudaEvent_t ev_frame[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    bool copied_valid[2] = {false, false};
    int pipe_flip = 0;

    uint32_t* d_unit_items = nullptr;           // sample-range units of k_render_lean<kGlobal> (kernels.h::RenderLaunch::units_log2): the launch's work items,
    float4* d_unit_carry = nullptr;             //   and the per-pixel hand-over: 32 bytes {sum, -, seed, unit_epoch + finished units, -, -}
    size_t unit_tiles_cap = 0;
    uint32_t unit_epoch = 0;
    float4* d_steal_scratch = nullptr;          // sample stealing in the drain of k_render_lean (kernels.h::RenderLaunch::steal_scratch): lanes x (spp + 1) float4
    uint32_t* d_steal_count = nullptr;          // one counter per lane of the grid, zero between launches
    size_t steal_scratch_cap = 0, steal_count_cap = 0;
    uint32_t* d_timeline = nullptr;             // 2048 words, VN_COUNTERS launches of k_render_lean (kernels.h::RenderLaunch::timeline)
    uint32_t* d_flags = nullptr;                // 64 words: [0..61] epoch flags for cross-process ordering (vn_signal / vn_wait_flags), [63] error word
    unsigned long long* d_counters = nullptr;   // 4 x u64 + work ticket (u32) at +32 bytes; [8..11) the launch timeline of the instrumented k_render_lean / k_render_async
    unsigned long long* h_counters = nullptr;   // pinned: kStatSlots x 256 bytes, one slot per launch in flight (VN_ASYNC renders never wait for each other on the host)
    cudaEvent_t ev_slot[8][2] = {};             // begin / end of the launch that owns the slot
    uint32_t slot_head = 0, slots_pending = 0;  // launches whose counters have not been folded into stats yet: slots [head - pending, head)
    uint32_t slot_launches[8] = {};

    WavefrontBuffers wf;
    uint64_t wf_sample_floats_ = 0;

    // options
    uint32_t leaf_size = 0;           // 0 = auto: 3 with SAH splits (small scenes), 2 with Karras splits
    float aabb_pad = 0.01f;
    uint32_t wide_max_prims = 16384;  // scenes up to this size also get 4-wide nodes (octant-sorted copies in shared memory when they fit); 0 = never.
                                      // Larger values build them for any scene (one launch pair per level) for the opt-in "wide_global" traversal
    bool wide_global = false;         // traverse canonical wide nodes from L2/HBM when the scene does not fit in shared memory (measured slower
                                      // than the pair nodes on 1 M / 16 M spheres: 1.10 vs 1.45 and 0.69 vs 0.80 Grays/s)
    GridScene grid;                   // uniform grid + oversize list (small scenes), built behind the LBVH on the same sorted spheres
    uint32_t accel = 1;               // closest-hit structure of the path kernel: 1 = BVH (default: what the north star specifies), 0 = auto (the grid
                                      // when the scene suits it: +3..6 % on RTIOW), 2 = grid
    uint32_t grid_vote = 0;           // see closest_hit_grid_vote (path_kernels.cu): 0 = while-while (fastest measured), n = voted sphere turns
    uint32_t grid_max_per_cell = 16;  // a cell with more spheres than this disqualifies the grid (clustered scenes: the BVH adapts, a grid does not)
    uint32_t last_accel = 0;          // what the last vn_render traversed: 1 pair nodes, 2 wide nodes (shared memory), 3 wide nodes (L2/HBM), 4 grid
    float huge_factor = 50.0f;        // spheres with radius > huge_factor x median are tested before the wide traversal (0 = none), lbvh_core.cuh::HugeList
    int wide_threads = 1024;          // CTA size of the wide-node path kernel (one CTA per SM): 512, 768 or 1024 (64 registers per lane at 1024)
    uint32_t leaf_vote = 0;           // see closest_hit_wide_vote (path_kernels.cu); 0 = while-while
    // cost-ordered tile schedule (prepare_tile_order): per-tile ray segments of the previous launch of the same view, sorted descending
    uint32_t tile_order_opt = 1;      // "tile_order": 0 = row-major tickets, 1 = tiles sorted by the ray segments of their 32 pixels, 2 = by their most expensive pixel, 3 = by max(sum / 8, most expensive pixel), 4 = like 1 with the all-miss tiles last
    uint32_t* d_tile_cost = nullptr;  // [2 * tile_cap]: sum | max of the pixels' ray segments
    uint32_t* d_tile_sort = nullptr;  // [4 * tile_cap]: keys, values and their alternates for the radix sort
    const uint32_t* d_tile_order = nullptr;
    uint32_t tile_cap = 0;
    bool tile_guess = false;          // the collecting launch of the current view may use the previous view's order
    uint32_t tile_guess_opt = 1;      // "tile_guess": 1 = after a camera / scene change the collecting launch hands the tiles out in the previous view's order (0: row-major).
                                      // (Scattered tickets -- t * golden-ratio mod n -- were tried for views without a predecessor: 6.06-6.26 instead of 5.79 ms, expensive tiles then start at random times up to the end)
    double tile_seg_per_path = 0.0;   // ray segments per path of the current view's cost-collecting launch (0: not measured)
    uint32_t tile_all_miss = 0;       // tiles of the current view in which no path hit anything while the costs were collected: the tail of d_tile_order
    int tile_state = 0;               // 0: nothing known (the next launch collects costs), 1: costs collected (sort before the next launch), 2: order valid
    struct TileSig { uint32_t w, h, r0, r1, spp, depth; float cam[13]; uint32_t pad_; uint64_t epoch; } tile_sig{};   // no implicit padding
    uint64_t bvh_epoch = 0;
    uint32_t hit_gate = 1;            // "hit_gate": 1 = scenes traversed from L2 / HBM apply the hit-point gate (vn_math.cuh::hit_gate_ok), 0 = never,
                                      // 2 = the pair-node kernels apply it to small scenes too (the shared-memory wide-node kernels never do)
    uint32_t lean = 1;                // "lean": k_render_lean (16-bit links, no per-lane statistics, no spills) when the launch qualifies; 0 = k_render_async
    float split_tail = 0.5f;          // "split_tail": fraction of the cost-ordered tiles (its cheap end) that a launch of the shared-memory path kernel hands to a SECOND
                                      // launch on another stream.  The first launch's drain -- a few lanes per SM finishing their last, heavy pixels while the rest of
                                      // the GPU idles -- then overlaps the second launch, whose CTAs start on every SM the first one leaves; the second launch's own
                                      // tiles drain in a few tens of microseconds.  Only tiles that saw nothing but sky while the view's costs were collected qualify
                                      // (uniform and cheap); the value caps their share.  Measured on RTIOW 1080p with a fixed fraction of the cost-ordered tiles:
                                      // 0 / 0.10 / 0.15 / 0.18 / 0.25 -> 5.13 / 5.09 / 5.09 / 5.31 / 5.33 ms per launch: beyond the sky (~17 % of the tiles) the second
                                      // launch has a heavy tail of its own; 0 = one launch
    uint32_t qnodes_opt = 1;          // "qnodes": scenes traversed from L2 / HBM are traversed through the quantised pairs (32 bytes = one 256-bit request per step
                                      // instead of two; 16-bit planes over the root box); 0 = the packed fp32 pairs
    uint32_t units = 4;               // "units": scenes traversed from L2 / HBM hand a tile's samples out in this many ranges (1, 2, 4, 8 or 16; path_kernels.cu, finish_unit):
                                      // a pixel of a million-sphere scene is 20 ms of one lane's time, and a launch ends with whole pixels that were started late
    double units_min_seg = 13.0;      // "units_min_seg": ... but only for views whose paths are at least this many ray segments long on average (measured by the view's
                                      // cost-collecting launch).  Units chain a pixel's ranges one behind the other; where most pixels are cheap and a few bounce 50
                                      // times those chains ARE the end of the launch, and whole pixels + the stealing drain do better.  Random scenes at 1080p,
                                      // ms per launch with 1 / 4 units (tools/units_probe.py): 2 k spheres, 1.6 segments per path: 8.5 / 13.1; 8 k, 2.9: 21.8 / 27.9;
                                      // 30 k, 6.9: 48.6 / 51.9; 100 k, 11.4: 83.6 / 84.2; 300 k, 15.4: 122.2 / 115.1; 1 M, 18.4: 169.1 / 156.5.
                                      // 0 = units whatever the view (and without a measurement)
    uint32_t multi_subframes = 64;    // "multi_subframes": vn_render_subframes renders up to this many subframes per launch (k_render_lean<kMulti>: one drain instead of
                                      // one per subframe; the pixel's subframes are blended in order through a tag in accum.w); 0 or 1 = one launch per subframe
    uint32_t steal = 1;               // "steal": once the tile tickets are exhausted, idle lanes of a warp take single samples of the pixels its other lanes still hold
                                      // (k_render_lean's drain, path_kernels.cu::lean_drain); the value = the fewest samples a lane must have left to give one away, 0 = off
    uint32_t steal_smem = 0;          // "steal_smem": also for scenes traversed from shared memory.  Off: measured on RTIOW 1080p the drain shrinks from 0.39 to 0.28 ms
                                      // but the kernel variant that carries the drain's code runs its first loop 2 % slower (18.75 -> 18.52 Grays/s); scenes traversed
                                      // from L2 / HBM gain 9 % (1 M spheres) and 4 % (16 M spheres)
    uint32_t warp_tiles = 1;          // "warp_tiles": k_render_async phase form hands whole tiles to warps (see path_kernels.cu); 0 = lanes take single pixels
    uint32_t async_done = 26;         // k_render_async: a traversal burst ends when this many lanes hold a finished ray (0 = k_render_persistent)
    int global_ctas = 6;              // "global_ctas": 4, 5 or 6 CTAs of 256 threads per SM for the L2 / HBM form (64 / 48 / 40 registers: more warps to hide latency,
                                      // a few spills each).  With the fp32 pairs 4 / 5 / 6 measured 3046 / 3037 / 2961 (1 M spheres) and 1726 / 1765 / 1779 (16 M); with the
                                      // quantised pairs the traversal is less bound by the memory system and more by latency: 3581 / 3886 / 3945 and - / 2225 / 2304 (8: 3372 / 1949)
    uint32_t global_done = 16;        // the same threshold for scenes traversed from L2 / HBM (k_render_lean<kGlobal>): long traversals, so shade earlier (measured:
                                      // 1 M spheres 2.21 -> 2.49 Grays/s, 16 M spheres 1.11 -> 1.60 against 26)
    uint32_t async_node = 0, async_leaf = 8;   // async_node 0 = phase form (no votes inside the node / leaf phases), the default
    bool wide_nodes = true;           // use them when they fit in shared memory
    uint32_t sah_max_prims = 4096;    // scenes up to this size get SAH splits (k_sah_small); 0 = always Karras
    int threads = 256;
    int blocks_per_sm = 0;            // 0 = occupancy
    size_t smem_scene_limit = 100 * 1024;
    uint32_t wavefront_slots = 1u << 21;
    uint32_t wavefront_wide = 0;      // "wavefront_wide": the wavefront kernel's extend phase traverses the path kernels' 4-wide nodes in shared memory (one 1024-thread CTA per SM).
                                      // Measured on RTIOW 1080p: 6.1 Grays/s against 7.1 on the pair nodes (several 256-thread CTAs per SM) and 19.0 for k_render_lean on the
                                      // same wide nodes: the queue traffic and the four grid barriers per bounce bound this schedule, not its closest-hit structure
    bool octant_nodes = true;         // stage the BVH nodes once per ray octant when 8 copies fit in shared memory

    vn_stats stats{};
};

namespace {

std::mutex g_err_mutex;
std::string g_create_error;

void set_global_error(const std::string& s) {
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_create_error = s;
}

int fail(vn_context* c, int status, const std::string& msg) {
    if (c) c->last_error = msg;
    set_global_error(msg);
    return status;
}

#define VN_CUDA(c, call)                                                                                      \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail((c), e_ == cudaErrorMemoryAllocation ? VN_ERR_OOM : VN_ERR_CUDA,                      \
                        std::string("CUDA call (") + #call + ") failed with error: '" + cudaGetErrorString(e_) + \
                            "' (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");                         \
    } while (0)

#define VN_REQUIRE(c, cond, msg) \
    do { if (!(cond)) return fail((c), VN_ERR_INVALID, std::string(msg)); } while (0)

inline f3 to_f3(const float* p) { return mk3(p[0], p[1], p[2]); }

void free_wavefront(WavefrontBuffers& w) {
    cudaFree(w.slab);
    cudaFree(w.counts);
    cudaFree(w.sample_rgb);
    w = WavefrontBuffers();
}

bool scene_fits_smem(const vn_context* c) {
    const size_t need = scene_smem_bytes((uint32_t)c->scene.num_nodes, (uint32_t)c->scene.n);
    return c->scene.n > 0 && need <= c->smem_scene_limit && need + 1024 <= c->smem_optin;
}

int fill_launch(vn_context* c, const vn_params* p, RenderLaunch& L) {
    memset(&L, 0, sizeof(L));
    L.cam.origin = to_f3(p->origin);
    L.cam.u = to_f3(p->u);
    L.cam.v = to_f3(p->v);
    L.cam.w = to_f3(p->w);
    // loop invariants of get_ray (RayTracer.cu:155-156); host IEEE float == the device's exact build
    L.cam.u_unit = normalize(L.cam.u);
    L.cam.v_unit = normalize(L.cam.v);
    L.cam.lens_radius = p->lens_radius;
    L.cam.wm1 = (float)(p->width - 1u);
    L.cam.hm1 = (float)(p->height - 1u);
    L.cam.inv_wm1 = 1.0f / L.cam.wm1;
    L.cam.inv_hm1 = 1.0f / L.cam.hm1;
    L.cam.div_exact = (div_by_const_ok(L.cam.wm1) && div_by_const_ok(L.cam.hm1)) ? 1u : 0u;
    L.width = p->width; L.height = p->height;
    L.spp = p->samples_per_pixel;
    L.subframe_index = p->subframe_index;
    L.max_depth = p->max_depth;
    L.row_begin = p->row_begin; L.row_end = p->row_end;
    if (L.row_begin == 0 && L.row_end == 0) L.row_end = p->height;
    if (p->flags & VN_ACCUM_SUM) { L.blend_mode = kBlendSum; L.blend_a = 0.0f; }
    else if (p->accum_count > 0) { L.blend_mode = kBlendLerp; L.blend_a = 1.0f / (float)(p->accum_count + 1u); }   // RayTracer.cu:210
    else { L.blend_mode = kBlendOverwrite; L.blend_a = 1.0f; }
    L.n_sub = 1u; L.tiles_per_sub = 0u; L.accum_count = p->accum_count; L.sub_stride = 1u;
    L.inv_spp = 1.0f / (float)p->samples_per_pixel;                                                                // vec_math.h:483-487
    L.accum = c->accum;
    L.nodes = c->scene.nodes; L.geom = c->scene.geom; L.mat = c->scene.mat; L.type = c->scene.type;
    L.root_link = c->scene.root_link;
    L.num_nodes = (uint32_t)c->scene.num_nodes;
    L.num_spheres = (uint32_t)c->scene.n;
    L.wide = c->scene.wide; L.num_wide = c->scene.num_wide; L.wide_root = 0u;
    L.huge = c->scene.huge;
    L.leaf_vote = c->leaf_vote;
    L.async_done = c->async_done; L.async_node = c->async_node; L.async_leaf = c->async_leaf;
    L.grid_vote = c->grid_vote;
    L.steal = c->steal;
    L.qnodes = nullptr;
    for (int a = 0; a < 3; a++) { L.q_lo[a] = 0.0f; L.q_scale[a] = 0.0f; }
    if (c->qnodes_opt && c->scene.qnodes && !scene_fits_smem(c)) {
        L.qnodes = c->scene.qnodes;
        for (int a = 0; a < 3; a++) {
            const float e = c->scene.bounds_hi[a] - c->scene.bounds_lo[a];
            L.q_lo[a] = c->scene.bounds_lo[a];
            L.q_scale[a] = e > 0.0f ? e / 65535.0f : 0.0f;
        }
    }
    L.timeline = nullptr;
    L.gate = (c->hit_gate == 2u || (c->hit_gate == 1u && !scene_fits_smem(c))) ? 1u : 0u;
    L.grid = c->grid.h; L.grid_start = c->grid.start; L.grid_refs = c->grid.refs;
    L.counters = c->d_counters;
    L.work_counter = reinterpret_cast<uint32_t*>(c->d_counters + 4);
    const uint32_t rows = L.row_end - L.row_begin;
    L.tile_order = nullptr; L.tile_cost = nullptr; L.tile_cost_stride = c->tile_cap;
    L.tiles_x = (p->width + 7u) / 8u;
    L.tiles_x_inv = L.tiles_x > 1u ? (uint32_t)(0x100000000ull / L.tiles_x) : 0xFFFFFFFFu;
    L.total_work = L.tiles_x * ((rows + 3u) / 4u) * 32u;
    return VN_OK;
}

}  // namespace

// scratch device buffer for the unit-level test entry points
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 4); }
    template <typename T> T* as() { return static_cast<T*>(p); }
};
}  // namespace

extern "C" void vn_destroy(vn_handle c);

// everything vn_create allocates; on failure the caller destroys the half-built context
static int create_resources(vn_context* c) {
    VN_CUDA(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    VN_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10)
        return fail(c, VN_ERR_NO_DEVICE, std::string("vn_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                             "; this library contains sm_100a code only");
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    VN_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& ev : c->ev) VN_CUDA(c, cudaEventCreate(&ev));
    VN_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    VN_CUDA(c, cudaStreamCreateWithFlags(&c->tail_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev_tail) VN_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        VN_CUDA(c, cudaEventCreateWithFlags(&c->ev_frame[i], cudaEventDisableTiming));
        VN_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    }
    VN_CUDA(c, cudaMalloc(&c->d_counters, 256));
    VN_CUDA(c, cudaMemset(c->d_counters, 0, 256));
    VN_CUDA(c, cudaMalloc(&c->d_timeline, 16384));
    VN_CUDA(c, cudaMemset(c->d_timeline, 0, 16384));
    VN_CUDA(c, cudaMalloc(&c->d_flags, 256));
    VN_CUDA(c, cudaMemset(c->d_flags, 0, 256));
    VN_CUDA(c, cudaHostAlloc(&c->h_counters, 256 * kStatSlots, cudaHostAllocDefault));
    memset(c->h_counters, 0, 256 * kStatSlots);
    for (uint32_t i = 0; i < kStatSlots; i++) for (int j = 0; j < 2; j++) VN_CUDA(c, cudaEventCreate(&c->ev_slot[i][j]));
    return VN_OK;
}

extern "C" {

const char* vn_version(void) { return "venusaur_b200 0.1 (sm_100a)"; }

int vn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* vn_last_error(vn_handle h) {
    if (h) return h->last_error.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mutex);
    copy = g_create_error;
    return copy.c_str();
}

int vn_create(int device, vn_handle* out) {
    if (!out) return fail(nullptr, VN_ERR_INVALID, "vn_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, VN_ERR_NO_DEVICE, std::string("vn_create: no CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                                   "); venusaur_b200 has no CPU path");
    }
    if (device < 0 || device >= n) return fail(nullptr, VN_ERR_INVALID, "vn_create: device index out of range");
    vn_context* c = new vn_context();
    c->device = device;
    const int rc = create_resources(c);
    if (rc != VN_OK) {                   // the message is already in the global slot (vn_last_error(NULL)); nothing may leak
        vn_destroy(c);
        return rc;
    }
    *out = c;
    return VN_OK;
}

void vn_destroy(vn_handle c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    lbvh_free(c->scene);
    grid_free(c->grid);
    lbvh_workspace_free(c->bvh_ws);
    free_wavefront(c->wf); c->wf_sample_floats_ = 0;
    cudaFree(c->d_tile_cost); cudaFree(c->d_tile_sort); cudaFree(c->d_spheres); cudaFree(c->accum_own); cudaFree(c->image_tmp); cudaFree(c->d_counters); cudaFree(c->d_flags); cudaFree(c->d_timeline); cudaFree(c->d_steal_scratch); cudaFree(c->d_steal_count); cudaFree(c->d_unit_items); cudaFree(c->d_unit_carry);
    cudaFreeHost(c->h_counters);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& pr : c->ev_slot) for (auto& ev : pr) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; i++) { cudaFree(c->image_pipe[i]); if (c->ev_frame[i]) cudaEventDestroy(c->ev_frame[i]); if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); }
    if (c->tail_stream) cudaStreamDestroy(c->tail_stream);
    for (auto& e : c->ev_tail) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* vn_stream(vn_handle c) { return c ? (void*)c->stream : nullptr; }

int vn_set_option(vn_handle c, const char* name, double value) {
    VN_REQUIRE(c, c && name, "vn_set_option: NULL argument");
    const std::string k(name);
    if (k == "leaf_size") { VN_REQUIRE(c, value >= 0 && value <= 8, "leaf_size must be in [0,8] (0 = auto)"); c->leaf_size = (uint32_t)value; c->bvh_valid = false; }
    else if (k == "aabb_pad") { VN_REQUIRE(c, value >= 0 && value < 1, "aabb_pad must be in [0,1)"); c->aabb_pad = (float)value; c->bvh_valid = false; }
    else if (k == "threads") { VN_REQUIRE(c, value == 64 || value == 128 || value == 256 || value == 512 || value == 1024, "threads must be 64, 128, 256, 512 or 1024"); c->threads = (int)value; }
    else if (k == "blocks_per_sm") { VN_REQUIRE(c, value >= 0 && value <= 32, "blocks_per_sm must be in [0,32]"); c->blocks_per_sm = (int)value; }
    else if (k == "smem_scene_limit") { VN_REQUIRE(c, value >= 0, "smem_scene_limit must be >= 0"); c->smem_scene_limit = (size_t)value; }
    else if (k == "wavefront_slots") { VN_REQUIRE(c, value >= 1024 && value <= (double)(1u << 26), "wavefront_slots out of range"); c->wavefront_slots = (uint32_t)value; free_wavefront(c->wf); c->wf_sample_floats_ = 0; }
    else if (k == "sah_max_prims") { VN_REQUIRE(c, value >= 0 && value <= 8192, "sah_max_prims must be in [0,8192]"); c->sah_max_prims = (uint32_t)value; c->bvh_valid = false; }
    else if (k == "wide_max_prims") { VN_REQUIRE(c, value >= 0 && value <= (double)(1u << 28), "wide_max_prims must be in [0,2^28]"); c->wide_max_prims = (uint32_t)value; c->bvh_valid = false; }
    else if (k == "wide_nodes") { c->wide_nodes = value != 0; }
    else if (k == "wide_global") { c->wide_global = value != 0; }
    else if (k == "accel") { VN_REQUIRE(c, value == 0 || value == 1 || value == 2, "accel must be 0 (auto), 1 (BVH) or 2 (grid)"); c->accel = (uint32_t)value; c->bvh_valid = false; }
    else if (k == "grid_vote") { VN_REQUIRE(c, value >= 0 && value <= 33, "grid_vote must be in [0,33]"); c->grid_vote = (uint32_t)value; }
    else if (k == "grid_max_per_cell") { VN_REQUIRE(c, value >= 1 && value <= 65535, "grid_max_per_cell must be in [1,65535]"); c->grid_max_per_cell = (uint32_t)value; c->bvh_valid = false; }
    else if (k == "huge_factor") { VN_REQUIRE(c, value >= 0, "huge_factor must be >= 0"); c->huge_factor = (float)value; c->bvh_valid = false; }
    else if (k == "wide_threads") { VN_REQUIRE(c, value == 512 || value == 768 || value == 1024, "wide_threads must be 512, 768 or 1024"); c->wide_threads = (int)value; }
    else if (k == "tile_order") { VN_REQUIRE(c, value >= 0 && value <= 4, "tile_order must be 0..4"); c->tile_order_opt = (uint32_t)value; c->tile_state = 0; }
    else if (k == "warp_tiles") { c->warp_tiles = value != 0 ? 1u : 0u; }
    else if (k == "tile_guess") { c->tile_guess_opt = value != 0 ? 1u : 0u; }
    else if (k == "wavefront_wide") { c->wavefront_wide = value != 0 ? 1u : 0u; }
    else if (k == "split_tail") { VN_REQUIRE(c, value >= 0 && value <= 0.9, "split_tail must be in [0,0.9]"); c->split_tail = (float)value; }
    else if (k == "qnodes") { c->qnodes_opt = value != 0 ? 1u : 0u; c->bvh_valid = false; }
    else if (k == "multi_subframes") { VN_REQUIRE(c, value >= 0 && value <= 64, "multi_subframes must be in [0,64]"); c->multi_subframes = (uint32_t)value; }
    else if (k == "units_min_seg") { VN_REQUIRE(c, value >= 0, "units_min_seg must be >= 0"); c->units_min_seg = value; }
    else if (k == "units") { VN_REQUIRE(c, value == 1 || value == 2 || value == 4 || value == 8 || value == 16, "units must be 1, 2, 4, 8 or 16"); c->units = (uint32_t)value; }
    else if (k == "steal_smem") { c->steal_smem = value != 0 ? 1u : 0u; }
    else if (k == "steal") { VN_REQUIRE(c, value >= 0 && value <= 1023, "steal must be in [0,1023]"); c->steal = (uint32_t)value; }
    else if (k == "lean") { c->lean = value != 0 ? 1u : 0u; }
    else if (k == "hit_gate") { VN_REQUIRE(c, value == 0 || value == 1 || value == 2, "hit_gate must be 0, 1 or 2"); c->hit_gate = (uint32_t)value; }
    else if (k == "global_ctas") { VN_REQUIRE(c, value == 4 || value == 5 || value == 6, "global_ctas must be 4, 5 or 6"); c->global_ctas = (int)value; }
    else if (k == "global_done") { VN_REQUIRE(c, value >= 1 && value <= 32, "global_done must be in [1,32]"); c->global_done = (uint32_t)value; }
    else if (k == "async_done") { VN_REQUIRE(c, value >= 0 && value <= 32, "async_done must be in [0,32]"); c->async_done = (uint32_t)value; }
    else if (k == "async_node") { VN_REQUIRE(c, value >= 0 && value <= 32, "async_node must be in [0,32] (0 = phase form: no votes inside the node / leaf phases)"); c->async_node = (uint32_t)value; }
    else if (k == "async_leaf") { VN_REQUIRE(c, value >= 1 && value <= 32, "async_leaf must be in [1,32]"); c->async_leaf = (uint32_t)value; }
    else if (k == "leaf_vote") { VN_REQUIRE(c, value >= 0 && value <= 32, "leaf_vote must be in [0,32]"); c->leaf_vote = (uint32_t)value; }
    else if (k == "octant_nodes") { c->octant_nodes = value != 0; }
    else return fail(c, VN_ERR_INVALID, "vn_set_option: unknown option '" + k + "'");
    return VN_OK;
}

int vn_set_spheres(vn_handle c, const vn_sphere* host_spheres, uint64_t n) {
    VN_REQUIRE(c, c, "vn_set_spheres: NULL handle");
    VN_REQUIRE(c, n == 0 || host_spheres, "vn_set_spheres: NULL spheres");
    for (uint64_t i = 0; i < n; i++) VN_REQUIRE(c, host_spheres[i].type <= 2u, "vn_set_spheres: material type must be 0, 1 or 2");
    VN_CUDA(c, cudaSetDevice(c->device));
    cudaFree(c->d_spheres);
    c->d_spheres = nullptr;
    c->bvh_valid = false;
    c->have_spheres = false;
    if (n) {
        VN_CUDA(c, cudaMalloc(&c->d_spheres, n * sizeof(vn_sphere)));
        VN_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
        VN_CUDA(c, cudaMemcpyAsync(c->d_spheres, host_spheres, n * sizeof(vn_sphere), cudaMemcpyHostToDevice, c->stream));
        VN_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
        VN_CUDA(c, cudaStreamSynchronize(c->stream));
        VN_CUDA(c, cudaEventElapsedTime(&c->stats.ms_upload, c->ev[0], c->ev[1]));
    }
    c->n_spheres = n;
    c->have_spheres = true;
    return VN_OK;
}

int vn_build_bvh(vn_handle c) {
    VN_REQUIRE(c, c, "vn_build_bvh: NULL handle");
    VN_REQUIRE(c, c->have_spheres, "vn_build_bvh: call vn_set_spheres first");
    VN_CUDA(c, cudaSetDevice(c->device));
    // a rebuild frees the old scene arrays before it can fail: nothing may render from them until the new build has succeeded
    c->bvh_valid = false;
    grid_free(c->grid);
    std::string err;
    uint32_t launches = 0;
    VN_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    // leaf_size 0 = auto.  Small (SAH-split) scenes: the smallest leaf whose 4-wide nodes still fit in shared memory next to the
    // spheres (one-sphere leaves need no sphere loop and test the fewest spheres; RTIOW: 238 wide nodes = 213 KB of the 227 KB);
    // larger scenes: 2.
    const bool small = c->n_spheres >= 2 && c->n_spheres <= c->sah_max_prims && c->n_spheres <= c->wide_max_prims && c->n_spheres <= 16384;
    int rc = 0;
    for (uint32_t leaf_size = c->leaf_size ? c->leaf_size : (small ? 1u : 2u);; leaf_size++) {
        rc = lbvh_build(c->d_spheres, c->n_spheres, leaf_size, c->aabb_pad, c->sah_max_prims, c->wide_max_prims, c->huge_factor, c->num_sms, c->stream, c->scene, c->bvh_ws, &launches, err);
        if (rc != 0 || c->leaf_size || !small) break;
        if (c->scene.num_wide > 0 && c->scene.wide_levels <= kWideMaxLevels && wide_smem_bytes(c->scene.num_wide, (uint32_t)c->scene.n) + 2048 <= c->smem_optin) break;
        if (leaf_size >= 4u) {      // the wide copies never fit: the pair nodes will be traversed, which like 3 spheres per leaf best
            rc = lbvh_build(c->d_spheres, c->n_spheres, 3u, c->aabb_pad, c->sah_max_prims, c->wide_max_prims, c->huge_factor, c->num_sms, c->stream, c->scene, c->bvh_ws, &launches, err);
            break;
        }
    }
    if (rc != 0) return fail(c, rc == -1 ? VN_ERR_INVALID : VN_ERR_CUDA, "vn_build_bvh: " + err);
    if (c->qnodes_opt && !scene_fits_smem(c)) {       // scenes traversed from L2 / HBM: the pairs again in 32 bytes each (lbvh.cu::k_quantize_pairs)
        if (lbvh_quantize(c->scene, c->stream, &launches, err) < 0) return fail(c, VN_ERR_CUDA, "vn_build_bvh: " + err);
    }
    if (c->accel != 1u && c->scene.geom) {
        const int grc = grid_build(c->scene.geom, c->scene.n, c->grid_max_per_cell, c->stream, c->grid, &launches, err);
        if (grc < 0) return fail(c, VN_ERR_CUDA, "vn_build_bvh: " + err);
    }
    VN_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    VN_CUDA(c, cudaEventElapsedTime(&c->stats.ms_build, c->ev[0], c->ev[1]));
    c->stats.kernel_launches_total += launches;
    c->bvh_valid = true;
    c->bvh_epoch += 1;
    return VN_OK;
}

int vn_update_spheres(vn_handle c, const vn_sphere* host_spheres, uint64_t n) {
    VN_REQUIRE(c, c && host_spheres, "vn_update_spheres: NULL argument");
    VN_REQUIRE(c, c->have_spheres && n == c->n_spheres && n > 0, "vn_update_spheres: the sphere count must be the one given to vn_set_spheres");
    for (uint64_t i = 0; i < n; i++) VN_REQUIRE(c, host_spheres[i].type <= 2u, "vn_update_spheres: material type must be 0, 1 or 2");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));                 // no launch may still be reading the old records through the BVH arrays
    VN_CUDA(c, cudaMemcpyAsync(c->d_spheres, host_spheres, n * sizeof(vn_sphere), cudaMemcpyHostToDevice, c->stream));
    if (!c->bvh_valid) return VN_OK;
    // refit-only rebuild: same hierarchy, new boxes (lbvh.cu::lbvh_refit); a full build when the hierarchy is no longer at hand
    c->bvh_valid = false;
    std::string err;
    uint32_t launches = 0;
    VN_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    const int rc = lbvh_refit(c->d_spheres, c->aabb_pad, c->huge_factor, c->stream, c->scene, c->bvh_ws, &launches, err);
    if (rc < 0) return fail(c, VN_ERR_CUDA, "vn_update_spheres: " + err);
    if (rc == 1 || c->grid.valid) return vn_build_bvh(c);
    if (c->qnodes_opt && !scene_fits_smem(c)) {
        if (lbvh_quantize(c->scene, c->stream, &launches, err) < 0) return fail(c, VN_ERR_CUDA, "vn_update_spheres: " + err);
    }
    VN_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    VN_CUDA(c, cudaEventElapsedTime(&c->stats.ms_build, c->ev[0], c->ev[1]));
    c->stats.kernel_launches_total += launches;
    c->bvh_valid = true;
    c->bvh_epoch += 1;
    return VN_OK;
}

int vn_get_bvh_info(vn_handle c, vn_bvh_info* out) {
    VN_REQUIRE(c, c && out, "vn_get_bvh_info: NULL argument");
    VN_REQUIRE(c, c->bvh_valid, "vn_get_bvh_info: no BVH (call vn_build_bvh)");
    memset(out, 0, sizeof(*out));
    out->num_spheres = c->scene.n;
    out->num_nodes = c->scene.num_nodes;
    out->max_leaf_size = c->scene.leaf_size;
    out->scene_in_smem = scene_fits_smem(c) ? 1u : 0u;
    for (int a = 0; a < 3; a++) { out->bounds_lo[a] = c->scene.bounds_lo[a]; out->bounds_hi[a] = c->scene.bounds_hi[a]; }
    return VN_OK;
}

int vn_read_bvh(vn_handle c, vn_node32* host_nodes, uint64_t cap_nodes, uint32_t* host_prim_order, uint64_t cap_prims) {
    VN_REQUIRE(c, c, "vn_read_bvh: NULL handle");
    VN_REQUIRE(c, c->bvh_valid, "vn_read_bvh: no BVH (call vn_build_bvh)");
    VN_CUDA(c, cudaSetDevice(c->device));
    if (host_nodes && c->scene.nodes)
        VN_CUDA(c, cudaMemcpy(host_nodes, c->scene.nodes, std::min<uint64_t>(cap_nodes, c->scene.num_nodes) * 32, cudaMemcpyDeviceToHost));
    if (host_prim_order && c->scene.orig)
        VN_CUDA(c, cudaMemcpy(host_prim_order, c->scene.orig, std::min<uint64_t>(cap_prims, c->scene.n) * 4, cudaMemcpyDeviceToHost));
    return VN_OK;
}

int vn_read_sched_counters(vn_handle c, uint64_t* out14) {
    VN_REQUIRE(c, c && out14, "vn_read_sched_counters: NULL argument");
    if (c->slots_pending) { const int rc = vn_synchronize(c); if (rc != VN_OK) return rc; }
    const unsigned long long* last = c->h_counters + 32u * ((c->slot_head + kStatSlots - 1u) % kStatSlots);
    for (int i = 0; i < 14; i++) out14[i] = last[8 + i];
    return VN_OK;
}

int vn_read_timeline(vn_handle c, uint32_t* out2048) {
    VN_REQUIRE(c, c && out2048, "vn_read_timeline: NULL argument");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaMemcpyAsync(out2048, c->d_timeline, 8192, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_read_timeline_ex(vn_handle c, uint32_t* out4096) {
    VN_REQUIRE(c, c && out4096, "vn_read_timeline_ex: NULL argument");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaMemcpyAsync(out4096, c->d_timeline, 16384, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_read_grid(vn_handle c, void* header104, uint16_t* host_start, uint64_t cap_start, uint16_t* host_refs, uint64_t cap_refs) {
    VN_REQUIRE(c, c, "vn_read_grid: NULL handle");
    VN_REQUIRE(c, c->bvh_valid, "vn_read_grid: no scene structure (call vn_build_bvh)");
    if (!c->grid.valid) return 0;
    VN_CUDA(c, cudaSetDevice(c->device));
    if (header104) memcpy(header104, &c->grid.h, sizeof(GridHeader));
    if (host_start) VN_CUDA(c, cudaMemcpy(host_start, c->grid.start, std::min<uint64_t>(cap_start, c->grid.h.n_cells + 1ull) * 2, cudaMemcpyDeviceToHost));
    if (host_refs) VN_CUDA(c, cudaMemcpy(host_refs, c->grid.refs, std::min<uint64_t>(cap_refs, c->grid.h.n_refs) * 2, cudaMemcpyDeviceToHost));
    return 1;
}

int vn_read_huge(vn_handle c, uint32_t* idx8) {
    VN_REQUIRE(c, c, "vn_read_huge: NULL handle");
    VN_REQUIRE(c, c->bvh_valid, "vn_read_huge: no BVH (call vn_build_bvh)");
    for (uint32_t i = 0; idx8 && i < c->scene.huge.n; i++) idx8[i] = c->scene.huge.idx[i];
    return (int)c->scene.huge.n;
}

int vn_last_accel(vn_handle c) { return c ? (int)c->last_accel : 0; }

int vn_read_wide_bvh(vn_handle c, float* host_nodes, uint64_t cap_nodes, uint32_t* num_nodes_out, uint32_t* levels_out) {
    VN_REQUIRE(c, c, "vn_read_wide_bvh: NULL handle");
    VN_REQUIRE(c, c->bvh_valid, "vn_read_wide_bvh: no BVH (call vn_build_bvh)");
    VN_CUDA(c, cudaSetDevice(c->device));
    if (num_nodes_out) *num_nodes_out = c->scene.num_wide;
    if (levels_out) *levels_out = c->scene.wide_levels;
    if (host_nodes && c->scene.wide && c->scene.num_wide)
        VN_CUDA(c, cudaMemcpy(host_nodes, c->scene.wide, std::min<uint64_t>(cap_nodes, c->scene.num_wide) * 128, cudaMemcpyDeviceToHost));
    return VN_OK;
}

int vn_morton_codes(vn_handle c, uint32_t* codes_out, uint64_t cap) {
    VN_REQUIRE(c, c && codes_out, "vn_morton_codes: NULL argument");
    VN_REQUIRE(c, c->bvh_valid, "vn_morton_codes: no BVH (call vn_build_bvh)");
    VN_CUDA(c, cudaSetDevice(c->device));
    if (c->scene.codes) VN_CUDA(c, cudaMemcpy(codes_out, c->scene.codes, std::min<uint64_t>(cap, c->scene.n) * 4, cudaMemcpyDeviceToHost));
    return VN_OK;
}

int vn_resize(vn_handle c, uint32_t width, uint32_t height) {
    VN_REQUIRE(c, c, "vn_resize: NULL handle");
    VN_REQUIRE(c, width >= 2 && height >= 2, "vn_resize: width and height must be >= 2 (the camera divides by width-1, RayTracer.cu:173)");
    VN_REQUIRE(c, (uint64_t)width * height < (1ull << 31), "vn_resize: too many pixels");
    VN_CUDA(c, cudaSetDevice(c->device));
    if (width != c->width || height != c->height || !c->accum_own) {
        VN_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->accum_own);
        c->accum_own = nullptr;
        VN_CUDA(c, cudaMalloc(&c->accum_own, (size_t)width * height * sizeof(float4)));
        c->accum = c->accum_own;   // an external buffer does not survive a resize
        c->width = width; c->height = height;
    }
    VN_CUDA(c, cudaMemsetAsync(c->accum, 0, (size_t)width * height * sizeof(float4), c->stream));
    return VN_OK;
}

int vn_reset_accum(vn_handle c) {
    VN_REQUIRE(c, c, "vn_reset_accum: NULL handle");
    VN_REQUIRE(c, c->accum, "vn_reset_accum: no accumulation buffer (call vn_resize)");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaMemsetAsync(c->accum, 0, (size_t)c->width * c->height * sizeof(float4), c->stream));
    return VN_OK;
}

int vn_set_accum_external(vn_handle c, void* dev_ptr) {
    VN_REQUIRE(c, c, "vn_set_accum_external: NULL handle");
    VN_REQUIRE(c, c->width && c->height, "vn_set_accum_external: call vn_resize first");
    c->accum = dev_ptr ? static_cast<float4*>(dev_ptr) : c->accum_own;
    return VN_OK;
}

int vn_accum_device_ptr(vn_handle c, void** dev_ptr) {
    VN_REQUIRE(c, c && dev_ptr, "vn_accum_device_ptr: NULL argument");
    VN_REQUIRE(c, c->accum, "vn_accum_device_ptr: no accumulation buffer (call vn_resize)");
    *dev_ptr = c->accum;
    return VN_OK;
}

int vn_read_accum(vn_handle c, float* host_rgba) {
    VN_REQUIRE(c, c && host_rgba, "vn_read_accum: NULL argument");
    VN_REQUIRE(c, c->accum, "vn_read_accum: no accumulation buffer");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaMemcpyAsync(host_rgba, c->accum, (size_t)c->width * c->height * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_write_accum(vn_handle c, const float* host_rgba) {
    VN_REQUIRE(c, c && host_rgba, "vn_write_accum: NULL argument");
    VN_REQUIRE(c, c->accum, "vn_write_accum: no accumulation buffer");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaMemcpyAsync(c->accum, host_rgba, (size_t)c->width * c->height * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

// Folds the counters of every finished launch into the statistics (the stream must have been synchronised): totals accumulate over all
// of them, the per-launch fields describe the newest one.
static int collect_render_stats(vn_context* c) {
    while (c->slots_pending) {
        const uint32_t slot = (c->slot_head + kStatSlots - c->slots_pending) % kStatSlots;
        const unsigned long long* hc = c->h_counters + 32u * slot;
        float ms = 0.0f;
        VN_CUDA(c, cudaEventElapsedTime(&ms, c->ev_slot[slot][0], c->ev_slot[slot][1]));
        c->stats.ms_render = ms;
        c->stats.ms_trace = ms;
        c->stats.segments = hc[0];
        c->stats.paths = hc[1];
        c->stats.node_visits = hc[2];
        c->stats.sphere_tests = hc[3];
        c->stats.segments_total += hc[0];
        c->stats.kernel_launches = c->slot_launches[slot];
        c->slots_pending -= 1u;
    }
    return VN_OK;
}

// kernels of the last vn_render done (statistics readable); pipelined frame copies may still be in flight
static int sync_kernels(vn_context* c) {
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    VN_CUDA(c, cudaGetLastError());
    return collect_render_stats(c);
}

int vn_synchronize(vn_handle c) {
    VN_REQUIRE(c, c, "vn_synchronize: NULL handle");
    VN_CUDA(c, cudaSetDevice(c->device));
    const int rc = sync_kernels(c);
    if (rc != VN_OK) return rc;
    VN_CUDA(c, cudaStreamSynchronize(c->copy_stream));     // VN_IMAGE_HOST | VN_ASYNC frames have landed in host memory
    return VN_OK;
}

static int ensure_image_pipe(vn_context* c, uint64_t pixels) {
    if (c->image_pipe_pixels >= pixels && c->image_pipe[0]) return VN_OK;
    VN_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    for (int i = 0; i < 2; i++) {
        cudaFree(c->image_pipe[i]);
        c->image_pipe[i] = nullptr;
        c->copied_valid[i] = false;
    }
    c->image_pipe_pixels = 0;
    for (int i = 0; i < 2; i++) VN_CUDA(c, cudaMalloc(&c->image_pipe[i], pixels * 4));
    c->image_pipe_pixels = pixels;
    return VN_OK;
}

static int ensure_image_tmp(vn_context* c, uint64_t pixels) {
    if (c->image_tmp_pixels >= pixels && c->image_tmp) return VN_OK;
    cudaFree(c->image_tmp);
    c->image_tmp = nullptr;
    c->image_tmp_pixels = 0;
    VN_CUDA(c, cudaMalloc(&c->image_tmp, pixels * 4));
    c->image_tmp_pixels = pixels;
    return VN_OK;
}

// per-(pixel, sample) radiance of one launch, summed in sample order by the accumulate kernel
static int ensure_sample_buffer(vn_context* c, uint64_t pixels, uint32_t spp) {
    WavefrontBuffers& w = c->wf;
    const uint64_t need = pixels * spp * 3ull;
    if (need > c->wf_sample_floats_) {
        cudaFree(w.sample_rgb);
        w.sample_rgb = nullptr;
        c->wf_sample_floats_ = 0;
        VN_CUDA(c, cudaMalloc(&w.sample_rgb, need * 4ull));
        c->wf_sample_floats_ = need;
    }
    return VN_OK;
}

static int ensure_wavefront(vn_context* c, uint64_t pixels, uint32_t spp) {
    WavefrontBuffers& w = c->wf;
    const uint32_t cap = c->wavefront_slots;
    if (w.capacity != cap) {
        free_wavefront(w);
        c->wf_sample_floats_ = 0;
        VN_CUDA(c, cudaMalloc(&w.slab, 4ull * cap * kWfArraysTotal));
        uint32_t* base = static_cast<uint32_t*>(w.slab);
        auto take = [&]() { uint32_t* p = base; base += cap; return p; };
        for (int q = 0; q < 2; q++) {
            WfState& s = w.st[q];
            float** fl[] = {&s.ox, &s.oy, &s.oz, &s.dx, &s.dy, &s.dz, &s.tr, &s.tg, &s.tb};
            for (float** p : fl) *p = reinterpret_cast<float*>(take());
            s.seed = take();
            s.ps = take();
            s.depth = reinterpret_cast<int32_t*>(take());
        }
        w.hit_t = reinterpret_cast<float*>(take());
        w.hit_prim = reinterpret_cast<int32_t*>(take());
        for (int i = 0; i < 4; i++) w.mat_queue[i] = take();
        VN_CUDA(c, cudaMalloc(&w.counts, 4 * kWfCountWords));
        w.capacity = cap;
    }
    return ensure_sample_buffer(c, pixels, spp);
}

// Longest-processing-time-first schedule for the persistent path kernels.  Lanes take 8x4-pixel tiles from a global ticket; with
// row-major tickets a launch ends with ~0.7 ms (11 % of a 1080p launch, measured with tools/tail_probe.py) in which the tickets
// are gone and ever fewer lanes finish the expensive pixels (glass: paths of up to max_depth segments) they took late.  Progressive
// rendering launches the same view again and again (Renderer::Draw, Renderer.h:35-78), so the first launch of a view counts every
// tile's ray segments, the tiles are sorted by that cost (hand-written radix sort, radix_sort.cuh) and later launches hand them out
// most expensive first: the drain then consists of the cheapest pixels.  Pixels are independent, so the image does not change.
static int prepare_tile_order(vn_handle c, const vn_params* p, RenderLaunch& L) {
    L.tile_order = nullptr;
    L.tile_cost = nullptr;
    const uint32_t n_tiles = L.total_work / 32u;
    if (!c->tile_order_opt || n_tiles < (uint32_t)c->num_sms * 32u) return VN_OK;      // small frames: every lane gets at most one tile anyway
    vn_context::TileSig sig;
    memset(&sig, 0, sizeof sig);          // the struct is compared with memcmp: padding bytes must be defined
    sig.w = p->width; sig.h = p->height; sig.r0 = L.row_begin; sig.r1 = L.row_end; sig.spp = p->samples_per_pixel; sig.depth = p->max_depth;
    const float cam[13] = {p->origin[0], p->origin[1], p->origin[2], p->u[0], p->u[1], p->u[2], p->v[0], p->v[1], p->v[2], p->w[0], p->w[1], p->w[2], p->lens_radius};
    memcpy(sig.cam, cam, sizeof cam);
    sig.epoch = c->bvh_epoch;
    if (n_tiles > c->tile_cap) {
        cudaFree(c->d_tile_cost); cudaFree(c->d_tile_sort);
        c->d_tile_cost = nullptr; c->d_tile_sort = nullptr; c->tile_cap = 0; c->tile_state = 0;
        VN_CUDA(c, cudaMalloc(&c->d_tile_cost, (size_t)n_tiles * 8));
        VN_CUDA(c, cudaMalloc(&c->d_tile_sort, (size_t)n_tiles * 16));
        c->tile_cap = n_tiles;
    }
    if (memcmp(&sig, &c->tile_sig, sizeof sig) != 0) {
        // a new view.  If only the camera (or the scene) changed, the previous view's order is the best guess for the collecting launch
        const bool same_frame = c->tile_state == 2 && sig.w == c->tile_sig.w && sig.h == c->tile_sig.h && sig.r0 == c->tile_sig.r0 && sig.r1 == c->tile_sig.r1;
        memcpy(&c->tile_sig, &sig, sizeof sig);
        c->tile_state = 0;
        c->tile_seg_per_path = 0.0;
        c->tile_guess = same_frame && c->tile_guess_opt;
    }
    if (c->tile_state == 0) {
        VN_CUDA(c, cudaMemsetAsync(c->d_tile_cost, 0, (size_t)c->tile_cap * 8, c->stream));
        L.tile_cost = c->d_tile_cost;
        L.tile_cost_stride = c->tile_cap;
        if (c->tile_guess) {
            L.tile_order = c->d_tile_order;                     // (stays valid until the sort of the next launch, which runs behind this one)
        }
        c->tile_guess = false;
        c->tile_state = 1;
        return VN_OK;
    }
    if (c->tile_state == 1) {
        uint32_t *k0 = c->d_tile_sort, *v0 = k0 + c->tile_cap, *k1 = v0 + c->tile_cap, *v1 = k1 + c->tile_cap;
        uint32_t* d_all_miss = reinterpret_cast<uint32_t*>(c->d_counters + 6);     // (a free word of the counter block; the launch's memset comes later)
        VN_CUDA(c, cudaMemsetAsync(d_all_miss, 0, 16, c->stream));
        VN_CUDA(c, exact::launch_tile_keys(c->d_tile_cost, c->tile_cap, n_tiles, c->tile_order_opt, p->samples_per_pixel, k0, v0, d_all_miss, c->stream));
        std::string err;
        uint32_t launches = 0;
        const int which = radix_sort_pairs_device(k0, v0, k1, v1, n_tiles, 24, c->num_sms, c->stream, &launches, err);
        if (which < 0) return fail(c, VN_ERR_CUDA, "vn_render: tile order sort failed: " + err);
        c->d_tile_order = which ? v1 : v0;
        // once per view: how many tiles saw nothing but sky in the collecting launch (they are the end of the order)
        // ... and how long its paths are (sample-range units pay when every pixel is expensive, see vn_render)
        unsigned long long back[2] = {0ull, 0ull};
        VN_CUDA(c, cudaMemcpyAsync(back, d_all_miss, 16, cudaMemcpyDeviceToHost, c->stream));
        VN_CUDA(c, cudaStreamSynchronize(c->stream));
        c->tile_all_miss = (uint32_t)(back[0] & 0xFFFFFFFFull);
        c->tile_seg_per_path = (double)back[1] / ((double)p->width * (double)(L.row_end - L.row_begin) * (double)p->samples_per_pixel);
        c->tile_state = 2;
    }
    L.tile_order = c->d_tile_order;
    return VN_OK;
}

// One call of the path kernel(s): subframe p->subframe_index, or -- when the caller has `want_n` > 1 subframes to render and the launch
// qualifies (see n_sub below) -- all of them in one launch.  *took = subframes rendered; the image is produced by the call that renders the last.
static int render_some(vn_handle c, const vn_params* p, uint32_t want_n, uint32_t stride, uint32_t* took) {
    *took = 1u;
    VN_REQUIRE(c, c && p, "vn_render: NULL argument");
    VN_REQUIRE(c, c->bvh_valid, "vn_render: no BVH (call vn_set_spheres + vn_build_bvh; Renderer::Init does both)");
    VN_REQUIRE(c, p->width >= 2 && p->height >= 2, "vn_render: width and height must be >= 2");
    VN_REQUIRE(c, p->samples_per_pixel >= 1, "vn_render: samples_per_pixel must be >= 1");
    VN_REQUIRE(c, p->max_depth >= 1, "vn_render: max_depth must be >= 1");
    VN_REQUIRE(c, p->row_begin <= p->row_end && p->row_end <= p->height, "vn_render: bad row range");
    VN_CUDA(c, cudaSetDevice(c->device));
    if (p->width != c->width || p->height != c->height || !c->accum) {
        const int rc = vn_resize(c, p->width, p->height);
        if (rc != VN_OK) return rc;
    }
    const uint64_t pixels = (uint64_t)p->width * p->height;
    const bool want_image = p->image && !(p->flags & VN_NO_TONEMAP) && !(p->flags & VN_ACCUM_SUM);
    const bool host_image = want_image && (p->flags & VN_IMAGE_HOST);
    RenderLaunch L;
    fill_launch(c, p, L);
    if (want_image) {
        if (host_image && (p->flags & VN_ASYNC)) { const int rc = ensure_image_pipe(c, pixels); if (rc != VN_OK) return rc; L.image = c->image_pipe[c->pipe_flip]; }
        else if (host_image) { const int rc = ensure_image_tmp(c, pixels); if (rc != VN_OK) return rc; L.image = c->image_tmp; }
        else L.image = static_cast<uint32_t*>(p->image);
    }
    const bool exact_build = !(p->flags & VN_FAST);
    const bool count = (p->flags & VN_COUNTERS) != 0;
    uint32_t launches = 0;
    if (count) { L.timeline = c->d_timeline; VN_CUDA(c, cudaMemsetAsync(c->d_timeline, 0, 16384, c->stream)); }

    const bool pipelined = host_image && (p->flags & VN_ASYNC);
    // a launch in flight owns a slot of the pinned counter ring; only a full ring makes the host wait (VN_ASYNC renders queue back to back)
    if (c->slots_pending >= kStatSlots) { const int rc = sync_kernels(c); if (rc != VN_OK) return rc; }
    const uint32_t slot = c->slot_head;
    // the staging buffer of this frame was last read by the copy of two frames ago
    if (pipelined && c->copied_valid[c->pipe_flip]) VN_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[c->pipe_flip], 0));
    VN_CUDA(c, cudaMemsetAsync(c->d_counters, 0, 256, c->stream));
    VN_CUDA(c, cudaEventRecord(c->ev_slot[slot][0], c->stream));
    bool emit = want_n == 1u;           // this call renders the last of the caller's subframes: tonemap, frame copy, and the host waits unless VN_ASYNC
    if (p->flags & VN_WAVEFRONT) {
        const uint32_t rows = L.row_end - L.row_begin;
        const int rc = ensure_wavefront(c, (uint64_t)p->width * rows, p->samples_per_pixel);
        if (rc != VN_OK) return rc;
        // the wide nodes of the path kernels when their eight octant copies fit in shared memory ("wavefront_wide"), else the pair nodes
        if (!(c->wavefront_wide && c->wide_nodes && L.wide && L.num_wide > 0 && c->scene.wide_levels <= kWideMaxLevels &&
              wide_smem_bytes(L.num_wide, L.num_spheres) + 2048 <= c->smem_optin)) { L.wide = nullptr; L.num_wide = 0; }
        c->last_accel = L.wide ? 2u : 1u;
        VN_CUDA(c, exact_build ? exact::launch_wavefront(L, c->wf, c->num_sms, c->stream, &launches)
                               : fast::launch_wavefront(L, c->wf, c->num_sms, c->stream, &launches));
    } else {
        { const int rc = prepare_tile_order(c, p, L); if (rc != VN_OK) return rc; }
        KernelConfig cfg;
        cfg.threads = c->threads;
        cfg.count = count;
        cfg.scene_in_smem = scene_fits_smem(c);
        cfg.smem_bytes = cfg.scene_in_smem ? scene_smem_bytes(L.num_nodes, L.num_spheres) : 0;
        // 8 octant-specialised copies of the nodes when they fit beside the spheres: one 1024-thread CTA per SM shares them
        const size_t oct_bytes = scene_smem_bytes(L.num_nodes, L.num_spheres, 8);
        cfg.octant = cfg.scene_in_smem && c->octant_nodes && oct_bytes + 2048 <= c->smem_optin;
        // preferred: 4-wide nodes, octant-sorted, 8 copies in shared memory (half the traversal steps, no distance compare)
        const size_t wide_bytes = wide_smem_bytes(L.num_wide, L.num_spheres);
        cfg.wide = c->wide_nodes && L.wide && L.num_wide > 0 && c->scene.wide_levels <= kWideMaxLevels && wide_bytes + 2048 <= c->smem_optin;
        // scenes too large for shared memory: the canonical wide nodes straight from L2/HBM (half the dependent fetches)
        const bool wide_global = !cfg.wide && !cfg.scene_in_smem && c->wide_global && c->wide_nodes && L.wide && L.num_wide > 0 && c->scene.wide_levels <= kWideGlobalMaxLevels;
        // small scenes of similar-sized spheres: uniform grid + oversize list (grid_core.cuh), 40 % of the BVH's instructions on RTIOW
        cfg.grid = c->accel != 1u && c->grid.valid && grid_smem_bytes(c->grid.h.n_cells, c->grid.h.n_refs, L.num_spheres) + 2048 <= c->smem_optin;
        if (cfg.grid) { cfg.scene_in_smem = true; cfg.octant = false; cfg.wide = false; cfg.smem_bytes = grid_smem_bytes(c->grid.h.n_cells, c->grid.h.n_refs, L.num_spheres); cfg.threads = c->wide_threads; }
        else if (cfg.wide) { cfg.scene_in_smem = true; cfg.octant = false; cfg.smem_bytes = wide_bytes; cfg.threads = c->wide_threads; cfg.async = c->async_done > 0u; cfg.warp_tiles = cfg.async && c->async_node == 0u && c->warp_tiles != 0u;
            cfg.lean = cfg.warp_tiles && c->lean != 0u && cfg.threads >= 768 && L.num_spheres < kLink16MaxPrims && L.num_wide < 32768u && p->samples_per_pixel < 65536u &&
                       p->max_depth < 65536u && p->width < 65536u && p->height < 65536u; }
        else if (wide_global) { cfg.wide = true; cfg.octant = false; if (cfg.threads > 256) cfg.threads = 256; }
        else if (cfg.octant) { cfg.smem_bytes = oct_bytes; cfg.threads = 1024; }
        else if (cfg.threads > 256) cfg.threads = 256;
        // scenes traversed from L2 / HBM (pair nodes): the asynchronous form of the path kernel (k_render_lean<kGlobal>, 256-thread CTAs)
        if (!cfg.scene_in_smem && !cfg.wide && !cfg.grid && c->lean != 0u && c->async_done > 0u && p->width < 65536u && p->height < 65536u) { cfg.lean = true; cfg.threads = 256; cfg.global_ctas = c->global_ctas; L.async_done = c->global_done; }
        int per_sm = (cfg.octant || (cfg.wide && cfg.scene_in_smem)) ? 1 : c->blocks_per_sm;
        if (per_sm <= 0) {
            per_sm = exact_build ? exact::max_blocks_per_sm(cfg.threads, cfg.smem_bytes, cfg.scene_in_smem, count, cfg.octant, cfg.wide, cfg.grid, cfg.lean, cfg.global_ctas)
                                 : fast::max_blocks_per_sm(cfg.threads, cfg.smem_bytes, cfg.scene_in_smem, count, cfg.octant, cfg.wide, cfg.grid, cfg.lean, cfg.global_ctas);
            if (per_sm <= 0) return fail(c, VN_ERR_CUDA, "vn_render: occupancy query failed for the path kernel");
        }
        cfg.blocks = c->num_sms * per_sm;
        c->last_accel = cfg.grid ? 4u : (cfg.wide ? (cfg.scene_in_smem ? 2u : 3u) : 1u);
        // never launch more lanes than there is work
        const uint64_t max_blocks = ((uint64_t)L.total_work + cfg.threads - 1) / cfg.threads;
        if ((uint64_t)cfg.blocks > max_blocks) cfg.blocks = (int)std::max<uint64_t>(1, max_blocks);
        L.units_log2 = 0u; L.carry = nullptr; L.unit_epoch = 0u;
        {
            // sample-range units for the L2 / HBM form of k_render_lean: every tile's samples in 2 or 4 ranges, all first ranges first
            uint32_t lu = c->units >= 16u ? 4u : (c->units >= 8u ? 3u : (c->units >= 4u ? 2u : (c->units >= 2u ? 1u : 0u)));
            while (lu > 0u && (p->samples_per_pixel % (1u << lu)) != 0u) lu -= 1u;
            const uint32_t n_tiles_u = L.total_work / 32u;
            if (c->units_min_seg > 0.0 && !(c->tile_state == 2 && c->tile_seg_per_path >= c->units_min_seg)) lu = 0u;
            if (lu > 0u && cfg.lean && !cfg.scene_in_smem && !L.tile_cost && p->width < 16384u && p->height < 16384u && n_tiles_u < (1u << 22) &&
                n_tiles_u >= (uint32_t)c->num_sms * 8u) {
                if (n_tiles_u > c->unit_tiles_cap) {
                    VN_CUDA(c, cudaStreamSynchronize(c->stream));
                    cudaFree(c->d_unit_items); cudaFree(c->d_unit_carry);
                    c->d_unit_items = nullptr; c->d_unit_carry = nullptr; c->unit_tiles_cap = 0;
                    VN_CUDA(c, cudaMalloc(&c->d_unit_items, (size_t)n_tiles_u * 16 * sizeof(uint32_t)));
                    VN_CUDA(c, cudaMalloc(&c->d_unit_carry, (size_t)n_tiles_u * 32 * 2 * sizeof(float4)));
                    VN_CUDA(c, cudaMemsetAsync(c->d_unit_carry, 0, (size_t)n_tiles_u * 32 * 2 * sizeof(float4), c->stream));
                    c->unit_tiles_cap = n_tiles_u;
                    c->unit_epoch = 0;
                }
                c->unit_epoch += 32u;
                if (c->unit_epoch > 0x7FFFFF00u) {               // (once in 2^28 launches: start over)
                    VN_CUDA(c, cudaMemsetAsync(c->d_unit_carry, 0, c->unit_tiles_cap * 32 * 2 * sizeof(float4), c->stream));
                    c->unit_epoch = 32u;
                }
                VN_CUDA(c, exact::launch_unit_items(L.tile_order, n_tiles_u, lu, c->d_unit_items, c->stream));
                L.tile_order = c->d_unit_items;
                L.total_work = (n_tiles_u << lu) * 32u;
                L.units_log2 = lu; L.carry = c->d_unit_carry; L.unit_epoch = c->unit_epoch;
                launches += 1;
            }
        }
        L.steal_scratch = nullptr; L.steal_count = nullptr;
        if (cfg.lean && c->steal != 0u && (!cfg.scene_in_smem || c->steal_smem != 0u) && p->samples_per_pixel > 1u && p->samples_per_pixel < 1024u) {
            // scratch slots of the drain's sample stealing: one per lane of the grid (a few tens of MB; 180 GB of HBM)
            const size_t lanes = (size_t)cfg.blocks * cfg.threads;
            const size_t need = lanes * (p->samples_per_pixel + 1u) * sizeof(float4);
            if (need > ((size_t)1 << 30)) {
                // (hundreds of samples per pixel and launch: the slots would take gigabytes, and launches that long do not need the drain shortened)
            } else {
            if (need > c->steal_scratch_cap) {
                VN_CUDA(c, cudaStreamSynchronize(c->stream));
                cudaFree(c->d_steal_scratch); c->d_steal_scratch = nullptr; c->steal_scratch_cap = 0;
                VN_CUDA(c, cudaMalloc(&c->d_steal_scratch, need));
                c->steal_scratch_cap = need;
            }
            if (lanes * sizeof(uint32_t) > c->steal_count_cap) {
                VN_CUDA(c, cudaStreamSynchronize(c->stream));
                cudaFree(c->d_steal_count); c->d_steal_count = nullptr; c->steal_count_cap = 0;
                VN_CUDA(c, cudaMalloc(&c->d_steal_count, lanes * sizeof(uint32_t)));
                c->steal_count_cap = lanes * sizeof(uint32_t);
                VN_CUDA(c, cudaMemsetAsync(c->d_steal_count, 0, c->steal_count_cap, c->stream));
            }
            L.steal_scratch = c->d_steal_scratch; L.steal_count = c->d_steal_count;
            }
        }
        // several subframes in one launch (k_render_lean<kMulti>, path_kernels.cu::finish_pixel_multi): the shared-memory form of the lean kernel,
        // a view whose tile costs are known (or a frame too small for a tile order), no instrumentation, no stealing drain
        uint32_t n_sub = 1u;
        if (want_n > 1u && c->multi_subframes > 1u && cfg.lean && cfg.scene_in_smem && cfg.wide && !cfg.grid && !L.tile_cost && !count &&
            p->width < 8192u && p->height < 8192u) {
            n_sub = std::min(std::min(want_n, 64u), c->multi_subframes);
            const uint64_t cap = 0xFFFFFFFFull / std::max<uint64_t>(1u, L.total_work);
            if ((uint64_t)n_sub > cap) n_sub = (uint32_t)std::max<uint64_t>(1u, cap);
        }
        if (n_sub > 1u) {
            L.n_sub = n_sub; L.tiles_per_sub = L.total_work / 32u; L.total_work *= n_sub; L.sub_stride = stride;
            L.steal_scratch = nullptr; L.steal_count = nullptr;
            *took = n_sub;
        }
        emit = n_sub == want_n;
        // split frame: the cheap end of the cost-ordered tile list goes to a second launch on tail_stream (same kernel, own ticket counter, the
        // statistics add up in the same counters); everything behind it on c->stream waits for both
        const uint32_t n_tiles = L.total_work / 32u;
        // ... and only tiles that saw nothing but sky when the costs were collected: uniform, cheap, no pixel that bounces for a millisecond --
        // a second launch that holds heavy-tailed tiles ends later than the first (measured: +0.2 ms as soon as it reached beyond the sky)
        const uint32_t n_tail = (L.tile_order && !L.tile_cost && !count && cfg.lean && cfg.scene_in_smem && c->split_tail > 0.0f && n_sub == 1u)
                                    ? std::min(c->tile_all_miss, (uint32_t)((double)n_tiles * c->split_tail)) : 0u;
        if (n_tail > 0u && n_tail < n_tiles) {
            RenderLaunch T = L;
            T.tile_order = L.tile_order + (n_tiles - n_tail);
            T.total_work = n_tail * 32u;
            T.work_counter = reinterpret_cast<uint32_t*>(c->d_counters + 5);
            L.total_work = (n_tiles - n_tail) * 32u;
            static const bool debug_split = getenv("VN_DEBUG_SPLIT") != nullptr;      // prints the two launches' end times (host sync: measurements only)
            cudaEvent_t dbg[3] = {nullptr, nullptr, nullptr};
            if (debug_split) { for (auto& e : dbg) cudaEventCreate(&e); cudaEventRecord(dbg[0], c->stream); }
            VN_CUDA(c, cudaEventRecord(c->ev_tail[0], c->stream));
            VN_CUDA(c, exact_build ? exact::launch_render_persistent(L, cfg, c->stream) : fast::launch_render_persistent(L, cfg, c->stream));
            if (debug_split) cudaEventRecord(dbg[1], c->stream);
            VN_CUDA(c, cudaStreamWaitEvent(c->tail_stream, c->ev_tail[0], 0));
            VN_CUDA(c, exact_build ? exact::launch_render_persistent(T, cfg, c->tail_stream) : fast::launch_render_persistent(T, cfg, c->tail_stream));
            if (debug_split) cudaEventRecord(dbg[2], c->tail_stream);
            VN_CUDA(c, cudaEventRecord(c->ev_tail[1], c->tail_stream));
            VN_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tail[1], 0));
            if (debug_split) {
                cudaEventSynchronize(dbg[1]); cudaEventSynchronize(dbg[2]);
                float a = 0.0f, b = 0.0f;
                cudaEventElapsedTime(&a, dbg[0], dbg[1]); cudaEventElapsedTime(&b, dbg[0], dbg[2]);
                fprintf(stderr, "split frame: first launch (%u tiles) ends at %.3f ms, second (%u tiles) at %.3f ms\n", n_tiles - n_tail, a, n_tail, b);
                for (auto& e : dbg) cudaEventDestroy(e);
            }
            L.total_work = n_tiles * 32u;
            launches += 1;
        } else {
            VN_CUDA(c, exact_build ? exact::launch_render_persistent(L, cfg, c->stream) : fast::launch_render_persistent(L, cfg, c->stream));
        }
        launches += 1;
        if (L.image && emit) {
            // sRGB + quantise of the rows just rendered (RayTracer.cu:216), as a coalesced kernel behind the path kernel
            const uint64_t begin = (uint64_t)L.row_begin * L.width, count = (uint64_t)(L.row_end - L.row_begin) * L.width;
            VN_CUDA(c, exact_build ? exact::launch_tonemap(L.accum + begin, 1.0f, L.image + begin, count, c->stream)
                                   : fast::launch_tonemap(L.accum + begin, 1.0f, L.image + begin, count, c->stream));
            launches += 1;
        }
    }
    VN_CUDA(c, cudaEventRecord(c->ev_slot[slot][1], c->stream));
    // only the rows this launch rendered were tonemapped into the staging buffer: copy those and leave the caller's other rows alone
    const uint64_t row_px0 = (uint64_t)L.row_begin * L.width, row_px = (uint64_t)(L.row_end - L.row_begin) * L.width;
    // (the 256-byte statistics copy goes first: queued behind the frame on the D2H engine it would delay the next launch)
    VN_CUDA(c, cudaMemcpyAsync(c->h_counters + 32u * slot, c->d_counters, 256, cudaMemcpyDeviceToHost, c->stream));
    if (pipelined && emit) {
        const int f = c->pipe_flip;
        VN_CUDA(c, cudaEventRecord(c->ev_frame[f], c->stream));
        VN_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_frame[f], 0));
        VN_CUDA(c, cudaMemcpyAsync(static_cast<uint32_t*>(p->image) + row_px0, c->image_pipe[f] + row_px0, row_px * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        VN_CUDA(c, cudaEventRecord(c->ev_copied[f], c->copy_stream));
        c->copied_valid[f] = true;
        c->pipe_flip = f ^ 1;
    } else if (host_image && emit) {
        VN_CUDA(c, cudaMemcpyAsync(static_cast<uint32_t*>(p->image) + row_px0, c->image_tmp + row_px0, row_px * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    c->slot_launches[slot] = launches;
    c->stats.kernel_launches = launches;
    c->stats.kernel_launches_total += launches;
    c->slot_head = (slot + 1u) % kStatSlots;
    c->slots_pending += 1u;
    if (!(p->flags & VN_ASYNC) && emit) return sync_kernels(c);
    return VN_OK;
}

int vn_render(vn_handle c, const vn_params* p) {
    uint32_t took = 0;
    return render_some(c, p, 1u, 1u, &took);
}

int vn_render_subframes_strided(vn_handle c, const vn_params* p, uint32_t n, uint32_t stride) {
    VN_REQUIRE(c, c && p, "vn_render_subframes: NULL argument");
    VN_REQUIRE(c, n >= 1u && n <= 65536u, "vn_render_subframes: n must be in [1,65536]");
    VN_REQUIRE(c, stride >= 1u && stride <= 65536u, "vn_render_subframes: stride must be in [1,65536]");
    vn_params q = *p;
    for (uint32_t done = 0; done < n;) {
        q.subframe_index = p->subframe_index + done * stride;
        q.accum_count = p->accum_count + ((p->flags & VN_ACCUM_SUM) ? 0u : done);
        uint32_t took = 0;
        const int rc = render_some(c, &q, n - done, stride, &took);
        if (rc != VN_OK) return rc;
        done += took;
    }
    return VN_OK;
}

int vn_render_subframes(vn_handle c, const vn_params* p, uint32_t n) { return vn_render_subframes_strided(c, p, n, 1u); }

int vn_tonemap(vn_handle c, float scale, void* image, uint32_t flags) {
    VN_REQUIRE(c, c && image, "vn_tonemap: NULL argument");
    VN_REQUIRE(c, c->accum, "vn_tonemap: no accumulation buffer");
    VN_CUDA(c, cudaSetDevice(c->device));
    const uint64_t pixels = (uint64_t)c->width * c->height;
    uint32_t* dst = static_cast<uint32_t*>(image);
    if (flags & VN_IMAGE_HOST) { const int rc = ensure_image_tmp(c, pixels); if (rc != VN_OK) return rc; dst = c->image_tmp; }
    VN_CUDA(c, !(flags & VN_FAST) ? exact::launch_tonemap(c->accum, scale, dst, pixels, c->stream) : fast::launch_tonemap(c->accum, scale, dst, pixels, c->stream));
    c->stats.kernel_launches_total += 1;
    if (flags & VN_IMAGE_HOST) VN_CUDA(c, cudaMemcpyAsync(image, c->image_tmp, pixels * 4, cudaMemcpyDeviceToHost, c->stream));
    if (!(flags & VN_ASYNC) || (flags & VN_IMAGE_HOST)) VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_reduce_tonemap_peers(vn_handle c, const void* const* peer_accum, uint32_t n_peers, float scale, uint32_t row_begin,
                            uint32_t row_end, void* image, uint32_t flags) {
    VN_REQUIRE(c, c, "vn_reduce_tonemap_peers: NULL argument");
    return vn_reduce_tonemap_peers_to(c, peer_accum, n_peers, scale, row_begin, row_end, c->accum, image, flags);
}

int vn_reduce_tonemap_peers_to(vn_handle c, const void* const* peer_accum, uint32_t n_peers, float scale, uint32_t row_begin,
                               uint32_t row_end, void* sum_out, void* image, uint32_t flags) {
    return vn_reduce_tonemap_peers_wait(c, peer_accum, n_peers, scale, row_begin, row_end, sum_out, image, nullptr, 0u, flags);
}

int vn_sync_flags(vn_handle c, void** dev_ptr) {
    VN_REQUIRE(c, c && dev_ptr, "vn_sync_flags: NULL argument");
    *dev_ptr = c->d_flags;
    return VN_OK;
}

int vn_signal(vn_handle c, uint32_t index, uint32_t value) {
    VN_REQUIRE(c, c, "vn_signal: NULL handle");
    VN_REQUIRE(c, index < 62u, "vn_signal: flag index must be < 62");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, exact::launch_signal(c->d_flags + index, value, c->stream));
    c->stats.kernel_launches_total += 1;
    return VN_OK;
}

int vn_wait_flags(vn_handle c, const void* const* flags, uint32_t n, uint32_t value) {
    VN_REQUIRE(c, c && flags, "vn_wait_flags: NULL argument");
    VN_REQUIRE(c, n >= 1 && n <= (uint32_t)kMaxPeers, "vn_wait_flags: 1..8 flags");
    VN_CUDA(c, cudaSetDevice(c->device));
    const uint32_t* f[kMaxPeers];
    for (uint32_t i = 0; i < n; i++) f[i] = static_cast<const uint32_t*>(flags[i]);
    VN_CUDA(c, exact::launch_wait_flags(f, n, value, c->d_flags + 63, c->stream));
    c->stats.kernel_launches_total += 1;
    return VN_OK;
}

int vn_check_flags(vn_handle c) {
    VN_REQUIRE(c, c, "vn_check_flags: NULL handle");
    VN_CUDA(c, cudaSetDevice(c->device));
    uint32_t err = 0;
    VN_CUDA(c, cudaMemcpyAsync(&err, c->d_flags + 63, 4, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    if (err) {
        VN_CUDA(c, cudaMemsetAsync(c->d_flags + 63, 0, 4, c->stream));
        return fail(c, VN_ERR_CUDA, "a device-side wait for peer " + std::to_string(err - 1u) + "'s epoch flag timed out (4 s): that rank never signalled");
    }
    return VN_OK;
}

int vn_reduce_tonemap_peers_wait(vn_handle c, const void* const* peer_accum, uint32_t n_peers, float scale, uint32_t row_begin,
                                 uint32_t row_end, void* sum_out, void* image, const void* const* peer_flags, uint32_t wait_value, uint32_t flags) {
    VN_REQUIRE(c, c && peer_accum, "vn_reduce_tonemap_peers: NULL argument");
    VN_REQUIRE(c, n_peers >= 1 && n_peers <= (uint32_t)kMaxPeers, "vn_reduce_tonemap_peers: 1..8 peers");
    VN_REQUIRE(c, c->accum, "vn_reduce_tonemap_peers: no accumulation buffer");
    VN_REQUIRE(c, row_begin <= row_end && row_end <= c->height, "vn_reduce_tonemap_peers: bad row range");
    VN_CUDA(c, cudaSetDevice(c->device));
    const float4* peers[kMaxPeers];
    for (uint32_t i = 0; i < n_peers; i++) peers[i] = static_cast<const float4*>(peer_accum[i]);
    const uint64_t begin = (uint64_t)row_begin * c->width, end = (uint64_t)row_end * c->width;
    uint32_t* img = static_cast<uint32_t*>(image);
    float4* sum = static_cast<float4*>(sum_out);
    const uint32_t* pf[kMaxPeers];
    for (uint32_t i = 0; peer_flags && i < n_peers; i++) pf[i] = static_cast<const uint32_t*>(peer_flags[i]);
    VN_CUDA(c, !(flags & VN_FAST) ? exact::launch_reduce_tonemap_peers(peers, n_peers, scale, begin, end, sum, img, peer_flags ? pf : nullptr, wait_value, c->d_flags + 63, c->stream)
                                  : fast::launch_reduce_tonemap_peers(peers, n_peers, scale, begin, end, sum, img, peer_flags ? pf : nullptr, wait_value, c->d_flags + 63, c->stream));
    c->stats.kernel_launches_total += 1;
    if (!(flags & VN_ASYNC)) VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_get_stats(vn_handle c, vn_stats* out) {
    VN_REQUIRE(c, c && out, "vn_get_stats: NULL argument");
    *out = c->stats;
    return VN_OK;
}

int vn_reset_stats(vn_handle c) {
    VN_REQUIRE(c, c, "vn_reset_stats: NULL handle");
    const float b = c->stats.ms_build, u = c->stats.ms_upload;
    c->stats = vn_stats{};
    c->stats.ms_build = b; c->stats.ms_upload = u;
    return VN_OK;
}

// ---- buffers / IPC
int vn_buffer_alloc(int device, uint64_t bytes, int zero_copy_host, void** dev_ptr, void** host_ptr) {
    if (!dev_ptr || !host_ptr) return fail(nullptr, VN_ERR_INVALID, "vn_buffer_alloc: NULL argument");
    *dev_ptr = *host_ptr = nullptr;
    VN_CUDA(nullptr, cudaSetDevice(device));
    if (zero_copy_host) {
        VN_CUDA(nullptr, cudaHostAlloc(host_ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        VN_CUDA(nullptr, cudaHostGetDevicePointer(dev_ptr, *host_ptr, 0));
    } else {
        VN_CUDA(nullptr, cudaMalloc(dev_ptr, bytes));
    }
    return VN_OK;
}

int vn_buffer_free(int device, void* dev_ptr, void* host_ptr, int zero_copy_host) {
    VN_CUDA(nullptr, cudaSetDevice(device));
    if (zero_copy_host) { if (host_ptr) VN_CUDA(nullptr, cudaFreeHost(host_ptr)); }
    else if (dev_ptr) VN_CUDA(nullptr, cudaFree(dev_ptr));
    return VN_OK;
}

int vn_buffer_copy_to_host(int device, void* host_dst, const void* dev_src, uint64_t bytes) {
    VN_CUDA(nullptr, cudaSetDevice(device));
    VN_CUDA(nullptr, cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
    return VN_OK;
}

int vn_stream_synchronize(int device, void* cuda_stream) {
    VN_CUDA(nullptr, cudaSetDevice(device));
    VN_CUDA(nullptr, cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    return VN_OK;
}

// A CUDA IPC handle names a whole ALLOCATION: for a pointer into the middle of one (a tensor carved out of a caching allocator's
// segment) cudaIpcGetMemHandle silently returns the handle of the segment, and the peer that opens it gets the segment's base.
// vn_ipc_export_at reports the pointer's offset inside its allocation (driver entry point cuMemGetAddressRange, looked up at run
// time: no link-time dependency on libcuda); vn_ipc_export refuses pointers that are not the base of theirs.
int vn_ipc_export_at(vn_handle c, void* dev_ptr, unsigned char handle_out[64], uint64_t* offset_out) {
    VN_REQUIRE(c, c && dev_ptr && handle_out && offset_out, "vn_ipc_export_at: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    VN_CUDA(c, cudaSetDevice(c->device));
    typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VN_CUDA(c, cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    VN_REQUIRE(c, fn && qres == cudaDriverEntryPointSuccess, "vn_ipc_export_at: cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    const int drc = reinterpret_cast<GetRange>(fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
    VN_REQUIRE(c, drc == 0 && base != 0, "vn_ipc_export_at: not a device allocation");
    cudaIpcMemHandle_t hd;
    VN_CUDA(c, cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>((uintptr_t)base)));
    memcpy(handle_out, &hd, 64);
    *offset_out = (uint64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
    return VN_OK;
}

int vn_ipc_export(vn_handle c, void* dev_ptr, unsigned char handle_out[64]) {
    uint64_t offset = 0;
    const int rc = vn_ipc_export_at(c, dev_ptr, handle_out, &offset);
    if (rc != VN_OK) return rc;
    VN_REQUIRE(c, offset == 0, "vn_ipc_export: the pointer is not the base of its allocation (an IPC handle names the whole allocation): use vn_ipc_export_at");
    return VN_OK;
}

int vn_ipc_open(vn_handle c, const unsigned char handle_in[64], void** dev_ptr) {
    VN_REQUIRE(c, c && handle_in && dev_ptr, "vn_ipc_open: NULL argument");
    VN_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle_in, 64);
    VN_CUDA(c, cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return VN_OK;
}

int vn_ipc_close(vn_handle c, void* dev_ptr) {
    VN_REQUIRE(c, c && dev_ptr, "vn_ipc_close: NULL argument");
    VN_CUDA(c, cudaSetDevice(c->device));
    VN_CUDA(c, cudaIpcCloseMemHandle(dev_ptr));
    return VN_OK;
}

// ---- host-side scene / camera helpers: thin C wrappers over the drop-in classes of include/venusaur/
uint32_t vn_scene_rtiow_final(vn_sphere* out, uint32_t cap) {
    venusaur::Scene scene;                                   // Scene.h:13-80
    const std::vector<vn_sphere> flat = scene.Flatten();
    if (out) for (uint32_t i = 0; i < flat.size() && i < cap; i++) out[i] = flat[i];
    return (uint32_t)flat.size();
}

void vn_scene_random(vn_sphere* out, uint64_t n, uint32_t seed, float S, uint32_t mix) {
    // sphere i is a pure function of (i, seed) through the reference's own tea<4>/lcg stream.
    for (uint64_t i = 0; i < n; i++) {
        uint32_t s = tea4((uint32_t)i, seed);
        vn_sphere o;
        const float rx = rnd(s); o.cx = S * (2.0f * rx - 1.0f);
        const float ry = rnd(s); o.cy = S * (2.0f * ry - 1.0f);
        const float rz = rnd(s); o.cz = S * (2.0f * rz - 1.0f);
        const float rr = rnd(s); o.r = 0.1f + 0.2f * rr;
        const float m = rnd(s);
        if (mix == 0) o.type = m < 0.80f ? VN_LAMBERTIAN : (m < 0.95f ? VN_METAL : VN_DIELECTRIC);
        else          o.type = m < 0.50f ? VN_DIELECTRIC : (m < 0.90f ? VN_LAMBERTIAN : VN_METAL);
        if (o.type == VN_LAMBERTIAN) {
            const float a0 = rnd(s), a1 = rnd(s), b0 = rnd(s), b1 = rnd(s), c0 = rnd(s), c1 = rnd(s);
            o.ax = a0 * a1; o.ay = b0 * b1; o.az = c0 * c1; o.fuzz_or_ir = 0.0f;
        } else if (o.type == VN_METAL) {
            const float a0 = rnd(s), a1 = rnd(s), a2 = rnd(s), f = rnd(s);
            o.ax = 0.5f + 0.5f * a0; o.ay = 0.5f + 0.5f * a1; o.az = 0.5f + 0.5f * a2; o.fuzz_or_ir = 0.5f * f;
        } else {
            o.ax = o.ay = o.az = 0.0f; o.fuzz_or_ir = 1.5f;
        }
        out[i] = o;
    }
}

void vn_camera_frame(const float lookfrom[3], const float forward[3], float vfov_deg, float aspect, float aperture, float focal_length,
                     float origin[3], float u[3], float v[3], float w[3], float* lens_radius) {
    venusaur::Camera cam(venusaur::vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vfov_deg, aspect, aperture, focal_length);   // Core.cpp:29
    cam.SetForward(venusaur::vec3(forward[0], forward[1], forward[2]));                                                     // Core.cpp:355
    venusaur::vec3 U, V, W;
    cam.UVWFrame(U, V, W);
    origin[0] = cam.GetPosition().x; origin[1] = cam.GetPosition().y; origin[2] = cam.GetPosition().z;
    u[0] = U.x; u[1] = U.y; u[2] = U.z;
    v[0] = V.x; v[1] = V.y; v[2] = V.z;
    w[0] = W.x; w[1] = W.y; w[2] = W.z;
    *lens_radius = cam.GetLensRadius();
}

// ---- unit-level test entry points

int vn_test_rng(vn_handle c, const uint32_t* v0, const uint32_t* v1, uint64_t n, uint32_t n_draws, uint32_t* seeds_out,
                uint32_t* lcg_out, float* rnd_out) {
    VN_REQUIRE(c, c && v0 && v1 && seeds_out && lcg_out && rnd_out, "vn_test_rng: NULL argument");
    VN_CUDA(c, cudaSetDevice(c->device));
    DevBuf a, b, s, l, r;
    VN_CUDA(c, a.alloc(4 * n)); VN_CUDA(c, b.alloc(4 * n)); VN_CUDA(c, s.alloc(4 * n));
    VN_CUDA(c, l.alloc(4 * n * n_draws)); VN_CUDA(c, r.alloc(4 * n * n_draws));
    VN_CUDA(c, cudaMemcpyAsync(a.p, v0, 4 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(b.p, v1, 4 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, fast::launch_test_rng(a.as<uint32_t>(), b.as<uint32_t>(), n, n_draws, s.as<uint32_t>(), l.as<uint32_t>(), r.as<float>(), c->stream));
    VN_CUDA(c, cudaMemcpyAsync(seeds_out, s.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(lcg_out, l.p, 4 * n * n_draws, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(rnd_out, r.p, 4 * n * n_draws, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_trace_rays(vn_handle c, const float* origins, const float* dirs, uint64_t n, float* t_out, int32_t* prim_out, uint32_t flags) {
    VN_REQUIRE(c, c && origins && dirs && t_out && prim_out, "vn_trace_rays: NULL argument");
    VN_REQUIRE(c, c->bvh_valid, "vn_trace_rays: no BVH");
    VN_CUDA(c, cudaSetDevice(c->device));
    DevBuf o, d, t, pr;
    VN_CUDA(c, o.alloc(12 * n)); VN_CUDA(c, d.alloc(12 * n)); VN_CUDA(c, t.alloc(4 * n)); VN_CUDA(c, pr.alloc(4 * n));
    VN_CUDA(c, cudaMemcpyAsync(o.p, origins, 12 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(d.p, dirs, 12 * n, cudaMemcpyHostToDevice, c->stream));
    RenderLaunch L;
    memset(&L, 0, sizeof(L));
    L.nodes = c->scene.nodes; L.geom = c->scene.geom; L.root_link = c->scene.root_link;
    L.gate = (c->hit_gate == 2u || (c->hit_gate == 1u && !scene_fits_smem(c))) ? 1u : 0u;
    const bool use_grid = (flags & VN_GRID) != 0;
    VN_REQUIRE(c, !use_grid || c->grid.valid, "vn_trace_rays: VN_GRID but the scene has no grid");
    L.grid = c->grid.h; L.grid_start = c->grid.start; L.grid_refs = c->grid.refs;
    VN_CUDA(c, !(flags & VN_FAST) ? exact::launch_trace_rays(L, o.as<float>(), d.as<float>(), n, t.as<float>(), pr.as<int32_t>(), c->scene.orig, use_grid, c->stream)
                                  : fast::launch_trace_rays(L, o.as<float>(), d.as<float>(), n, t.as<float>(), pr.as<int32_t>(), c->scene.orig, use_grid, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(t_out, t.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(prim_out, pr.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_sort_pairs(vn_handle c, uint32_t* keys, uint32_t* values, uint64_t n, uint32_t key_bits) {
    VN_REQUIRE(c, c && (n == 0 || (keys && values)), "vn_sort_pairs: NULL argument");
    VN_REQUIRE(c, key_bits >= 1 && key_bits <= 32, "vn_sort_pairs: key_bits must be in [1,32]");
    VN_REQUIRE(c, n < (1ull << 30), "vn_sort_pairs: n must be < 2^30");
    if (n == 0) return VN_OK;
    VN_CUDA(c, cudaSetDevice(c->device));
    DevBuf k0, v0, k1, v1;
    VN_CUDA(c, k0.alloc(4 * n)); VN_CUDA(c, v0.alloc(4 * n)); VN_CUDA(c, k1.alloc(4 * n)); VN_CUDA(c, v1.alloc(4 * n));
    VN_CUDA(c, cudaMemcpyAsync(k0.p, keys, 4 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(v0.p, values, 4 * n, cudaMemcpyHostToDevice, c->stream));
    std::string err;
    uint32_t launches = 0;
    const int which = radix_sort_pairs_device(k0.as<uint32_t>(), v0.as<uint32_t>(), k1.as<uint32_t>(), v1.as<uint32_t>(), (uint32_t)n, (int)key_bits,
                                              c->num_sms, c->stream, &launches, err);
    if (which < 0) return fail(c, VN_ERR_CUDA, "vn_sort_pairs: " + err);
    c->stats.kernel_launches_total += launches;
    VN_CUDA(c, cudaMemcpyAsync(keys, which ? k1.p : k0.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(values, which ? v1.p : v0.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_test_make_color(vn_handle c, const float* rgb, uint64_t n, uint8_t* rgba_out, uint32_t flags) {
    VN_REQUIRE(c, c && rgb && rgba_out, "vn_test_make_color: NULL argument");
    VN_CUDA(c, cudaSetDevice(c->device));
    DevBuf in, out;
    VN_CUDA(c, in.alloc(12 * n)); VN_CUDA(c, out.alloc(4 * n));
    VN_CUDA(c, cudaMemcpyAsync(in.p, rgb, 12 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, !(flags & VN_FAST) ? exact::launch_make_color(in.as<float>(), n, out.as<uint32_t>(), c->stream)
                                  : fast::launch_make_color(in.as<float>(), n, out.as<uint32_t>(), c->stream));
    VN_CUDA(c, cudaMemcpyAsync(rgba_out, out.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

int vn_test_scatter(vn_handle c, uint32_t material_type, const float albedo_fuzz_ir[4], const float* dirs, const float* normals,
                    const uint8_t* front, const uint32_t* seeds, uint64_t n, float* dirs_out, uint8_t* scattered_out,
                    uint32_t* seeds_out, uint32_t flags) {
    VN_REQUIRE(c, c && albedo_fuzz_ir && dirs && normals && front && seeds && dirs_out && scattered_out && seeds_out, "vn_test_scatter: NULL argument");
    VN_REQUIRE(c, material_type <= 2u, "vn_test_scatter: bad material type");
    VN_CUDA(c, cudaSetDevice(c->device));
    DevBuf d, nr, fr, sd, dout, sc, sout;
    VN_CUDA(c, d.alloc(12 * n)); VN_CUDA(c, nr.alloc(12 * n)); VN_CUDA(c, fr.alloc(n)); VN_CUDA(c, sd.alloc(4 * n));
    VN_CUDA(c, dout.alloc(12 * n)); VN_CUDA(c, sc.alloc(n)); VN_CUDA(c, sout.alloc(4 * n));
    VN_CUDA(c, cudaMemcpyAsync(d.p, dirs, 12 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(nr.p, normals, 12 * n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(fr.p, front, n, cudaMemcpyHostToDevice, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(sd.p, seeds, 4 * n, cudaMemcpyHostToDevice, c->stream));
    // same packing as k_gather: {albedo.xyz, fuzz} or {ir, 0, 0, 0}
    const float4 mat = material_type == VN_DIELECTRIC ? make_float4(albedo_fuzz_ir[3], 0.f, 0.f, 0.f)
                                                      : make_float4(albedo_fuzz_ir[0], albedo_fuzz_ir[1], albedo_fuzz_ir[2], albedo_fuzz_ir[3]);
    VN_CUDA(c, !(flags & VN_FAST) ? exact::launch_scatter(material_type, mat, d.as<float>(), nr.as<float>(), fr.as<uint8_t>(), sd.as<uint32_t>(), n,
                                                          dout.as<float>(), sc.as<uint8_t>(), sout.as<uint32_t>(), c->stream)
                                  : fast::launch_scatter(material_type, mat, d.as<float>(), nr.as<float>(), fr.as<uint8_t>(), sd.as<uint32_t>(), n,
                                                         dout.as<float>(), sc.as<uint8_t>(), sout.as<uint32_t>(), c->stream));
    VN_CUDA(c, cudaMemcpyAsync(dirs_out, dout.p, 12 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(scattered_out, sc.p, n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaMemcpyAsync(seeds_out, sout.p, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    VN_CUDA(c, cudaStreamSynchronize(c->stream));
    return VN_OK;
}

}  // extern "C"
