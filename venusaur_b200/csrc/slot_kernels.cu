// slot_kernels.cu -- the slot-scheduled path kernel: every lane owns K path slots in shared memory and the WARP votes,
// instruction block by instruction block, on what to execute next.  Compiled twice like path_kernels.cu.
//
// Why (profiles/): in k_render_persistent a lane is bound to one path, so a warp's traversal lasts as long as its
// slowest ray, sphere tests run for the few lanes that hold a leaf and shading for the few that just finished: 10 of 32
// lanes are active per issued instruction.  Here the five kinds of work of the Monte-Carlo loop are separate
// warp-wide operations and a lane takes part in an operation whenever ANY of its K slots needs it:
//
//   N  one 4-wide node step        lanes whose current ray stands on an internal node          (optixTrace, RT cores)
//   L  one leaf = sphere tests     lanes whose current ray stands on a leaf    (__intersection__hit_sphere, :229-270)
//   W  retire + fetch              lanes whose ray is finished (hit point -> slot) or that hold no ray but a ready slot
//   O/D/M  shade one slot          opaque (Lambertian+metal, :272-366) / dielectric (:381-440) / miss (:442-450): ONE
//                                  program per operation = the material-sorted shade queue, without a queue
//   R  camera ray into a free slot (__raygen__rg, :163-204), next sample of the lane's pixel or the next pixel
//
// The scheduler keeps the node step running while >= TN lanes want it and otherwise runs the operation most lanes are
// waiting for (one REDUX over packed 6-bit counters).  State of a slot: 13 words SoA in shared memory (o|p, d,
// throughput, seed, prim, sample<<16|depth, pixel); the ray being traversed lives in registers.  A lane's slots hold
// consecutive samples of its pixel, finished in any order, so each path's radiance goes to the per-(sample, pixel)
// buffer and k_wf_accumulate adds the samples in the reference's order (RayTracer.cu:203-216): bit-identical output.
#include <cooperative_groups.h>

#include <algorithm>

#include "kernels.h"
#include "lbvh_core.cuh"

namespace cg = cooperative_groups;

#if VN_EXACT
#define VN_NS exact
#else
#define VN_NS fast
#endif

namespace vn {
namespace VN_NS {

namespace {

constexpr unsigned kFull = 0xffffffffu;
enum SlotField : int { kOx = 0, kOy, kOz, kDx, kDy, kDz, kTr, kTg, kTb, kSeed, kPrim, kMeta, kPix, kSlotFields };
// lane mask register: one nibble per slot state
constexpr int kShReady = 0, kShOpaque = 4, kShDiel = 8, kShMiss = 12, kShEmpty = 16;

__device__ __forceinline__ uint32_t fetch_ticket(uint32_t* counter) {
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, g.size());
    base = g.shfl(base, 0);
    return base + g.thread_rank();
}

template <int K, int THREADS, bool kCount>
__global__ void __launch_bounds__(THREADS, 1) k_render_slots(const __grid_constant__ RenderLaunch p, float* __restrict__ sample_rgb,
                                                            const SlotTune tune) {
    extern __shared__ float4 s_mem[];
    const uint32_t node_f4s = kWideNodeF4 * p.num_wide;
    float4* s_nodes = s_mem;
    float4* s_geom = s_nodes + (size_t)node_f4s * 8;
    float4* s_mat = s_geom + p.num_spheres;
    uint8_t* s_type = reinterpret_cast<uint8_t*>(s_mat + p.num_spheres);
    uint32_t* s_slots = reinterpret_cast<uint32_t*>(s_type + ((p.num_spheres + 15u) & ~15u));
    for (uint32_t i = threadIdx.x; i < 8u * p.num_wide; i += THREADS) {
        const uint32_t k = i / p.num_wide, j = i - k * p.num_wide;
        float4 canon[8], out[kWideNodeF4];
#pragma unroll
        for (int q = 0; q < 8; q++) canon[q] = p.wide[8ull * j + q];
        wide_octant_node(canon, k, out);
#pragma unroll
        for (int q = 0; q < (int)kWideNodeF4; q++) s_nodes[(size_t)k * node_f4s + kWideNodeF4 * j + q] = out[q];
    }
    for (uint32_t i = threadIdx.x; i < p.num_spheres; i += THREADS) { s_geom[i] = p.geom[i]; s_mat[i] = p.mat[i]; s_type[i] = p.type[i]; }
    __syncthreads();

    // field f of slot j of this lane: conflict-free (consecutive lanes, consecutive words)
    uint32_t* const my = s_slots + threadIdx.x;
#define SLOT_U(f, j) my[((f) * K + (j)) * THREADS]
#define SLOT_F(f, j) reinterpret_cast<float*>(my)[((f) * K + (j)) * THREADS]

    const uint32_t region_pixels = p.width * (p.row_end - p.row_begin);
    const uint32_t region_first = p.row_begin * p.width;

    // ---- lane state
    uint32_t m = ((1u << K) - 1u) << kShEmpty;          // every slot empty
    int cur_slot = -1;
    uint32_t cur = kEmptyScene;                         // node / leaf link of the ray in registers
    f3 o = mk3(0.0f), d = mk3(0.0f), idir = mk3(0.0f), ood = mk3(0.0f);
    float a = 1.0f, inv_a = 1.0f, tbest = kTMax;
    int prim = -1, sp = 0;
    const float4* wn = s_nodes;
    uint32_t stack[kStackSize];
    // the lane's pixel: samples are handed to slots in order, the camera seed chain is the reference's (RayTracer.cu:169-183)
    uint32_t pix = 0, pxy = 0, cam_seed = 0, s_left = 0;
    bool exhausted = false;                             // the global ticket has no pixel left for this lane
    uint32_t n_seg = 0, n_path = 0;
    TraceCounters cnt{0u, 0u};
    unsigned long long n_nodes = 0, n_sph = 0;
    uint32_t op_count[8] = {0, 0, 0, 0, 0, 0, 0, 0}, op_lanes[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // kCount only (warp-uniform)

    const uint32_t TN = tune.node_threshold, TL = tune.leaf_threshold, TW = tune.switch_threshold, TS = tune.shade_threshold,
                   TR = tune.regen_threshold;

    bool force_n = false;                               // the service pass found nothing above threshold but node work
    for (;;) {
        // ---- N burst: node steps while at least TN lanes stand on an internal node (one vote per step)
        uint32_t nN;
        for (;;) {
            const bool atN = cur_slot >= 0 && !(cur & kLeafFlag);
            nN = __popc(__ballot_sync(kFull, atN));
            if (nN < TN && !force_n) break;
            force_n = false;
            if (kCount) { op_count[0] += 1u; op_lanes[0] += nN; }
            if (atN) {
                if (kCount) cnt.nodes += 1;
                cur = wide_node_step(wn, cur, idir, ood, tbest, stack, sp);
            }
        }
        // ---- service pass: ONE packed reduction counts the lanes waiting for every other operation; every operation
        // above its threshold runs once, in pipeline order (leaf -> retire/fetch -> shade by material -> camera)
        const bool holds = cur_slot >= 0;
        uint32_t todo;
        {
            const bool atL = holds && (cur & kLeafFlag) && cur != kEmptyScene;
            const bool wantW = (holds && cur == kEmptyScene) || (!holds && ((m >> kShReady) & 0xFu));
            const bool wantO = (m >> kShOpaque) & 0xFu, wantD = (m >> kShDiel) & 0xFu, wantM = (m >> kShMiss) & 0xFu;
            const bool wantR = ((m >> kShEmpty) & 0xFu) && !(exhausted && s_left == 0u);
            const uint32_t packed = (atL ? 1u : 0u) | (wantW ? 1u << 6 : 0u) | (wantO ? 1u << 12 : 0u) | (wantD ? 1u << 18 : 0u) |
                                    (wantM ? 1u << 24 : 0u);
            const uint32_t c = __reduce_add_sync(kFull, packed);
            const uint32_t nR = __popc(__ballot_sync(kFull, wantR));
            const uint32_t nL = c & 63u, nW = (c >> 6) & 63u, nO = (c >> 12) & 63u, nD = (c >> 18) & 63u, nM = (c >> 24) & 63u;
            todo = (nL >= TL ? 2u : 0u) | (nW >= TW ? 4u : 0u) | (nO >= TS ? 8u : 0u) | (nD >= TS ? 16u : 0u) | (nM >= TS ? 32u : 0u) |
                   (nR >= TR ? 64u : 0u);
            if (todo == 0u) {
                uint32_t best = nN;
                int op = 0;
                if (nL > best) { best = nL; op = 1; }
                if (nW > best) { best = nW; op = 2; }
                if (nO > best) { best = nO; op = 3; }
                if (nM > best) { best = nM; op = 5; }
                if (nR > best) { best = nR; op = 6; }
                if (nD > best) { best = nD; op = 4; }
                if (best == 0u) break;                  // nothing left anywhere in this warp
                todo = 1u << op;
            }
            if (kCount) {
                const uint32_t lanes[7] = {nN, nL, nW, nO, nD, nM, nR};
#pragma unroll
                for (int i = 1; i < 7; i++) if (todo & (1u << i)) { op_count[i] += 1u; op_lanes[i] += lanes[i]; }
            }
        }
        if (todo & 1u) force_n = true;
        if (todo & 2u) {
            // ---- L: ONE sphere of the leaf the lane stands on (no inner loop: every lane does the same work)
            if (holds && (cur & kLeafFlag) && cur != kEmptyScene) {
                const uint32_t idx = (cur & 0x7FFFFFFFu) >> 3;
                const float4 g = s_geom[idx];
                if (kCount) cnt.spheres += 1;
                const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                if (t >= 0.0f) { tbest = t; prim = (int)idx; }
                cur = (cur & 7u) ? cur + 7u : (sp ? stack[--sp] : kEmptyScene);     // first + 1, count - 1 -- or pop
            }
        }
        if (todo & 4u) {
            // ---- W: retire the finished ray into its slot, then take the next ready slot
            if (cur_slot >= 0 && cur == kEmptyScene) {
                int cls = kShMiss;
                if (prim >= 0) {
                    const f3 hp = hit_point(o, d, tbest);               // RayTracer.cu:256
                    SLOT_F(kOx, cur_slot) = hp.x; SLOT_F(kOy, cur_slot) = hp.y; SLOT_F(kOz, cur_slot) = hp.z;
                    cls = s_type[prim] == 2u ? kShDiel : kShOpaque;
                }
                SLOT_U(kPrim, cur_slot) = (uint32_t)prim;
                m |= 1u << (cls + cur_slot);
                cur_slot = -1;
                n_seg += 1u;
            }
            if (cur_slot < 0 && ((m >> kShReady) & 0xFu)) {
                const int j = __ffs((m >> kShReady) & 0xFu) - 1;
                m &= ~(1u << (kShReady + j));
                o = mk3(SLOT_F(kOx, j), SLOT_F(kOy, j), SLOT_F(kOz, j));
                d = mk3(SLOT_F(kDx, j), SLOT_F(kDy, j), SLOT_F(kDz, j));
                idir = slab_idir(d);
                wn = s_nodes + ray_octant(d) * node_f4s;
                ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
                a = dot(d, d);
                inv_a = rcp(a);
                tbest = kTMax;
                prim = -1;
                for (uint32_t i = 0; i < p.huge.n; i++) {                  // huge spheres are not in the wide nodes (lbvh_core.cuh::HugeList)
                    const uint32_t hs = p.huge.idx[i];
                    const float4 g = s_geom[hs];
                    if (kCount) cnt.spheres += 1;
                    const float th = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                    if (th >= 0.0f) { tbest = th; prim = (int)hs; }
                }
                sp = 0;
                cur = p.wide_root;
                cur_slot = j;
            }
        }
#pragma unroll 1
        for (int op = 3; op <= 5; op++) {
            if (!(todo & (1u << op))) continue;
            // ---- O / D / M: shade one waiting slot with ONE program
            const int sh = op == 3 ? kShOpaque : (op == 4 ? kShDiel : kShMiss);
            if ((m >> sh) & 0xFu) {
                const int j = __ffs((m >> sh) & 0xFu) - 1;
                m &= ~(1u << (sh + j));
                PathState st;
                st.d = mk3(SLOT_F(kDx, j), SLOT_F(kDy, j), SLOT_F(kDz, j));
                st.thr = mk3(SLOT_F(kTr, j), SLOT_F(kTg, j), SLOT_F(kTb, j));
                const uint32_t meta = SLOT_U(kMeta, j);
                st.depth = (int)(meta & 0xFFFFu);
                bool cont = false;
                f3 result = mk3(0.0f);
                if (op == 5) {
                    result = shade_miss(st.thr, normalize(st.d));                    // RayTracer.cu:442-450
                } else if (st.depth > 0) {                                           // RayTracer.cu:275,324,384
                    const int hp = (int)SLOT_U(kPrim, j);
                    const f3 pt = mk3(SLOT_F(kOx, j), SLOT_F(kOy, j), SLOT_F(kOz, j));
                    const float4 g = s_geom[hp], mt = s_mat[hp];
                    st.seed = SLOT_U(kSeed, j);
                    f3 n;
                    bool front;
                    hit_normal(pt, st.d, g, n, front);
                    if (op == 3) {
                        const uint32_t type = s_type[hp];
                        f3 unit_direction = mk3(0.0f);
                        if (type == 1u) unit_direction = normalize(st.d);
                        cont = shade_opaque(type, mt, unit_direction, n, st);
                        if (cont) { SLOT_F(kTr, j) = st.thr.x; SLOT_F(kTg, j) = st.thr.y; SLOT_F(kTb, j) = st.thr.z; }
                    } else {
                        { DielectricConsts dc; dc.ir = mt.x; dc.inv_ir = mt.y; dc.r0_front = mt.z; dc.r0_back = mt.w; st.d = scatter_dielectric(normalize(st.d), n, front, dc, st.seed); }
                        cont = true;
                    }
                    if (cont) {
                        SLOT_F(kDx, j) = st.d.x; SLOT_F(kDy, j) = st.d.y; SLOT_F(kDz, j) = st.d.z;
                        SLOT_U(kSeed, j) = st.seed;
                        SLOT_U(kMeta, j) = meta - 1u;                                // depth -= 1; the origin slot already holds p
                    }
                }
                if (cont) {
                    m |= 1u << (kShReady + j);
                } else {
                    float* out = sample_rgb + 3ull * ((uint64_t)(meta >> 16) * region_pixels + SLOT_U(kPix, j));
                    out[0] = result.x; out[1] = result.y; out[2] = result.z;         // pixel_color += ... happens in k_wf_accumulate
                    m |= 1u << (kShEmpty + j);
                }
            }
        }
        if (todo & 64u) {
            // ---- R: next sample of the lane's pixel (or the next pixel) into a free slot
            if (((m >> kShEmpty) & 0xFu) && !(exhausted && s_left == 0u)) {
                if (s_left == 0u) {
                    for (;;) {
                        const uint32_t w = fetch_ticket(p.work_counter);
                        if (w >= p.total_work) { exhausted = true; break; }
                        const uint32_t tile = w >> 5, in = w & 31u;
                        const uint32_t ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
                        const uint32_t px = tx * 8u + (in & 7u), py = p.row_begin + ty * 4u + (in >> 3);
                        if (px < p.width && py < p.row_end) {
                            pix = py * p.width + px;
                            pxy = px | (py << 16);
                            cam_seed = tea4(pix, p.subframe_index);     // RayTracer.cu:169
                            s_left = p.spp;
                            break;
                        }
                    }
                }
                if (s_left != 0u) {
                    const int j = __ffs((m >> kShEmpty) & 0xFu) - 1;
                    f3 co, cd;
                    camera_ray(p.cam, pxy & 0xFFFFu, pxy >> 16, cam_seed, co, cd);   // RayTracer.cu:173-177
                    SLOT_F(kOx, j) = co.x; SLOT_F(kOy, j) = co.y; SLOT_F(kOz, j) = co.z;
                    SLOT_F(kDx, j) = cd.x; SLOT_F(kDy, j) = cd.y; SLOT_F(kDz, j) = cd.z;
                    SLOT_F(kTr, j) = 1.0f; SLOT_F(kTg, j) = 1.0f; SLOT_F(kTb, j) = 1.0f;
                    SLOT_U(kSeed, j) = cam_seed;                                     // prd.seed = seed: a copy (:183)
                    SLOT_U(kMeta, j) = ((p.spp - s_left) << 16) | (p.max_depth - 1u); // depth = max_depth - 1 (:184)
                    SLOT_U(kPix, j) = pix - region_first;
                    s_left -= 1u;
                    n_path += 1u;
                    m = (m & ~(1u << (kShEmpty + j))) | (1u << (kShReady + j));
                }
            }
        }
    }
#undef SLOT_U
#undef SLOT_F

    if (kCount) { n_nodes = cnt.nodes; n_sph = cnt.spheres; }
    {
        unsigned long long seg = n_seg, path = n_path;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            seg += __shfl_xor_sync(kFull, seg, off);
            path += __shfl_xor_sync(kFull, path, off);
            if (kCount) { n_nodes += __shfl_xor_sync(kFull, n_nodes, off); n_sph += __shfl_xor_sync(kFull, n_sph, off); }
        }
        if ((threadIdx.x & 31u) == 0u) {
            atomicAdd(&p.counters[0], seg);
            atomicAdd(&p.counters[1], path);
            if (kCount) {
                atomicAdd(&p.counters[2], n_nodes);
                atomicAdd(&p.counters[3], n_sph);
                for (int i = 0; i < 7; i++) {
                    atomicAdd(&p.counters[8 + 2 * i], (unsigned long long)op_count[i]);
                    atomicAdd(&p.counters[9 + 2 * i], (unsigned long long)op_lanes[i]);
                }
            }
        }
    }
}

typedef void (*SlotKernel)(const RenderLaunch, float*, const SlotTune);

SlotKernel pick_slot_kernel(int slots, int threads, bool count) {
#define VN_SLOT_CASE(K, T) if (slots == K && threads == T) return count ? k_render_slots<K, T, true> : k_render_slots<K, T, false>;
    VN_SLOT_CASE(2, 1024)
    VN_SLOT_CASE(2, 768)
    VN_SLOT_CASE(3, 768)
    VN_SLOT_CASE(3, 512)
    VN_SLOT_CASE(4, 512)
    VN_SLOT_CASE(4, 384)
#undef VN_SLOT_CASE
    return nullptr;
}

}  // namespace

size_t slot_smem_bytes(uint32_t num_wide, uint32_t num_spheres, int slots, int threads) {
    return wide_smem_bytes(num_wide, num_spheres) + (size_t)kSlotFields * 4u * (size_t)slots * (size_t)threads;
}

bool slot_config_supported(int slots, int threads) { return pick_slot_kernel(slots, threads, false) != nullptr; }

cudaError_t launch_render_slots(const RenderLaunch& p, float* sample_rgb, int slots, int threads, int blocks, const SlotTune& tune, bool count,
                                cudaStream_t stream) {
    SlotKernel k = pick_slot_kernel(slots, threads, count);
    if (!k) return cudaErrorInvalidValue;
    const size_t smem = slot_smem_bytes(p.num_wide, p.num_spheres, slots, threads);
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k<<<blocks, threads, smem, stream>>>(p, sample_rgb, tune);
    return cudaGetLastError();
}

}  // namespace VN_NS
}  // namespace vn
