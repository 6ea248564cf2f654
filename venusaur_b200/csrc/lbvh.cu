// lbvh.cu -- LBVH builder kernels for sm_100a.  Replaces Renderer::BuildAccelerationStructures
// (Renderer.h:160-255: optixAccelComputeMemoryUsage / optixAccelBuild / optixAccelCompact over per-sphere AABBs)
// and the per-sphere SBT records of Renderer::CreateSBT (Renderer.h:452-520).
//
// Pipeline (all on one stream, no host round trip except the final node count):
//   k_centroid_bounds   centroid AABB of the scene            (warp shuffle reduce + ordered-int atomics)
//   k_morton            30-bit Morton code per sphere + index
//   rs::sort_pairs      hand-written onesweep radix sort (radix_sort.cuh)
//   k_gather            sphere records permuted into Morton order (SoA: geom, material, type) + padded leaf AABBs
//   k_karras            Karras 2012 hierarchy, one thread per internal node
//   k_refit             bottom-up AABB refit, one thread per leaf, atomic arrival counters
//   k_mark_kept + scan  which Karras nodes survive leaf collapsing, and their rank
//   k_pack              32-byte nodes, sibling pairs adjacent and 64-byte aligned
// Compiled with -fmad=false so that Morton quantisation matches the CPU emulation in tests bit for bit.
#include "lbvh.h"

#include <cstring>
#include <vector>

#include "lbvh_core.cuh"
#include "radix_sort.cuh"

namespace vn {

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ uint32_t enc_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(uint32_t e) {
    const uint32_t u = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    return __uint_as_float(u);
}

// bounds[0..2] = min centroid (ordered encoding), bounds[3..5] = max centroid.  Pre-set to 0xFFFFFFFF / 0.
__global__ void __launch_bounds__(kBlock) k_centroid_bounds(const vn_sphere* __restrict__ s, uint32_t n, uint32_t* __restrict__ bounds) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        const float c[3] = {s[i].cx, s[i].cy, s[i].cz};
#pragma unroll
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], c[a]); hi[a] = fmaxf(hi[a], c[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            atomicMin(&bounds[a], enc_ordered(lo[a]));
            atomicMax(&bounds[3 + a], enc_ordered(hi[a]));
        }
    }
}

__global__ void __launch_bounds__(kBlock) k_morton(const vn_sphere* __restrict__ s, uint32_t n, const uint32_t* __restrict__ bounds,
                                                  uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    float clo[3], cinv[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float l = dec_ordered(bounds[a]), h = dec_ordered(bounds[3 + a]);
        clo[a] = l;
        cinv[a] = h > l ? 1.0f / (h - l) : 0.0f;
    }
    codes[i] = morton30(s[i].cx, s[i].cy, s[i].cz, clo, cinv);
    idx[i] = i;
}

__global__ void __launch_bounds__(kBlock) k_gather(const vn_sphere* __restrict__ s, const uint32_t* __restrict__ sorted_idx, uint32_t n,
                                                  float pad_rel, float4* __restrict__ geom, float4* __restrict__ mat,
                                                  uint8_t* __restrict__ type, uint32_t* __restrict__ orig,
                                                  f4* __restrict__ leaf_lo, f4* __restrict__ leaf_hi) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = sorted_idx[i];
    const vn_sphere p = s[src];
    geom[i] = make_float4(p.cx, p.cy, p.cz, p.r);
    // MaterialData (RayTracer.h:27-38): {albedo, fuzz} or {ir} aliasing albedo.x
    if (p.type == VN_DIELECTRIC) {
        const DielectricConsts dc = dielectric_consts(p.fuzz_or_ir);      // vn_math.cuh: 1/ir and Schlick's r0 for both faces, IEEE
        mat[i] = make_float4(dc.ir, dc.inv_ir, dc.r0_front, dc.r0_back);
    } else {
        mat[i] = make_float4(p.ax, p.ay, p.az, p.fuzz_or_ir);
    }
    type[i] = (uint8_t)p.type;
    orig[i] = src;
    const float pad = leaf_pad(p.cx, p.cy, p.cz, p.r, pad_rel);       // lbvh_core.cuh
    leaf_lo[i] = f4{p.cx - pad, p.cy - pad, p.cz - pad, 0.0f};
    leaf_hi[i] = f4{p.cx + pad, p.cy + pad, p.cz + pad, 0.0f};
}

__global__ void __launch_bounds__(kBlock) k_karras(const uint32_t* __restrict__ codes, uint32_t n, KarrasNode* __restrict__ kn,
                                                  uint32_t* __restrict__ parent_internal, uint32_t* __restrict__ parent_leaf) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i + 1 >= n) return;
    const KarrasNode k = karras_node(codes, (int)n, (int)i);
    kn[i] = k;
    if (k.left & kChildLeaf) parent_leaf[k.left & ~kChildLeaf] = i; else parent_internal[k.left] = i;
    if (k.right & kChildLeaf) parent_leaf[k.right & ~kChildLeaf] = i; else parent_internal[k.right] = i;
}

// ---- stage 4': SAH splits for small scenes, one CTA, level by level (see lbvh_core.cuh for the per-element bodies).
// Every level: (1) each (axis, position) candidate of every active node evaluates its SAH cost by brute force over the
// node's range and does a 64-bit atomicMin on the node's slot; (2) every position is moved by a stable partition around
// its node's winning candidate, the node is emitted with Karras' id convention and its children become active.
constexpr int kSahThreads = 1024;
constexpr uint32_t kMaxSahHeight = 48;   // < kStackSize (64) with margin
constexpr uint32_t kSahInactive = 0xFFFFFFFFu;
constexpr uint32_t kWideSingleCtaMax = 16384;

__global__ void __launch_bounds__(kSahThreads) k_sah_small(uint32_t n, const f4* __restrict__ cen, const f4* __restrict__ blo, const f4* __restrict__ bhi,
                                                          uint32_t* permA, uint32_t* permB, uint32_t* ownerA, uint32_t* ownerB,
                                                          unsigned long long* best, uint32_t* nfirst, uint32_t* nlast, KarrasNode* kn,
                                                          uint32_t* parent_internal, uint32_t* parent_leaf, uint32_t* perm_final, uint32_t* height_out) {
    const uint32_t tid = threadIdx.x;
    uint32_t levels = 0;
    for (uint32_t p = tid; p < n; p += kSahThreads) { permA[p] = p; ownerA[p] = n >= 2u ? 0u : kSahInactive; }
    if (tid == 0) { nfirst[0] = 0u; nlast[0] = n - 1u; }
    __syncthreads();
    uint32_t *perm = permA, *pnext = permB, *own = ownerA, *onext = ownerB;
    for (uint32_t level = 0; level < n; level++) {
        for (uint32_t p = tid; p < n; p += kSahThreads) {
            const uint32_t id = own[p];
            if (id != kSahInactive && p == nfirst[id]) best[id] = ~0ull;
        }
        __syncthreads();
        for (uint32_t idx = tid; idx < 3u * n; idx += kSahThreads) {
            const uint32_t a = idx / n, p = idx - a * n;
            const uint32_t id = own[p];
            if (id == kSahInactive) continue;
            const uint32_t first = nfirst[id], last = nlast[id];
            uint32_t nl;
            const float cost = sah_candidate_cost(perm, first, last, p, (int)a, cen, blo, bhi, &nl);
            if (cost < 3e38f) atomicMin(&best[id], sah_pack(cost, (int)a, p - first));
        }
        __syncthreads();
        bool any = false;
        for (uint32_t p = tid; p < n; p += kSahThreads) {
            const uint32_t id = own[p];
            const uint32_t e = perm[p];
            if (id == kSahInactive) { pnext[p] = e; onext[p] = kSahInactive; continue; }
            any = true;
            const uint32_t first = nfirst[id], last = nlast[id];
            // best[] was updated with atomics, which are performed in L2 and do not refresh this SM's L1 copy of the
            // line (it holds the ~0 written above): read it with ld.cg
            const unsigned long long b = __ldcg(&best[id]);
            const int a = (int)((b >> 28) & 3ull);
            const uint32_t es = perm[first + (uint32_t)(b & 0xFFFFFFFull)];
            const SahKey key{axis_of(cen[es], a), es};
            uint32_t nL = 0, before_l = 0, before_r = 0;
            for (uint32_t q = first; q <= last; q++) {
                const uint32_t eq = perm[q];
                const bool l = sah_key_le(axis_of(cen[eq], a), eq, key);
                nL += l ? 1u : 0u;
                if (q < p) { if (l) before_l++; else before_r++; }
            }
            const bool me_left = sah_key_le(axis_of(cen[e], a), e, key);
            const uint32_t gamma = first + nL - 1u, nR = last - gamma;
            const uint32_t newpos = me_left ? first + before_l : first + nL + before_r;
            pnext[newpos] = e;
            onext[newpos] = me_left ? (nL >= 2u ? gamma : kSahInactive) : (nR >= 2u ? gamma + 1u : kSahInactive);
            if (p == first) {
                KarrasNode k;
                k.left = nL == 1u ? (kChildLeaf | first) : gamma;
                k.right = nR == 1u ? (kChildLeaf | last) : gamma + 1u;
                k.first = first; k.last = last;
                kn[id] = k;
                if (nL == 1u) parent_leaf[first] = id; else { parent_internal[gamma] = id; nfirst[gamma] = first; nlast[gamma] = gamma; }
                if (nR == 1u) parent_leaf[last] = id; else { parent_internal[gamma + 1u] = id; nfirst[gamma + 1u] = gamma + 1u; nlast[gamma + 1u] = last; }
            }
        }
        any = __syncthreads_or(any ? 1 : 0) != 0;
        uint32_t* t = perm; perm = pnext; pnext = t;
        t = own; own = onext; onext = t;
        if (!any) break;
        levels += 1u;
    }
    for (uint32_t p = tid; p < n; p += kSahThreads) perm_final[p] = perm[p];
    if (tid == 0) *height_out = levels;   // levels of internal nodes = entries the traversal stack may need
}

__global__ void __launch_bounds__(kBlock) k_compose(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i < n) out[i] = idx[perm[i]];
}

__device__ __forceinline__ f4 ldcg4(const f4* p) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    return f4{v.x, v.y, v.z, v.w};
}

// One thread per leaf walks towards the root.  The first thread to reach a node leaves; the second one finds both
// children complete (the first made its box visible with a fence before bumping the counter) and merges them.
// Child boxes are read with ld.cg: L1 is not coherent across SMs.
__global__ void __launch_bounds__(kBlock) k_refit(uint32_t n, const KarrasNode* __restrict__ kn, const uint32_t* __restrict__ parent_internal,
                                                 const uint32_t* __restrict__ parent_leaf, const f4* __restrict__ leaf_lo,
                                                 const f4* __restrict__ leaf_hi, f4* ilo, f4* ihi, uint32_t* __restrict__ arrivals) {
    const uint32_t leaf = blockIdx.x * kBlock + threadIdx.x;
    if (leaf >= n || n < 2) return;
    uint32_t node = parent_leaf[leaf];
    while (true) {
        __threadfence();
        if (atomicAdd(&arrivals[node], 1u) == 0u) return;
        const KarrasNode k = kn[node];
        const f4 la = (k.left & kChildLeaf) ? leaf_lo[k.left & ~kChildLeaf] : ldcg4(&ilo[k.left]);
        const f4 ha = (k.left & kChildLeaf) ? leaf_hi[k.left & ~kChildLeaf] : ldcg4(&ihi[k.left]);
        const f4 lb = (k.right & kChildLeaf) ? leaf_lo[k.right & ~kChildLeaf] : ldcg4(&ilo[k.right]);
        const f4 hb = (k.right & kChildLeaf) ? leaf_hi[k.right & ~kChildLeaf] : ldcg4(&ihi[k.right]);
        const float4 lo = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.0f);
        const float4 hi = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
        __stcg(reinterpret_cast<float4*>(&ilo[node]), lo);
        __stcg(reinterpret_cast<float4*>(&ihi[node]), hi);
        if (node == 0u) return;
        node = parent_internal[node];
    }
}

__global__ void __launch_bounds__(kBlock) k_mark_kept(const KarrasNode* __restrict__ kn, uint32_t n_internal, uint32_t leaf_size,
                                                     uint32_t* __restrict__ kept) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n_internal) return;
    kept[i] = (kn[i].last - kn[i].first + 1u) > leaf_size ? 1u : 0u;
}

// ---- exclusive scan of uint32 (three small kernels; 1024 elements per block)
constexpr int kScanItems = 4;
constexpr int kScanTile = kBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < kBlock / 32; w++) {
        const uint32_t t = s_warp[w];
        if (w < (int)warp) base += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kBlock) k_scan_block_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_warp[kBlock / 32];
    uint32_t v = 0;
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < n) v += in[base + k];
    uint32_t total;
    block_exclusive_scan(v, s_warp, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of sums[0..nb) in place; sums[nb] = grand total
__global__ void __launch_bounds__(kBlock) k_scan_sums(uint32_t* __restrict__ sums, uint32_t nb) {
    __shared__ uint32_t s_warp[kBlock / 32];
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += kBlock) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? sums[i] : 0u;
        uint32_t total;
        const uint32_t e = block_exclusive_scan(v, s_warp, &total);
        if (i < nb) sums[i] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0) sums[nb] = carry;
}

__global__ void __launch_bounds__(kBlock) k_scan_final(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ sums,
                                                      uint32_t* __restrict__ out) {
    __shared__ uint32_t s_warp[kBlock / 32];
    uint32_t x[kScanItems];
    uint32_t v = 0;
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { x[k] = base + k < n ? in[base + k] : 0u; v += x[k]; }
    uint32_t total;
    uint32_t e = sums[blockIdx.x] + block_exclusive_scan(v, s_warp, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { if (base + k < n) out[base + k] = e; e += x[k]; }
}

__device__ __forceinline__ void store_node(float4* nodes, uint32_t index, const PackedNode& p) {
    nodes[2 * index] = make_float4(p.a.x, p.a.y, p.a.z, p.a.w);
    nodes[2 * index + 1] = make_float4(p.b.x, p.b.y, p.b.z, p.b.w);
}

__global__ void __launch_bounds__(kBlock) k_pack(uint32_t n, const KarrasNode* __restrict__ kn, const uint32_t* __restrict__ rank,
                                                const f4* __restrict__ ilo, const f4* __restrict__ ihi, const f4* __restrict__ leaf_lo,
                                                const f4* __restrict__ leaf_hi, uint32_t leaf_size, float4* __restrict__ nodes) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i == 0) {
        // node 0: padding; node 1: root
        nodes[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        nodes[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        PackedNode root;
        if (n == 1) {
            root.a = leaf_lo[0]; root.b = leaf_hi[0];
            root.a.w = u2f(leaf_link(0u, 1u)); root.b.w = u2f(1u);
        } else {
            root.a = ilo[0]; root.b = ihi[0];
            root.a.w = u2f(n > leaf_size ? 2u : leaf_link(0u, n)); root.b.w = u2f(n);
        }
        store_node(nodes, 1u, root);
    }
    if (i + 1 >= n) return;
    if ((kn[i].last - kn[i].first + 1u) <= leaf_size) return;
    const uint32_t base = 2u + 2u * rank[i];
    store_node(nodes, base, pack_child(kn[i].left, kn, rank, ilo, ihi, leaf_lo, leaf_hi, leaf_size));
    store_node(nodes, base + 1u, pack_child(kn[i].right, kn, rank, ilo, ihi, leaf_lo, leaf_hi, leaf_size));
}

// ---- stage 7 (small scenes): 4-wide nodes from the packed pairs, breadth-first in ONE CTA so that the node order is
// deterministic (prefix sums, no atomics): each level's wide nodes open their pair (lbvh_core.cuh::wide_collapse) and
// their internal children receive consecutive indices in entry order.  result[1] = wide nodes, result[2] = levels.
__global__ void __launch_bounds__(kBlock) k_wide_build(const float4* __restrict__ nodes, uint32_t* src, uint32_t cap, float4* __restrict__ wide,
                                                      uint32_t* __restrict__ result, const HugeList huge) {
    __shared__ uint32_t s_warp[kBlock / 32];
    const uint32_t root_link = __float_as_uint(nodes[2].w);
    if (root_link & kLeafFlag) {            // a single leaf or the empty scene: nothing to widen
        if (threadIdx.x == 0) { result[1] = 0u; result[2] = 0u; }
        return;
    }
    if (threadIdx.x == 0) src[0] = root_link;
    __syncthreads();
    uint32_t lo = 0, hi = 1, levels = 0;
    while (lo < hi) {
        levels += 1u;
        uint32_t next = hi;
        for (uint32_t base = lo; base < hi; base += kBlock) {
            const uint32_t e = base + threadIdx.x;
            uint32_t ch[4], n = 0, n_internal = 0;
            if (e < hi) {
                n = wide_collapse(nodes, src[e], ch);
                for (uint32_t c = 0; c < n; c++) n_internal += (__float_as_uint(nodes[2 * ch[c]].w) & kLeafFlag) ? 0u : 1u;
            }
            uint32_t total;
            uint32_t at = next + block_exclusive_scan(n_internal, s_warp, &total);
            if (e < hi) {
                for (uint32_t c = 0; c < 4u; c++) {
                    float4 a = make_float4(3e38f, 3e38f, 3e38f, __uint_as_float(kWideEmpty)), b = make_float4(-3e38f, -3e38f, -3e38f, 0.0f);
                    if (c < n && !huge_leaf(__float_as_uint(nodes[2 * ch[c]].w), huge)) {     // huge spheres: tested before the traversal
                        a = nodes[2 * ch[c]]; b = nodes[2 * ch[c] + 1];
                        const uint32_t link = __float_as_uint(a.w);
                        if (!(link & kLeafFlag)) { if (at < cap) src[at] = link; a.w = __uint_as_float(at); at += 1u; }
                    }
                    wide[8ull * e + 2 * c] = a; wide[8ull * e + 2 * c + 1] = b;
                }
            }
            next += total;
        }
        __syncthreads();
        lo = hi; hi = next;
    }
    if (threadIdx.x == 0) { result[1] = hi; result[2] = levels; }
}

// The same for scenes of any size, one level per launch pair: count the internal children of every wide node of the level,
// exclusive scan, emit.  Node order = the single-CTA kernel's (and the host emulation's) breadth-first order.
__global__ void __launch_bounds__(kBlock) k_wide_count(const float4* __restrict__ nodes, const uint32_t* __restrict__ src, uint32_t m, uint32_t* __restrict__ cnt) {
    const uint32_t e = blockIdx.x * kBlock + threadIdx.x;
    if (e >= m) return;
    uint32_t ch[4];
    const uint32_t n = wide_collapse(nodes, src[e], ch);
    uint32_t n_internal = 0;
    for (uint32_t c = 0; c < n; c++) n_internal += (__float_as_uint(nodes[2 * ch[c]].w) & kLeafFlag) ? 0u : 1u;
    cnt[e] = n_internal;
}
__global__ void __launch_bounds__(kBlock) k_wide_emit(const float4* __restrict__ nodes, uint32_t* src, uint32_t lo, uint32_t m, const uint32_t* __restrict__ off,
                                                     uint32_t next, uint32_t cap, float4* __restrict__ wide) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= m) return;
    const uint32_t e = lo + i;
    uint32_t ch[4];
    const uint32_t n = wide_collapse(nodes, src[e], ch);
    uint32_t at = next + off[i];
    for (uint32_t c = 0; c < 4u; c++) {
        float4 a = make_float4(3e38f, 3e38f, 3e38f, __uint_as_float(kWideEmpty)), b = make_float4(-3e38f, -3e38f, -3e38f, 0.0f);
        if (c < n) {
            a = nodes[2 * ch[c]]; b = nodes[2 * ch[c] + 1];
            const uint32_t link = __float_as_uint(a.w);
            if (!(link & kLeafFlag)) { if (at < cap) src[at] = link; a.w = __uint_as_float(at); at += 1u; }
        }
        wide[8ull * e + 2 * c] = a; wide[8ull * e + 2 * c + 1] = b;
    }
}

#define LB_CHECK(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); goto fail; } \
    } while (0)

inline uint32_t blocks_for(uint64_t n) { return (uint32_t)((n + kBlock - 1) / kBlock); }

}  // namespace

void lbvh_free(LbvhScene& sc) {
    cudaFree(sc.arena);
    cudaFree(sc.wide_alloc);
    cudaFree(sc.qnodes);
    sc = LbvhScene();
}

void lbvh_workspace_free(LbvhWorkspace& ws) {
    cudaFree(ws.ptr);
    ws = LbvhWorkspace();
}

namespace {
// Bump allocator over one cudaMalloc: a build used to spend 96 % of its time in ~25 cudaMalloc/cudaFree calls.
struct Arena {
    char* base = nullptr;
    size_t off = 0;
    template <typename T> T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return p;
    }
};
// The builder's temporaries inside the cached workspace; the layout is a pure function of n, so a later refit of the same scene finds
// the hierarchy (kn, parents, ranks) where the build left it.
struct WsPtrs {
    uint32_t *bounds, *codes0, *codes1, *idx0, *idx1, *parent_i, *parent_l, *arrivals, *kept, *rank, *sums, *result, *sah_u32;
    unsigned long long* sah_best;
    void* sortws;
    KarrasNode* kn;
    f4 *leaf_lo, *leaf_hi, *ilo, *ihi;
};
void carve_ws(Arena& t, uint32_t n, WsPtrs& w) {
    const uint32_t ni = n - 1;
    const uint32_t nb = (uint32_t)((ni + kScanTile - 1) / kScanTile);
    w.bounds = t.take<uint32_t>(8);
    w.result = t.take<uint32_t>(8);
    w.codes0 = t.take<uint32_t>(n); w.codes1 = t.take<uint32_t>(n); w.idx0 = t.take<uint32_t>(n); w.idx1 = t.take<uint32_t>(n);
    w.sortws = t.take<char>(rs::workspace_bytes(n));
    w.leaf_lo = t.take<f4>(n); w.leaf_hi = t.take<f4>(n);
    w.kn = t.take<KarrasNode>(ni + 1);
    w.ilo = t.take<f4>(ni + 1); w.ihi = t.take<f4>(ni + 1);
    w.parent_i = t.take<uint32_t>(ni + 1); w.parent_l = t.take<uint32_t>(n);
    w.arrivals = t.take<uint32_t>(ni + 1); w.kept = t.take<uint32_t>(ni + 1); w.rank = t.take<uint32_t>(ni + 1);
    w.sums = t.take<uint32_t>(nb + 2);
    w.sah_u32 = t.take<uint32_t>(8ull * n);
    w.sah_best = t.take<unsigned long long>(n);
}
}  // namespace

int lbvh_build(const vn_sphere* d_spheres, uint64_t n64, uint32_t leaf_size, float pad_rel, uint32_t sah_max_prims, uint32_t wide_max_prims, float huge_factor,
               int num_sms, cudaStream_t stream,
               LbvhScene& out, LbvhWorkspace& ws, uint32_t* launches, std::string& err) {
    lbvh_free(out);
    if (n64 > (1ull << 28)) { err = "too many spheres (limit 2^28)"; return -1; }
    const uint32_t n = (uint32_t)n64;
    if (leaf_size < 1) leaf_size = 1;
    if (leaf_size > 8) leaf_size = 8;
    out.n = n;
    out.leaf_size = leaf_size;
    uint32_t launched = 0;

    if (n == 0) {
        LB_CHECK(cudaMalloc(&out.arena, 4 * sizeof(float4)));
        out.nodes = static_cast<float4*>(out.arena);
        LB_CHECK(cudaMemsetAsync(out.nodes, 0, 4 * sizeof(float4), stream));
        out.num_nodes = 2;
        out.root_link = kEmptyScene;
        return 0;
    }
    {
        const uint32_t ni = n - 1;
        const uint32_t nb = (uint32_t)((ni + kScanTile - 1) / kScanTile);
        // ---- outputs: one allocation (nodes at their upper bound 2n+2, so the node count needs no mid-build sync)
        Arena oa;
        const bool want_wide = n <= wide_max_prims && n <= kWideSingleCtaMax;        // one-CTA breadth-first kernel, nodes in the arena
        const bool want_wide_large = n <= wide_max_prims && n > kWideSingleCtaMax;   // one launch pair per level, own allocation
        oa.take<float4>(2 * (2ull * n + 2)); oa.take<float4>(n); oa.take<float4>(n); oa.take<uint32_t>(n); oa.take<uint32_t>(n); oa.take<uint8_t>(n);
        if (want_wide) oa.take<float4>(8ull * n);
        const size_t out_bytes = oa.off + 256;
        LB_CHECK(cudaMalloc(&out.arena, out_bytes));
        oa = Arena{static_cast<char*>(out.arena), 0};
        out.nodes = oa.take<float4>(2 * (2ull * n + 2));
        out.geom = oa.take<float4>(n);
        out.mat = oa.take<float4>(n);
        out.orig = oa.take<uint32_t>(n);
        out.codes = oa.take<uint32_t>(n);
        out.type = oa.take<uint8_t>(n);
        out.wide = want_wide ? oa.take<float4>(8ull * n) : nullptr;
        // ---- temporaries: one cached workspace that only grows
        Arena ta;
        WsPtrs w;
        carve_ws(ta, n, w);
        const size_t need = ta.off + 256;
        if (ws.bytes < need) {
            cudaFree(ws.ptr);
            ws = LbvhWorkspace();
            LB_CHECK(cudaMalloc(&ws.ptr, need));
            ws.bytes = need;
        }
        ta = Arena{static_cast<char*>(ws.ptr), 0};
        carve_ws(ta, n, w);
        ws.topology_n = 0;                       // the workspace no longer describes the previous scene; set again when this build succeeds
        uint32_t *bounds = w.bounds, *codes0 = w.codes0, *codes1 = w.codes1, *idx0 = w.idx0, *idx1 = w.idx1, *parent_i = w.parent_i, *parent_l = w.parent_l,
                 *arrivals = w.arrivals, *kept = w.kept, *rank = w.rank, *sums = w.sums, *result = w.result, *sah_u32 = w.sah_u32;
        unsigned long long* sah_best = w.sah_best;
        void* sortws = w.sortws;
        KarrasNode* kn = w.kn;
        f4 *leaf_lo = w.leaf_lo, *leaf_hi = w.leaf_hi, *ilo = w.ilo, *ihi = w.ihi;

        const uint32_t init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
        LB_CHECK(cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        LB_CHECK(cudaMemsetAsync(arrivals, 0, 4ull * (ni + 1), stream));
        LB_CHECK(cudaMemsetAsync(rank, 0, 4ull * (ni + 1), stream));
        LB_CHECK(cudaMemsetAsync(sums, 0, 4ull * (nb + 2), stream));

        const uint32_t red_blocks = (uint32_t)std::min<uint64_t>((uint64_t)num_sms * 8ull, blocks_for(n));
        k_centroid_bounds<<<red_blocks, kBlock, 0, stream>>>(d_spheres, n, bounds);
        k_morton<<<blocks_for(n), kBlock, 0, stream>>>(d_spheres, n, bounds, codes0, idx0);
        launched += 2;
        const int which = rs::sort_pairs(codes0, idx0, codes1, idx1, n, 30, sortws, stream, num_sms, &launched);
        uint32_t* codes = which ? codes1 : codes0;
        uint32_t* idx = which ? idx1 : idx0;
        k_gather<<<blocks_for(n), kBlock, 0, stream>>>(d_spheres, idx, n, pad_rel, out.geom, out.mat, out.type, out.orig, leaf_lo, leaf_hi);
        launched += 1;
        LB_CHECK(cudaMemcpyAsync(out.codes, codes, 4ull * n, cudaMemcpyDeviceToDevice, stream));   // kept for vn_morton_codes()
        uint32_t* height = result;               // written by k_sah_small only
        const bool use_sah = ni > 0 && n <= sah_max_prims;
        out.sah = use_sah;
        if (use_sah) {
            // small scene: SAH splits instead of Karras' spatial medians, then re-gather in the final primitive order
            uint32_t *permA = sah_u32, *permB = sah_u32 + n, *ownA = sah_u32 + 2ull * n, *ownB = sah_u32 + 3ull * n, *nfirst = sah_u32 + 4ull * n,
                     *nlast = sah_u32 + 5ull * n, *perm_final = sah_u32 + 6ull * n, *final_idx = sah_u32 + 7ull * n;
            k_sah_small<<<1, kSahThreads, 0, stream>>>(n, reinterpret_cast<const f4*>(out.geom), leaf_lo, leaf_hi, permA, permB, ownA, ownB, sah_best,
                                                      nfirst, nlast, kn, parent_i, parent_l, perm_final, height);
            k_compose<<<blocks_for(n), kBlock, 0, stream>>>(perm_final, idx, n, final_idx);
            // the SAH kernel read geom/leaf boxes in Morton order; they are rewritten in the final order on the stream after it
            k_gather<<<blocks_for(n), kBlock, 0, stream>>>(d_spheres, final_idx, n, pad_rel, out.geom, out.mat, out.type, out.orig, leaf_lo, leaf_hi);
            launched += 3;
        }
        if (ni > 0) {
            if (!use_sah) k_karras<<<blocks_for(ni), kBlock, 0, stream>>>(codes, n, kn, parent_i, parent_l);
            k_refit<<<blocks_for(n), kBlock, 0, stream>>>(n, kn, parent_i, parent_l, leaf_lo, leaf_hi, ilo, ihi, arrivals);
            k_mark_kept<<<blocks_for(ni), kBlock, 0, stream>>>(kn, ni, leaf_size, kept);
            k_scan_block_sums<<<nb, kBlock, 0, stream>>>(kept, ni, sums);
            k_scan_sums<<<1, kBlock, 0, stream>>>(sums, nb);
            k_scan_final<<<nb, kBlock, 0, stream>>>(kept, ni, sums, rank);
            launched += 6;
        }
        k_pack<<<blocks_for(n), kBlock, 0, stream>>>(n, kn, rank, ilo, ihi, leaf_lo, leaf_hi, leaf_size, out.nodes);
        launched += 1;
        if (want_wide) {
            // huge spheres (host decision from one read-back of the sorted spheres: small scenes only) stay out of the wide nodes
            std::vector<node_f4> hgeom(n);
            LB_CHECK(cudaMemcpyAsync(hgeom.data(), out.geom, 16ull * n, cudaMemcpyDeviceToHost, stream));
            LB_CHECK(cudaStreamSynchronize(stream));
            huge_list_from_geom(hgeom.data(), n, leaf_size, out.huge, huge_factor);
            // sah_u32 is free again here (the SAH pass, if any, has been consumed by the second gather): queue of pair links
            k_wide_build<<<1, kBlock, 0, stream>>>(out.nodes, sah_u32, n, out.wide, result, out.huge);
            launched += 1;
        }
        // one host round trip at the end: kept-node count, root link, root bounds
        uint32_t host[12] = {0};
        LB_CHECK(cudaMemcpyAsync(&host[0], sums + nb, 4, cudaMemcpyDeviceToHost, stream));
        LB_CHECK(cudaMemcpyAsync(&host[4], reinterpret_cast<const char*>(out.nodes) + 2 * sizeof(float4), 2 * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        if (use_sah) LB_CHECK(cudaMemcpyAsync(&host[1], height, 4, cudaMemcpyDeviceToHost, stream));
        if (want_wide) LB_CHECK(cudaMemcpyAsync(&host[2], result + 1, 8, cudaMemcpyDeviceToHost, stream));
        LB_CHECK(cudaStreamSynchronize(stream));
        LB_CHECK(cudaGetLastError());
        const uint32_t total_kept = ni > 0 ? host[0] : 0u;
        out.num_nodes = 2ull + 2ull * total_kept;
        out.root_link = host[7];
        memcpy(out.bounds_lo, &host[4], 12);
        memcpy(out.bounds_hi, &host[8], 12);
        out.num_wide = want_wide ? host[2] : 0u;
        out.wide_levels = want_wide ? host[3] : 0u;
        if (want_wide_large && total_kept > 0u && !(out.root_link & kLeafFlag)) {
            const uint32_t cap = total_kept;                        // every wide node is a kept internal pair node
            const uint32_t nbw = (cap + kScanTile - 1) / kScanTile;
            uint32_t *src = nullptr, *cnt = nullptr, *off = nullptr, *wsums = nullptr;
            LB_CHECK(cudaMalloc(&out.wide_alloc, 128ull * cap));
            out.wide = static_cast<float4*>(out.wide_alloc);
            LB_CHECK(cudaMalloc(&src, 4ull * (3ull * cap + nbw + 8)));
            cnt = src + cap; off = cnt + cap; wsums = off + cap;
            cudaError_t werr = cudaMemcpyAsync(src, &out.root_link, 4, cudaMemcpyHostToDevice, stream);
            uint32_t lo = 0, hi = 1, levels = 0;
            while (werr == cudaSuccess && lo < hi) {
                const uint32_t m = hi - lo, nbm = (m + kScanTile - 1) / kScanTile;
                k_wide_count<<<blocks_for(m), kBlock, 0, stream>>>(out.nodes, src + lo, m, cnt);
                k_scan_block_sums<<<nbm, kBlock, 0, stream>>>(cnt, m, wsums);
                k_scan_sums<<<1, kBlock, 0, stream>>>(wsums, nbm);
                k_scan_final<<<nbm, kBlock, 0, stream>>>(cnt, m, wsums, off);
                k_wide_emit<<<blocks_for(m), kBlock, 0, stream>>>(out.nodes, src, lo, m, off, hi, cap, out.wide);
                launched += 5;
                uint32_t total = 0;
                werr = cudaMemcpyAsync(&total, wsums + nbm, 4, cudaMemcpyDeviceToHost, stream);
                if (werr == cudaSuccess) werr = cudaStreamSynchronize(stream);
                lo = hi; hi += total; levels += 1u;
                if (hi > cap) { werr = cudaErrorUnknown; break; }
            }
            cudaFree(src);
            LB_CHECK(werr);
            out.num_wide = hi; out.wide_levels = levels;
        }
        out.height = use_sah ? host[1] : 0u;     // 0 = not measured (Karras: bounded by the 30-bit key + index tie-break)
        if (use_sah && out.height > kMaxSahHeight) {
            // a degenerate scene (e.g. hundreds of coincident spheres) makes SAH peel one primitive per level; the
            // traversal stack holds kStackSize entries, so fall back to Karras, whose height is bounded by the key width
            if (launches) *launches += launched;
            return lbvh_build(d_spheres, n64, leaf_size, pad_rel, 0u, wide_max_prims, huge_factor, num_sms, stream, out, ws, launches, err);
        }
    }
    ws.topology_n = n;
    ws.topology_wide = out.wide != nullptr && out.wide_alloc == nullptr;
    if (launches) *launches += launched;
    return 0;
fail:
    lbvh_free(out);
    return -2;
}

// Refit: the spheres moved (same count, same order in the caller's array), the hierarchy stays.  Re-gathers the sphere records in the
// existing traversal order, recomputes the leaf boxes, re-runs the bottom-up refit over the hierarchy the last build left in the
// workspace, repacks the nodes and rebuilds the 4-wide nodes of a small scene: 4-5 launches instead of ~18 and no sort.  The result is
// a valid BVH of the moved spheres (traversal == brute force); its quality degrades as the spheres leave their old neighbourhoods, which
// is when the caller rebuilds.  Returns 1 when the workspace no longer holds this scene's hierarchy (the caller then rebuilds).
int lbvh_refit(const vn_sphere* d_spheres, float pad_rel, float huge_factor, cudaStream_t stream, LbvhScene& out, LbvhWorkspace& ws,
               uint32_t* launches, std::string& err) {
    const uint32_t n = (uint32_t)out.n;
    if (n < 2 || ws.topology_n != n || !ws.ptr || out.wide_alloc != nullptr) return 1;
    Arena ta{static_cast<char*>(ws.ptr), 0};
    WsPtrs w;
    carve_ws(ta, n, w);
    const uint32_t ni = n - 1;
    uint32_t launched = 0;
    LB_CHECK(cudaMemsetAsync(w.arrivals, 0, 4ull * (ni + 1), stream));
    k_gather<<<blocks_for(n), kBlock, 0, stream>>>(d_spheres, out.orig, n, pad_rel, out.geom, out.mat, out.type, out.orig, w.leaf_lo, w.leaf_hi);
    k_refit<<<blocks_for(n), kBlock, 0, stream>>>(n, w.kn, w.parent_i, w.parent_l, w.leaf_lo, w.leaf_hi, w.ilo, w.ihi, w.arrivals);
    k_pack<<<blocks_for(n), kBlock, 0, stream>>>(n, w.kn, w.rank, w.ilo, w.ihi, w.leaf_lo, w.leaf_hi, out.leaf_size, out.nodes);
    launched += 3;
    if (out.wide) {
        std::vector<node_f4> hgeom(n);
        LB_CHECK(cudaMemcpyAsync(hgeom.data(), out.geom, 16ull * n, cudaMemcpyDeviceToHost, stream));
        LB_CHECK(cudaStreamSynchronize(stream));
        huge_list_from_geom(hgeom.data(), n, out.leaf_size, out.huge, huge_factor);
        k_wide_build<<<1, kBlock, 0, stream>>>(out.nodes, w.sah_u32, n, out.wide, w.result, out.huge);
        launched += 1;
    }
    {
        uint32_t host[12] = {0};
        LB_CHECK(cudaMemcpyAsync(&host[4], reinterpret_cast<const char*>(out.nodes) + 2 * sizeof(float4), 2 * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        if (out.wide) LB_CHECK(cudaMemcpyAsync(&host[2], w.result + 1, 8, cudaMemcpyDeviceToHost, stream));
        LB_CHECK(cudaStreamSynchronize(stream));
        LB_CHECK(cudaGetLastError());
        memcpy(out.bounds_lo, &host[4], 12);
        memcpy(out.bounds_hi, &host[8], 12);
        if (out.wide) { out.num_wide = host[2]; out.wide_levels = host[3]; }
    }
    if (launches) *launches += launched;
    return 0;
fail:
    return -2;
}

// ---- quantised pair nodes.  A pair of the packed hierarchy is 64 bytes -- two 256-bit requests per traversal step, and the traversal of
// scenes that live in L2 / HBM is bound by the memory system's request rate (anything that ADDS requests loses).  The same
// pair in 32 bytes: the twelve planes of the two child boxes as 16-bit fixed point over the ROOT box (lo rounded down and hi rounded up, one
// more quantum of slack each, which covers every rounding of the decode), then the two links.  No per-node origin or exponent: a plane
// is q * scale + root_lo, so its slab parameter is q * (scale * idir) + (root_lo - o) * idir with both factors formed once per RAY
// (vn_math.cuh::pair_node_step_q_dev).  One quantum is extent / 65535: 3 mm on the 200-unit scene, 8 mm on the 500-unit one, against sphere
// radii of 0.1-0.3.  Entry `cur` and `cur + 1` of the uint4 array hold the pair whose packed nodes are `cur` and `cur + 1`.
__global__ void __launch_bounds__(256) k_quantize_pairs(const float4* __restrict__ nodes, uint32_t num_nodes, uint4* __restrict__ q) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;          // pair index: packed nodes 2p, 2p + 1
    if (p < 1u || 2u * p + 1u >= num_nodes) return;
    const float4 rlo = nodes[2], rhi = nodes[3];                           // node 1 = the root: the box everything is quantised over
    const float lo3[3] = {rlo.x, rlo.y, rlo.z}, hi3[3] = {rhi.x, rhi.y, rhi.z};
    float inv[3];
    for (int a = 0; a < 3; a++) { const float e = hi3[a] - lo3[a]; inv[a] = e > 0.0f ? 65535.0f / e : 0.0f; }
    uint32_t w[8];
    uint32_t qv[12];
    for (uint32_t c = 0; c < 2u; c++) {
        const float4 lo = nodes[2u * (2u * p + c)], hi = nodes[2u * (2u * p + c) + 1u];
        const float l3[3] = {lo.x, lo.y, lo.z}, h3[3] = {hi.x, hi.y, hi.z};
        for (int a = 0; a < 3; a++) {
            const float ql = floorf((l3[a] - lo3[a]) * inv[a]) - 1.0f, qh = ceilf((h3[a] - lo3[a]) * inv[a]) + 1.0f;
            qv[6u * c + a] = (uint32_t)fminf(fmaxf(ql, 0.0f), 65535.0f);
            qv[6u * c + 3 + a] = (uint32_t)fminf(fmaxf(qh, 0.0f), 65535.0f);
        }
        w[6u + c] = __float_as_uint(lo.w);
    }
    for (int k = 0; k < 6; k++) w[k] = qv[2 * k] | (qv[2 * k + 1] << 16);
    q[2u * p] = make_uint4(w[0], w[1], w[2], w[3]);
    q[2u * p + 1u] = make_uint4(w[4], w[5], w[6], w[7]);
}

int lbvh_quantize(LbvhScene& sc, cudaStream_t stream, uint32_t* launches, std::string& err) {
    if (!sc.nodes || sc.num_nodes < 4) return 0;
    if (sc.qnodes_cap < sc.num_nodes) {
        cudaFree(sc.qnodes); sc.qnodes = nullptr; sc.qnodes_cap = 0;
        const cudaError_t e = cudaMalloc(&sc.qnodes, sc.num_nodes * sizeof(uint4));
        if (e != cudaSuccess) { err = std::string("cudaMalloc(quantised nodes): ") + cudaGetErrorString(e); return -2; }
        sc.qnodes_cap = sc.num_nodes;
    }
    const uint32_t pairs = (uint32_t)(sc.num_nodes / 2);
    k_quantize_pairs<<<(pairs + 255u) / 256u, 256, 0, stream>>>(sc.nodes, (uint32_t)sc.num_nodes, sc.qnodes);
    if (launches) *launches += 1;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("k_quantize_pairs: ") + cudaGetErrorString(e); return -2; }
    return 0;
}

int radix_sort_pairs_device(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, int key_bits, int num_sms,
                            cudaStream_t stream, uint32_t* launches, std::string& err) {
    void* ws = nullptr;
    cudaError_t e = cudaMalloc(&ws, rs::workspace_bytes(n ? n : 1));
    if (e != cudaSuccess) { err = std::string("cudaMalloc(sort workspace): ") + cudaGetErrorString(e); return -1; }
    const int which = rs::sort_pairs(k0, v0, k1, v1, n, key_bits, ws, stream, num_sms, launches);
    e = cudaStreamSynchronize(stream);
    cudaFree(ws);
    if (e != cudaSuccess) { err = std::string("radix sort: ") + cudaGetErrorString(e); return -1; }
    return which;
}

}  // namespace vn
