// lbvh_core.cuh -- per-element bodies of the LBVH builder kernels (lbvh.cu): Morton codes, Karras 2012 hierarchy
// emission, node packing.  Written as host/device functions so the tests can also drive them in a CPU loop.
//
// Replaces optixAccelBuild / optixAccelCompact (Renderer.h:160-255), which is closed-source OptiX: the algorithm
// here is ours (30-bit Morton codes of centroids -> radix sort -> Karras hierarchy -> bottom-up refit -> pack).
#pragma once

#include "vn_math.cuh"

namespace vn {

// ---- stage 2: 30-bit Morton code (10 bits per axis) of a centroid normalised to the centroid bounds
VN_HD uint32_t expand_bits10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
VN_HD uint32_t quantize10(float c, float lo, float inv_extent) {
    float q = (c - lo) * inv_extent * 1024.0f;
    q = fminf(fmaxf(q, 0.0f), 1023.0f);
    return (uint32_t)q;
}
VN_HD uint32_t morton30(float cx, float cy, float cz, const float* clo, const float* cinv) {
    return (expand_bits10(quantize10(cx, clo[0], cinv[0])) << 2) |
           (expand_bits10(quantize10(cy, clo[1], cinv[1])) << 1) |
           expand_bits10(quantize10(cz, clo[2], cinv[2]));
}

VN_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}

// ---- stage 4: Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (2012).
// Keys are the sorted Morton codes; equal codes are disambiguated by their position, i.e. the key is (code, i).
VN_HD int karras_delta(const uint32_t* __restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint32_t a = codes[i], b = codes[j];
    if (a == b) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz32(a ^ b);
}

constexpr uint32_t kChildLeaf = 0x80000000u;

// Half-size of a sphere's leaf box: |r| (sphere.h:17-28 computes fabsf(radius) and then forgets to use it) grown by pad_rel
// (default 1 %) + 1e-6 + 2^-17 of the coordinates' magnitude.  The slab tests of the traversal only have to be conservative, and they are
// by this margin: the box contains the hit-point gate's box (vn_math.cuh::hit_gate_ok: 0.5 % + 2^-19 of the magnitude) with room for
// the rounding of the slab test itself (a few 2^-23 of the distance travelled), at any scene scale.
VN_HD float leaf_pad(float cx, float cy, float cz, float r, float pad_rel) {
    return fabsf(r) * (1.0f + pad_rel) + 1e-6f + (fabsf(cx) + fabsf(cy) + fabsf(cz) + fabsf(r)) * 7.62939453125e-06f;
}

struct KarrasNode {
    uint32_t left, right;    // kChildLeaf | leaf index, or internal node index
    uint32_t first, last;    // covered range of sorted primitives (inclusive)
};

VN_HD KarrasNode karras_node(const uint32_t* __restrict__ codes, int n, int i) {
    const int d = (karras_delta(codes, n, i, i + 1) - karras_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = karras_delta(codes, n, i, i - d);
    int lmax = 2;
    while (karras_delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = karras_delta(codes, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    KarrasNode k;
    k.left = (lo == gamma) ? (kChildLeaf | (uint32_t)gamma) : (uint32_t)gamma;
    k.right = (hi == gamma + 1) ? (kChildLeaf | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    k.first = (uint32_t)lo;
    k.last = (uint32_t)hi;
    return k;
}

// ---- stage 6: pack into 32-byte nodes.  Layout (vn_node32 in include/venusaur_b200.h):
//   node 0 unused (keeps child pairs 64-byte aligned), node 1 = root, pair of kept internal node with rank r at
//   (2+2r, 3+2r).  A Karras subtree with <= leaf_size primitives collapses into one leaf (its range is contiguous).
struct PackedNode { f4 a, b; };   // a = {lo.xyz, link}, b = {hi.xyz, aux}

VN_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

VN_HD uint32_t leaf_link(uint32_t first, uint32_t count) { return kLeafFlag | (first << 3) | (count - 1u); }

// Describes one child of a kept internal node.
VN_HD PackedNode pack_child(uint32_t child, const KarrasNode* __restrict__ kn, const uint32_t* __restrict__ rank,
                            const f4* __restrict__ ilo, const f4* __restrict__ ihi,
                            const f4* __restrict__ leaf_lo, const f4* __restrict__ leaf_hi, uint32_t leaf_size) {
    PackedNode out;
    if (child & kChildLeaf) {
        const uint32_t li = child & ~kChildLeaf;
        out.a = leaf_lo[li]; out.b = leaf_hi[li];
        out.a.w = u2f(leaf_link(li, 1u));
        out.b.w = u2f(1u);
    } else {
        const KarrasNode c = kn[child];
        const uint32_t count = c.last - c.first + 1u;
        out.a = ilo[child]; out.b = ihi[child];
        out.a.w = u2f(count > leaf_size ? (2u + 2u * rank[child]) : leaf_link(c.first, count));
        out.b.w = u2f(count);
    }
    return out;
}

}  // namespace vn

// ---- optional stage 4': surface-area-heuristic splits for SMALL scenes (n <= sah_max_prims), replacing Karras' spatial-
// median splits.  The tree keeps Karras' conventions (contiguous primitive ranges, left child id = last index of its
// range, right child id = first index of its range, root id 0), so refit and packing are unchanged.  Candidate split
// = (axis, primitive): everything whose centroid key is <= that primitive's goes left.  Per-element bodies:
namespace vn {

struct SahKey { float c; uint32_t id; };
VN_HD bool sah_key_le(float c, uint32_t id, SahKey k) { return c < k.c || (c == k.c && id <= k.id); }
VN_HD float box_area(float lx, float ly, float lz, float hx, float hy, float hz) {
    const float dx = fmaxf(hx - lx, 0.0f), dy = fmaxf(hy - ly, 0.0f), dz = fmaxf(hz - lz, 0.0f);
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}
VN_HD float axis_of(const f4& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// SAH cost of splitting range [first,last] of perm[] at candidate (axis a, position p); returns +inf-like (3e38) when the
// split would leave one side empty.  cen = primitive centres, blo/bhi = padded primitive boxes (all indexed by primitive).
VN_HD float sah_candidate_cost(const uint32_t* __restrict__ perm, uint32_t first, uint32_t last, uint32_t p, int a,
                               const f4* __restrict__ cen, const f4* __restrict__ blo, const f4* __restrict__ bhi, uint32_t* n_left_out) {
    const uint32_t ep = perm[p];
    const SahKey key{axis_of(cen[ep], a), ep};
    float llx = 3e38f, lly = 3e38f, llz = 3e38f, lhx = -3e38f, lhy = -3e38f, lhz = -3e38f;
    float rlx = 3e38f, rly = 3e38f, rlz = 3e38f, rhx = -3e38f, rhy = -3e38f, rhz = -3e38f;
    uint32_t nl = 0, nr = 0;
    for (uint32_t q = first; q <= last; q++) {
        const uint32_t e = perm[q];
        const f4 lo = blo[e], hi = bhi[e];
        if (sah_key_le(axis_of(cen[e], a), e, key)) {
            llx = fminf(llx, lo.x); lly = fminf(lly, lo.y); llz = fminf(llz, lo.z);
            lhx = fmaxf(lhx, hi.x); lhy = fmaxf(lhy, hi.y); lhz = fmaxf(lhz, hi.z);
            nl++;
        } else {
            rlx = fminf(rlx, lo.x); rly = fminf(rly, lo.y); rlz = fminf(rlz, lo.z);
            rhx = fmaxf(rhx, hi.x); rhy = fmaxf(rhy, hi.y); rhz = fmaxf(rhz, hi.z);
            nr++;
        }
    }
    *n_left_out = nl;
    if (nr == 0u) return 3e38f;
    return box_area(llx, lly, llz, lhx, lhy, lhz) * (float)nl + box_area(rlx, rly, rlz, rhx, rhy, rhz) * (float)nr;
}

// (cost, axis, position-in-range) packed so that an unsigned 64-bit min picks the cheapest candidate, ties broken by
// axis then position: deterministic on CPU and GPU.
VN_HD unsigned long long sah_pack(float cost, int a, uint32_t rel_pos) {
    return ((unsigned long long)f2u(cost) << 32) | ((unsigned long long)a << 28) | (unsigned long long)rel_pos;
}

}  // namespace vn

// ---- optional stage 7: 4-wide nodes derived from the packed pairs, for scenes that are traversed out of shared memory.
// A wide node is a packed internal node whose pair has been opened greedily (largest surface area first) until it has
// four children or only leaves are left.  Canonical form (global memory, what vn_read_wide_bvh returns): 8 float4 per
// wide node, child c = { {lo.xyz, link}, {hi.xyz, count} }; link = wide-node index, a leaf link (as in the packed
// nodes) or kWideEmpty with an inverted box.  Half as many traversal steps as the pairs, four slab tests per step.
namespace vn {

constexpr uint32_t kWideEmpty = 0xFFFFFFFFu;
constexpr uint32_t kWideMaxLevels = 20;    // 3 pushes per level + 1 <= kStackSize
constexpr uint32_t kWideGlobalMaxLevels = 42;   // same bound for kWideStackSize (traversal from global memory)

// Spheres so large that nearly every ray enters their box (RTIOW's radius-1000 ground: radius > 50 x the median radius) are
// left out of the WIDE nodes and tested by every ray before the traversal instead -- by all lanes of a warp together, where a
// leaf visit costs a divergent leaf turn.  The pair nodes keep them, so every other kernel is unaffected.  Needs one-sphere
// leaves (leaf_size 1, the automatic choice for small scenes); at most 8.
constexpr uint32_t kHugeMax = 8;
constexpr float kHugeFactor = 50.0f;
struct HugeList { uint32_t n; uint32_t idx[kHugeMax]; };
VN_HD bool huge_leaf(uint32_t link, const HugeList& h) {
    if (!(link & kLeafFlag) || (link & 7u) != 0u || link == kWideEmpty) return false;
    const uint32_t first = (link & 0x7FFFFFFFu) >> 3;
    bool is = false;
    for (uint32_t k = 0; k < kHugeMax; k++) is = is || (k < h.n && h.idx[k] == first);
    return is;
}

VN_HD float packed_area(const node_f4& a, const node_f4& b) { return box_area(a.x, a.y, a.z, b.x, b.y, b.z); }

// Opens the pair at `pair_link`; out[] = packed-node indices of the (2..4) children in stable left-to-right order.
VN_HD uint32_t wide_collapse(const node_f4* __restrict__ nodes, uint32_t pair_link, uint32_t* out) {
    uint32_t n = 2;
    out[0] = pair_link; out[1] = pair_link + 1u;
    while (n < 4u) {
        int best = -1;
        float best_area = -1.0f;
        for (uint32_t i = 0; i < n; i++) {
            const node_f4 a = nodes[2 * out[i]], b = nodes[2 * out[i] + 1];
            if (f2u(a.w) & kLeafFlag) continue;
            const float area = packed_area(a, b);
            if (area > best_area) { best_area = area; best = (int)i; }
        }
        if (best < 0) break;
        const uint32_t link = f2u(nodes[2 * out[best]].w);
        for (uint32_t i = n; i > (uint32_t)best + 1u; i--) out[i] = out[i - 1];
        out[best] = link; out[best + 1] = link + 1u;
        n++;
    }
    return n;
}

// Octant-specialised shared-memory form of one wide node: 7 float4 = near.x[4] near.y[4] near.z[4] far.x[4] far.y[4]
// far.z[4] link[4], children sorted front to back for rays of that octant (by box centre along the octant diagonal,
// ties by canonical position; empty slots last), near/far planes already selected per axis.  With the order fixed per
// octant the traversal needs no distance compare and no per-axis min/max.
VN_HD void wide_octant_node(const node_f4* __restrict__ canon /* 8 float4 */, uint32_t octant, node_f4* out /* 7 float4 */) {
    float key[4];
    int ord[4];
    for (int c = 0; c < 4; c++) {
        const node_f4 lo = canon[2 * c], hi = canon[2 * c + 1];
        const float kx = lo.x + hi.x, ky = lo.y + hi.y, kz = lo.z + hi.z;
        key[c] = f2u(lo.w) == kWideEmpty ? 3e38f : ((octant & 1u) ? -kx : kx) + ((octant & 2u) ? -ky : ky) + ((octant & 4u) ? -kz : kz);
        ord[c] = c;
    }
    for (int i = 1; i < 4; i++)                                       // stable insertion sort
        for (int j = i; j > 0 && key[ord[j]] < key[ord[j - 1]]; j--) { const int t = ord[j]; ord[j] = ord[j - 1]; ord[j - 1] = t; }
    float v[7][4];
    for (int s = 0; s < 4; s++) {
        const node_f4 lo = canon[2 * ord[s]], hi = canon[2 * ord[s] + 1];
        v[0][s] = (octant & 1u) ? hi.x : lo.x; v[3][s] = (octant & 1u) ? lo.x : hi.x;
        v[1][s] = (octant & 2u) ? hi.y : lo.y; v[4][s] = (octant & 2u) ? lo.y : hi.y;
        v[2][s] = (octant & 4u) ? hi.z : lo.z; v[5][s] = (octant & 4u) ? lo.z : hi.z;
        v[6][s] = lo.w;
    }
    for (int r = 0; r < 7; r++) { out[r].x = v[r][0]; out[r].y = v[r][1]; out[r].z = v[r][2]; out[r].w = v[r][3]; }
}

}  // namespace vn

#include <algorithm>
#include <vector>
namespace vn {
// host: the huge list from the sorted spheres {c.xyz, r} of a small scene with one-sphere leaves (empty otherwise)
inline void huge_list_from_geom(const node_f4* geom, uint32_t n, uint32_t leaf_size, HugeList& h, float factor = kHugeFactor) {
    h = HugeList();
    if (leaf_size != 1u || n < 16u) return;
    std::vector<float> r(n);
    for (uint32_t i = 0; i < n; i++) r[i] = fabsf(geom[i].w);
    std::vector<float> tmp = r;
    std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
    if (!(factor > 0.0f)) return;
    const float limit = factor * tmp[n / 2];
    for (uint32_t i = 0; i < n && h.n < kHugeMax; i++) if (r[i] > limit) h.idx[h.n++] = i;
}
}  // namespace vn
