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
