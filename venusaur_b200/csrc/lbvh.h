// lbvh.h -- internal interface between the C ABI (vn_api.cu) and the LBVH builder (lbvh.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <string>

#include "../../include/venusaur_b200.h"
#include "lbvh_core.cuh"

namespace vn {

// Device-resident scene in traversal order (Morton-sorted), as the trace kernels read it.
struct LbvhScene {
    void* arena = nullptr;       // one allocation holding every array below
    uint64_t n = 0;
    float4* geom = nullptr;      // {cx, cy, cz, r}
    float4* mat = nullptr;       // {albedo.xyz, fuzz} or {ir, 0, 0, 0}
    uint8_t* type = nullptr;     // material.h:9-14
    uint32_t* orig = nullptr;    // sorted position -> index in the caller's sphere array
    uint32_t* codes = nullptr;   // sorted Morton codes
    float4* nodes = nullptr;     // num_nodes x 32 B (vn_node32)
    uint64_t num_nodes = 0;
    uint32_t root_link = 0xFFFFFFFFu;
    uint32_t leaf_size = 0;
    float4* wide = nullptr;      // num_wide x 128 B: 4-wide nodes derived from the pairs (small scenes only), see lbvh_core.cuh
    void* wide_alloc = nullptr;  // large scenes: the wide nodes have their own allocation
    uint32_t num_wide = 0, wide_levels = 0;
    HugeList huge{};             // spheres left out of the wide nodes: every ray tests them before the traversal
    uint32_t height = 0;         // levels of internal nodes (the traversal stack must hold that many entries)
    bool sah = false;            // splits chosen by the surface-area heuristic (small scenes) instead of Karras' spatial medians
    float bounds_lo[3] = {0, 0, 0}, bounds_hi[3] = {0, 0, 0};
    uint4* qnodes = nullptr;     // optional (lbvh_quantize): the pairs again, 32 B each = one 256-bit load: child boxes as 16-bit fixed point over the root box
    uint64_t qnodes_cap = 0;     //   (uint4 entries allocated)
};

// Quantised copy of the packed pair nodes for scenes traversed from L2 / HBM (see lbvh.cu::k_quantize_pairs).  0 or a negative status.
int lbvh_quantize(LbvhScene& sc, cudaStream_t stream, uint32_t* launches, std::string& err);

void lbvh_free(LbvhScene& sc);

// Scratch memory of the builder, cached across builds (grows on demand).
struct LbvhWorkspace {
    void* ptr = nullptr;
    size_t bytes = 0;
    uint64_t topology_n = 0;     // != 0: the workspace still holds the hierarchy (Karras nodes, parents, ranks) of the last build of that many spheres
    bool topology_wide = false;
};
void lbvh_workspace_free(LbvhWorkspace& ws);

// Builds the packed LBVH for n spheres already resident on the device.  Returns 0 or a negative status with `err` set.
int lbvh_build(const vn_sphere* d_spheres, uint64_t n, uint32_t leaf_size, float pad_rel, uint32_t sah_max_prims, uint32_t wide_max_prims, float huge_factor,
               int num_sms, cudaStream_t stream,
               LbvhScene& out, LbvhWorkspace& ws, uint32_t* launches, std::string& err);

// The spheres moved: same hierarchy, new boxes (see lbvh.cu).  0 = done, 1 = not possible (rebuild), negative = error.
int lbvh_refit(const vn_sphere* d_spheres, float pad_rel, float huge_factor, cudaStream_t stream, LbvhScene& out, LbvhWorkspace& ws,
               uint32_t* launches, std::string& err);

// Onesweep sort of device (key, value) pairs; returns 0/1 = which buffer pair holds the result, or -1.
int radix_sort_pairs_device(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, int key_bits, int num_sms,
                            cudaStream_t stream, uint32_t* launches, std::string& err);

}  // namespace vn
