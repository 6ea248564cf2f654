// grid.h -- internal interface between the C ABI (vn_api.cu) and the grid builder (grid.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "grid_core.cuh"

namespace vn {

constexpr uint64_t kGridMaxPrims = 16384;

// Device-resident grid: 16-bit cell starts (n_cells + 1) and sphere references (indices into the Morton-sorted arrays).
struct GridScene {
    GridHeader h{};
    void* alloc = nullptr;
    const uint16_t* start = nullptr;
    const uint16_t* refs = nullptr;
    uint32_t max_per_cell = 0;
    bool valid = false;
};

void grid_free(GridScene& gs);

// Returns 1 when the grid was built, 0 when the scene does not suit it (more than 8 oversize spheres, too many cells or
// references, a cell with more than `max_per_cell` spheres), negative on a CUDA error with `err` set.
int grid_build(const float4* d_geom_sorted, uint64_t n, uint32_t max_per_cell, cudaStream_t stream, GridScene& out, uint32_t* launches, std::string& err);

}  // namespace vn
