// grid.cu -- builder of the uniform grid + oversize list (grid_core.cuh) for small scenes.  Runs behind the LBVH build on the
// Morton-sorted sphere array, so both structures share primitive indices.
//
// Host: 104 bytes of parameters from one read-back of the (<= 16384) sorted spheres -- oversize = radius > 3 x median, grid box,
// resolution (grid_make_header; cbrt/ceil are evaluated on the host only, so the test emulation gets the same numbers).
// Device: ONE CTA does count -> exclusive scan -> fill -> per-cell sort (ascending sphere index: deterministic) -> 16-bit
// arrays, with block barriers between the phases.  RTIOW: 48 x 1 x 48 cells, 1688 references, 4 oversize spheres, ~30 us.
#include <cstring>
#include <string>
#include <vector>

#include "grid.h"

namespace vn {

namespace {

constexpr int kGridThreads = 1024;

__global__ void __launch_bounds__(kGridThreads) k_grid_build(const GridHeader g, const float4* __restrict__ geom, uint32_t n, uint32_t* counts /* n_cells + 1 */,
                                                            uint32_t* cursor /* n_cells */, uint32_t* refs32, uint32_t cap_refs, uint16_t* __restrict__ start16,
                                                            uint16_t* __restrict__ refs16, uint32_t* __restrict__ result /* [0] refs, [1] max per cell */) {
    __shared__ uint32_t s_warp[kGridThreads / 32];
    __shared__ uint32_t s_carry, s_max;
    const uint32_t tid = threadIdx.x;
    auto is_big = [&](uint32_t i) { bool b = false; for (uint32_t k = 0; k < g.n_big; k++) b = b || g.big[k] == i; return b; };
    for (uint32_t c = tid; c <= g.n_cells; c += kGridThreads) counts[c] = 0u;
    if (tid == 0) { s_carry = 0u; s_max = 0u; }
    __syncthreads();
    // ---- count
    for (uint32_t i = tid; i < n; i += kGridThreads) {
        if (is_big(i)) continue;
        int c0[3], c1[3];
        const float4 s = geom[i];
        grid_sphere_cells(g, s, c0, c1);
        for (int z = c0[2]; z <= c1[2]; z++) for (int y = c0[1]; y <= c1[1]; y++) for (int x = c0[0]; x <= c1[0]; x++)
            atomicAdd(&counts[grid_cell_index(g, x, y, z)], 1u);
    }
    __syncthreads();
    // ---- exclusive scan of counts[0..n_cells) in chunks of the CTA; counts[n_cells] = total; cursor = copy of the starts
    for (uint32_t base = 0; base < g.n_cells; base += kGridThreads) {
        const uint32_t c = base + tid;
        const uint32_t v = c < g.n_cells ? counts[c] : 0u;
        uint32_t inc = v;
        const unsigned lane = tid & 31u, warp = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, tot = 0;
        for (int w = 0; w < kGridThreads / 32; w++) { const uint32_t t = s_warp[w]; if (w < (int)warp) wbase += t; tot += t; }
        const uint32_t carry = s_carry;
        if (c < g.n_cells) { const uint32_t e = carry + wbase + inc - v; counts[c] = e; cursor[c] = e; atomicMax(&s_max, v); }
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
    const uint32_t total = s_carry;
    if (tid == 0) { counts[g.n_cells] = total; result[0] = total; result[1] = s_max; }
    if (total > cap_refs || total > kGridMaxRefs) return;        // the host rejects the grid (uniform exit: same value in every thread)
    __syncthreads();
    // ---- fill
    for (uint32_t i = tid; i < n; i += kGridThreads) {
        if (is_big(i)) continue;
        int c0[3], c1[3];
        const float4 s = geom[i];
        grid_sphere_cells(g, s, c0, c1);
        for (int z = c0[2]; z <= c1[2]; z++) for (int y = c0[1]; y <= c1[1]; y++) for (int x = c0[0]; x <= c1[0]; x++)
            refs32[atomicAdd(&cursor[grid_cell_index(g, x, y, z)], 1u)] = i;
    }
    __syncthreads();
    // ---- per cell: insertion sort (lists are a handful of entries), then the 16-bit copies the kernels stage
    for (uint32_t c = tid; c < g.n_cells; c += kGridThreads) {
        const uint32_t b = counts[c], e = counts[c + 1u];
        for (uint32_t i = b + 1u; i < e; i++) {
            const uint32_t v = refs32[i];
            uint32_t j = i;
            while (j > b && refs32[j - 1u] > v) { refs32[j] = refs32[j - 1u]; j--; }
            refs32[j] = v;
        }
        for (uint32_t i = b; i < e; i++) refs16[i] = (uint16_t)refs32[i];
        start16[c] = (uint16_t)b;
    }
    if (tid == 0) start16[g.n_cells] = (uint16_t)total;
}

}  // namespace

void grid_free(GridScene& gs) {
    cudaFree(gs.alloc);
    gs = GridScene();
}

int grid_build(const float4* d_geom, uint64_t n64, uint32_t max_per_cell, cudaStream_t stream, GridScene& out, uint32_t* launches, std::string& err) {
    grid_free(out);
    if (n64 < 2 || n64 > kGridMaxPrims) return 0;                 // no grid: not an error
    const uint32_t n = (uint32_t)n64;
    std::vector<node_f4> geom(n);
    cudaError_t e = cudaMemcpyAsync(geom.data(), d_geom, 16ull * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { err = std::string("grid_build: read-back of the spheres: ") + cudaGetErrorString(e); return -2; }
    GridHeader g;
    if (!grid_make_header(geom.data(), n, g)) return 0;
    // one allocation: counts | cursor | refs32 | result | start16 | refs16
    const size_t cap_refs = kGridMaxRefs;
    const size_t words = (size_t)(g.n_cells + 1) + g.n_cells + cap_refs + 8;
    const size_t bytes = 4 * words + 2 * ((size_t)g.n_cells + 2 + cap_refs) + 64;
    e = cudaMalloc(&out.alloc, bytes);
    if (e != cudaSuccess) { err = std::string("grid_build: cudaMalloc: ") + cudaGetErrorString(e); out = GridScene(); return -2; }
    uint32_t* counts = static_cast<uint32_t*>(out.alloc);
    uint32_t* cursor = counts + g.n_cells + 1;
    uint32_t* refs32 = cursor + g.n_cells;
    uint32_t* result = refs32 + cap_refs;
    uint16_t* start16 = reinterpret_cast<uint16_t*>(result + 8);
    uint16_t* refs16 = start16 + ((g.n_cells + 2u) & ~1u);
    k_grid_build<<<1, kGridThreads, 0, stream>>>(g, d_geom, n, counts, cursor, refs32, (uint32_t)cap_refs, start16, refs16, result);
    if (launches) *launches += 1;
    uint32_t host[2] = {0, 0};
    e = cudaMemcpyAsync(host, result, 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("grid_build: ") + cudaGetErrorString(e); grid_free(out); return -2; }
    if (host[0] > kGridMaxRefs || host[1] > max_per_cell) { grid_free(out); return 0; }   // crowded cells: the BVH is the better structure
    g.n_refs = host[0];
    out.h = g;
    out.start = start16;
    out.refs = refs16;
    out.max_per_cell = host[1];
    out.valid = true;
    return 1;
}

}  // namespace vn
