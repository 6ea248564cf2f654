// grid_core.cuh -- the second closest-hit structure for SMALL scenes: a uniform grid over the similar-sized spheres plus a short
// list of oversize spheres that every ray tests first.  Per-element bodies of the builder (grid.cu) and the traversal, as
// host/device functions so the tests can drive them on the CPU (tests/host_harness.cpp).
//
// Why next to the LBVH: optixTrace's contract is "closest hit in [1e-3, 1e16]" (RayTracer.cu:190-202), not a particular
// structure.  The RTIOW final scene is 482 spheres of radius 0.2 on a 22 x 22 lattice, three of radius 1 and one of radius
// 1000.  Its SAH-split 4-wide BVH costs 5.0 node steps (4 slab tests each) + 2.0 sphere tests per ray segment; a 48 x 1 x 48
// grid with the four large spheres in the oversize list costs 2.1 cell steps + 5.7 sphere tests, about 40 % of the thread
// instructions, and most of them (the oversize tests) are executed by all lanes of a warp together.  Same sphere_root(),
// same primitive indices (the Morton-sorted arrays of the LBVH build): the two structures return bit-identical hits.
#pragma once

#include "vn_math.cuh"

namespace vn {

constexpr uint32_t kGridMaxBig = 8;          // oversize spheres tested by every ray
constexpr uint32_t kGridMaxCells = 32768;    // start[] and refs[] are 16-bit: the structure is meant to live in shared memory
constexpr uint32_t kGridMaxRefs = 65535;
constexpr float kGridBigFactor = 3.0f;       // oversize = radius > 3 x median radius
constexpr float kGridLambda = 4.0f;          // target cells per (non-oversize) sphere

struct GridHeader {                          // 104 bytes, lives in constant memory with the launch parameters
    float lo[3], inv_cell[3], cell[3], hi[3];
    uint32_t res[3];
    uint32_t n_cells, n_refs, n_big;
    uint32_t big[kGridMaxBig];
};

// radius used for cell overlap: |r| + 1 % + an absolute epsilon of the cell size (the DDA below is float arithmetic: a ray that
// clips a cell by less than that must still find the sphere in the neighbouring cell)
// (+ 2^-17 of the coordinates' magnitude, like the BVH's leaf boxes, lbvh_core.cuh::leaf_pad: grazing hits whose float quadratic is noise
// at the scale of the coordinates -- 33 of 272 M segments on the RTIOW frame, all on the radius-1000 ground -- are found by both structures)
VN_HD float grid_pad_radius(const node_f4& s, float min_cell) {
    return fabsf(s.w) * 1.01f + 1e-4f * min_cell + (fabsf(s.x) + fabsf(s.y) + fabsf(s.z) + fabsf(s.w)) * 7.62939453125e-06f;
}

VN_HD int grid_cell_coord(float p, float lo, float inv_cell, uint32_t res) {
    const float q = floorf((p - lo) * inv_cell);
    const float c = fminf(fmaxf(q, 0.0f), (float)(res - 1u));
    return (int)c;
}

// Cell range covered by a (non-oversize) sphere.
VN_HD void grid_sphere_cells(const GridHeader& g, const node_f4& s, int* c0, int* c1) {
    const float min_cell = fminf(g.cell[0], fminf(g.cell[1], g.cell[2]));
    const float r = grid_pad_radius(s, min_cell);
    const float c[3] = {s.x, s.y, s.z};
    for (int a = 0; a < 3; a++) {
        c0[a] = grid_cell_coord(c[a] - r, g.lo[a], g.inv_cell[a], g.res[a]);
        c1[a] = grid_cell_coord(c[a] + r, g.lo[a], g.inv_cell[a], g.res[a]);
    }
}
VN_HD uint32_t grid_cell_index(const GridHeader& g, int x, int y, int z) { return ((uint32_t)z * g.res[1] + (uint32_t)y) * g.res[0] + (uint32_t)x; }

}  // namespace vn
#include <algorithm>
#include <vector>
namespace vn {
// Resolution from the box of the non-oversize spheres and their number: res_a = ceil(extent_a * cbrt(lambda * n / volume)),
// at least 1, shrunk uniformly until the cell count fits.  Returns false when there is nothing to put in a grid.
inline bool grid_resolution(const float* lo, const float* hi, uint32_t n_small, GridHeader& g) {
    if (n_small == 0u) return false;
    float ext[3];
    for (int a = 0; a < 3; a++) {
        const float pad = 1e-3f * (hi[a] - lo[a]) + 1e-4f;
        g.lo[a] = lo[a] - pad; g.hi[a] = hi[a] + pad;
        ext[a] = g.hi[a] - g.lo[a];
    }
    float k = cbrtf(kGridLambda * (float)n_small / (ext[0] * ext[1] * ext[2]));
    for (int iter = 0; iter < 64; iter++) {
        unsigned long long cells = 1;
        for (int a = 0; a < 3; a++) {
            const float r = ceilf(ext[a] * k);
            g.res[a] = (uint32_t)fminf(fmaxf(r, 1.0f), 1024.0f);
            cells *= g.res[a];
        }
        if (cells <= kGridMaxCells) break;
        k *= 0.9f;
    }
    g.n_cells = g.res[0] * g.res[1] * g.res[2];
    for (int a = 0; a < 3; a++) { g.cell[a] = ext[a] / (float)g.res[a]; g.inv_cell[a] = 1.0f / g.cell[a]; }
    return g.n_cells <= kGridMaxCells;
}

// Host side of the build: picks the oversize spheres (radius > 3 x median), the grid box and its resolution from the sorted
// sphere array {c.xyz, r}.  104 bytes of parameters; counting / filling the cells is done by the kernels of grid.cu.  Returns
// false when the scene does not suit the structure (too many oversize spheres, nothing left for the grid, too many cells).
inline bool grid_make_header(const node_f4* geom, uint32_t n, GridHeader& g) {
    g = GridHeader();
    if (n < 2u) return false;
    std::vector<float> r(n);
    for (uint32_t i = 0; i < n; i++) r[i] = fabsf(geom[i].w);
    std::vector<float> tmp = r;
    std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
    const float limit = kGridBigFactor * tmp[n / 2];
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    uint32_t n_small = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (r[i] > limit) {
            if (g.n_big == kGridMaxBig) return false;
            g.big[g.n_big++] = i;
            continue;
        }
        const float c[3] = {geom[i].x, geom[i].y, geom[i].z};
        for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], c[a] - r[i]); hi[a] = std::max(hi[a], c[a] + r[i]); }
        n_small++;
    }
    return grid_resolution(lo, hi, n_small, g);
}
inline bool grid_is_big(const GridHeader& g, uint32_t i) {
    for (uint32_t k = 0; k < g.n_big; k++) if (g.big[k] == i) return true;
    return false;
}

// ---- traversal state of one ray (3-D DDA).  Split into setup / step so the path kernel can interleave lanes by vote.
struct GridRay {
    float tmax[3], tdelta[3];
    int c[3], step[3];
    float t_exit;        // parameter where the ray leaves the grid box (or the best hit so far, whichever is smaller at setup)
    uint32_t q, q_end;   // cursor into refs[] for the current cell
    bool alive;
};

VN_HD void grid_load_cell(const GridHeader& g, const uint16_t* __restrict__ start, GridRay& r) {
    const uint32_t ci = grid_cell_index(g, r.c[0], r.c[1], r.c[2]);
    r.q = start[ci];
    r.q_end = start[ci + 1u];
}

// Clips the ray to the grid box and positions the DDA on the first cell.  tbest = closest oversize hit so far.
VN_HD void grid_ray_setup(const GridHeader& g, const uint16_t* __restrict__ start, f3 o, f3 d, float tbest, GridRay& r) {
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    float idir[3];
    bool pos[3];                         // direction sign per axis, taken from the guarded component (so -0 walks like -1e-30)
    float t0 = 0.0f, t1 = tbest;
    for (int k = 0; k < 3; k++) {
        const float dk = fabsf(dd[k]) < 1e-30f ? copysignf(1e-30f, dd[k]) : dd[k];
        pos[k] = dk > 0.0f;
        idir[k] = 1.0f / dk;
        const float ta = (g.lo[k] - oo[k]) * idir[k], tb = (g.hi[k] - oo[k]) * idir[k];
        t0 = fmaxf(t0, fminf(ta, tb));
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    r.alive = t0 <= t1;
    r.t_exit = t1;
    r.q = r.q_end = 0u;
    if (!r.alive) return;
    for (int k = 0; k < 3; k++) {
        const float p = oo[k] + dd[k] * t0;
        r.c[k] = grid_cell_coord(p, g.lo[k], g.inv_cell[k], g.res[k]);
        r.step[k] = pos[k] ? 1 : -1;
        const float next_plane = g.lo[k] + (float)(r.c[k] + (pos[k] ? 1 : 0)) * g.cell[k];
        r.tmax[k] = (next_plane - oo[k]) * idir[k];
        r.tdelta[k] = g.cell[k] * fabsf(idir[k]);
    }
    grid_load_cell(g, start, r);
}

// Leaves the current cell (all its spheres have been tested): stops when the best hit lies inside it, with a slack for the
// float DDA, or when the ray leaves the grid; otherwise loads the next cell's sphere range.
VN_HD void grid_ray_advance(const GridHeader& g, const uint16_t* __restrict__ start, float tbest, GridRay& r) {
    const int k = r.tmax[0] < r.tmax[1] ? (r.tmax[0] < r.tmax[2] ? 0 : 2) : (r.tmax[1] < r.tmax[2] ? 1 : 2);
    const float t_cell_exit = k == 0 ? r.tmax[0] : (k == 1 ? r.tmax[1] : r.tmax[2]);
    if (tbest < t_cell_exit * (1.0f - 1e-5f) - 1e-6f) { r.alive = false; return; }
    const int nc = (k == 0 ? r.c[0] : (k == 1 ? r.c[1] : r.c[2])) + (k == 0 ? r.step[0] : (k == 1 ? r.step[1] : r.step[2]));
    const int lim = (int)(k == 0 ? g.res[0] : (k == 1 ? g.res[1] : g.res[2]));
    if (nc < 0 || nc >= lim) { r.alive = false; return; }
    if (k == 0) { r.c[0] = nc; r.tmax[0] += r.tdelta[0]; }
    else if (k == 1) { r.c[1] = nc; r.tmax[1] += r.tdelta[1]; }
    else { r.c[2] = nc; r.tmax[2] += r.tdelta[2]; }
    grid_load_cell(g, start, r);
}

// Whole closest-hit query (host tests, trace-rays entry point; the path kernel inlines the same pieces around a warp vote).
template <bool kCount>
VN_HD void closest_hit_grid(const GridHeader& g, const uint16_t* __restrict__ start, const uint16_t* __restrict__ refs,
                            const node_f4* __restrict__ geom, f3 o, f3 d, float& t_out, int& prim_out, TraceCounters& cnt, bool gate = false) {
    float tbest = kTMax;
    int prim = -1;
    const float a = dot(d, d);
    const float inv_a = rcp(a);
    for (uint32_t i = 0; i < g.n_big; i++) {
        const uint32_t s = g.big[i];
        const node_f4 sp = geom[s];
        if (kCount) cnt.spheres += 1;
        const float t = sphere_root(o, d, a, inv_a, sp.x, sp.y, sp.z, sp.w, kTMin, tbest);
        if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, sp.x, sp.y, sp.z, sp.w))) { tbest = t; prim = (int)s; }
    }
    if (g.n_cells != 0u) {
        GridRay r;
        grid_ray_setup(g, start, o, d, tbest, r);
        while (r.alive) {
            if (kCount) cnt.nodes += 1;
            for (; r.q < r.q_end; r.q++) {
                const uint32_t s = refs[r.q];
                const node_f4 sp = geom[s];
                if (kCount) cnt.spheres += 1;
                const float t = sphere_root(o, d, a, inv_a, sp.x, sp.y, sp.z, sp.w, kTMin, tbest);
                if (t >= 0.0f && (!gate || hit_gate_ok(o, d, t, sp.x, sp.y, sp.z, sp.w))) { tbest = t; prim = (int)s; }
            }
            grid_ray_advance(g, start, tbest, r);
        }
    }
    t_out = tbest;
    prim_out = prim;
}

}  // namespace vn
