// host_harness.cpp -- TEST ONLY.  Compiles the product's device math (venusaur_b200/csrc/vn_math.cuh, exact build)
// and the per-element LBVH builder bodies (lbvh_core.cuh) with g++ and drives them in plain CPU loops, so the
// logic that the CUDA kernels wrap can be checked against the oracle in the GPU-less container.  The product never
// uses this: libvenusaur_b200.so has no CPU path.
//
// g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -Ivenusaur_b200/csrc tests/host_harness.cpp
#define VN_EXACT 1
#include "lbvh_core.cuh"
#include "grid_core.cuh"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

using namespace vn;

struct hh_sphere { float cx, cy, cz, r, ax, ay, az, fuzz_or_ir; uint32_t type; };

struct hh_params {
    uint32_t width, height, spp, subframe_index, max_depth;
    float origin[3], u[3], v[3], w[3], lens_radius;
};

static inline node_f4 dielectric_mat(float ir) { const DielectricConsts dc = dielectric_consts(ir); return node_f4{dc.ir, dc.inv_ir, dc.r0_front, dc.r0_back}; }
static uint32_t g_sah_max = 4096;   // same default as the library's "sah_max_prims" option
static int g_use_oct = 0;
static int g_gate = 0;          // hh_set_gate(1): pair-node / global wide traversals apply the hit-point gate (vn_math.cuh::hit_gate_ok)
static int g_use_grid = 0;         // hh_set_grid(1): closest hit through the uniform grid + oversize list
static int g_use_wide = 0;         // hh_set_wide(2): canonical wide nodes with distance sort (closest_hit_wide_global); hh_set_wide(1): traverse the 4-wide octant-sorted nodes like k_render_persistent<.., kWide>
static int g_seq_postpone = 0;   // hh_set_oct(1): traverse octant-mirrored node copies like k_render_persistent<.., kOct=true>

struct HostBvh {
    std::vector<node_f4> nodes_oct;
    std::vector<node_f4> nodes, geom, mat;
    std::vector<uint8_t> type;
    std::vector<uint32_t> orig, codes;
    std::vector<KarrasNode> kn;
    uint32_t root_link = kEmptyScene;
    std::vector<node_f4> wide;        // canonical 4-wide nodes (8 float4 each), BFS order
    std::vector<node_f4> wide_oct;    // 8 octant-specialised copies (7 float4 per node)
    uint32_t wide_root = kEmptyScene, wide_levels = 0;
    HugeList huge{};                  // spheres left out of the wide nodes (tested before the traversal)
    uint32_t leaf_size = 0;
    GridHeader grid;                  // uniform grid + oversize list (grid_core.cuh); grid_ok = the scene suits it
    bool grid_ok = false;
    std::vector<uint16_t> grid_start, grid_refs;
};

// Sequential emulation of k_sah_small (lbvh.cu): same per-element functions, same level-by-level order.
static void sah_small_host(uint32_t n, const f4* cen, const f4* blo, const f4* bhi, std::vector<KarrasNode>& kn, std::vector<uint32_t>& perm_final) {
    const uint32_t INACT = 0xFFFFFFFFu;
    std::vector<uint32_t> permA(n), permB(n), ownA(n), ownB(n), nfirst(n), nlast(n);
    std::vector<unsigned long long> best(n);
    for (uint32_t p = 0; p < n; p++) { permA[p] = p; ownA[p] = n >= 2 ? 0u : INACT; }
    nfirst[0] = 0; nlast[0] = n - 1;
    uint32_t *perm = permA.data(), *pnext = permB.data(), *own = ownA.data(), *onext = ownB.data();
    for (uint32_t level = 0; level < n; level++) {
        for (uint32_t p = 0; p < n; p++) { const uint32_t id = own[p]; if (id != INACT && p == nfirst[id]) best[id] = ~0ull; }
        for (uint32_t idx = 0; idx < 3 * n; idx++) {
            const uint32_t a = idx / n, p = idx - a * n, id = own[p];
            if (id == INACT) continue;
            uint32_t nl;
            const float cost = sah_candidate_cost(perm, nfirst[id], nlast[id], p, (int)a, cen, blo, bhi, &nl);
            if (cost < 3e38f) best[id] = std::min(best[id], sah_pack(cost, (int)a, p - nfirst[id]));
        }
        bool any = false;
        // children write nfirst/nlast for ids that are not active in this level (see the kernel), so in-place is safe
        for (uint32_t p = 0; p < n; p++) {
            const uint32_t id = own[p], e = perm[p];
            if (id == INACT) { pnext[p] = e; onext[p] = INACT; continue; }
            any = true;
            const uint32_t first = nfirst[id], last = nlast[id];
            const unsigned long long b = best[id];
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
            onext[newpos] = me_left ? (nL >= 2u ? gamma : INACT) : (nR >= 2u ? gamma + 1u : INACT);
        }
        // node emission after the partition pass (the kernel does it inside; nfirst/nlast of current nodes are not touched)
        for (uint32_t p = 0; p < n; p++) {
            const uint32_t id = own[p];
            if (id == INACT || p != nfirst[id]) continue;
            const uint32_t first = nfirst[id], last = nlast[id];
            const unsigned long long b = best[id];
            const int a = (int)((b >> 28) & 3ull);
            const uint32_t es = perm[first + (uint32_t)(b & 0xFFFFFFFull)];
            const SahKey key{axis_of(cen[es], a), es};
            uint32_t nL = 0;
            for (uint32_t q = first; q <= last; q++) nL += sah_key_le(axis_of(cen[perm[q]], a), perm[q], key) ? 1u : 0u;
            const uint32_t gamma = first + nL - 1u, nR = last - gamma;
            KarrasNode k;
            k.left = nL == 1u ? (kChildLeaf | first) : gamma;
            k.right = nR == 1u ? (kChildLeaf | last) : gamma + 1u;
            k.first = first; k.last = last;
            kn[id] = k;
            if (nL != 1u) { nfirst[gamma] = first; nlast[gamma] = gamma; }
            if (nR != 1u) { nfirst[gamma + 1u] = gamma + 1u; nlast[gamma + 1u] = last; }
        }
        std::swap(perm, pnext);
        std::swap(own, onext);
        if (!any) break;
    }
    perm_final.assign(perm, perm + n);
}

// Sequential emulation of k_wide_build (lbvh.cu): breadth-first, level by level; the internal children of a level's
// entries receive consecutive indices in entry order, then canonical child order.
static void build_wide(HostBvh& B) {
    B.wide.clear(); B.wide_oct.clear(); B.wide_levels = 0;
    B.wide_root = B.root_link;
    huge_list_from_geom(B.geom.data(), (uint32_t)B.geom.size(), B.leaf_size, B.huge);
    if (B.geom.size() > 16384) B.huge = HugeList();          // the product only does this for the one-CTA wide build (<= 16384 spheres)
    if (B.root_link & kLeafFlag) return;                 // a single leaf (or the empty scene): no wide nodes
    B.wide_root = 0;
    std::vector<uint32_t> src{B.root_link};              // packed pair link of every wide node
    size_t lo = 0;
    while (lo < src.size()) {
        const size_t hi = src.size();
        B.wide_levels++;
        for (size_t e = lo; e < hi; e++) {
            uint32_t ch[4];
            const uint32_t n = wide_collapse(B.nodes.data(), src[e], ch);
            B.wide.resize(8 * (e + 1), node_f4{0, 0, 0, 0});
            for (uint32_t c = 0; c < 4; c++) {
                node_f4 a{3e38f, 3e38f, 3e38f, u2f(kWideEmpty)}, b{-3e38f, -3e38f, -3e38f, u2f(0u)};
                if (c < n && !huge_leaf(f2u(B.nodes[2 * ch[c]].w), B.huge)) {
                    a = B.nodes[2 * ch[c]]; b = B.nodes[2 * ch[c] + 1];
                    const uint32_t link = f2u(a.w);
                    if (!(link & kLeafFlag)) { a.w = u2f((uint32_t)src.size()); src.push_back(link); }
                }
                B.wide[8 * e + 2 * c] = a; B.wide[8 * e + 2 * c + 1] = b;
            }
        }
        lo = hi;
    }
    const size_t W = src.size();
    if (W > 65536) return;                                 // large scenes are traversed in canonical form (closest_hit_wide_global)
    B.wide_oct.resize(8 * 7 * W);
    for (uint32_t k = 0; k < 8; k++)
        for (size_t j = 0; j < W; j++) wide_octant_node(&B.wide[8 * j], k, &B.wide_oct[(k * W + j) * 7]);
}

// Sequential emulation of grid_build (grid.cu): header on the host (the product computes it there too), then count / scan /
// fill with every cell's references in ascending sphere order.
static void build_grid(HostBvh& B) {
    B.grid_ok = false; B.grid_start.clear(); B.grid_refs.clear();
    const uint32_t n = (uint32_t)B.geom.size();
    if (!grid_make_header(B.geom.data(), n, B.grid)) return;
    GridHeader& g = B.grid;
    std::vector<std::vector<uint32_t>> cells(g.n_cells);
    for (uint32_t i = 0; i < n; i++) {
        if (grid_is_big(g, i)) continue;
        int c0[3], c1[3];
        grid_sphere_cells(g, B.geom[i], c0, c1);
        for (int z = c0[2]; z <= c1[2]; z++) for (int y = c0[1]; y <= c1[1]; y++) for (int x = c0[0]; x <= c1[0]; x++)
            cells[grid_cell_index(g, x, y, z)].push_back(i);
    }
    size_t refs = 0;
    for (auto& c : cells) refs += c.size();
    if (refs > kGridMaxRefs) return;
    g.n_refs = (uint32_t)refs;
    B.grid_start.resize(g.n_cells + 1);
    size_t off = 0;
    for (uint32_t c = 0; c < g.n_cells; c++) { B.grid_start[c] = (uint16_t)off; for (uint32_t i : cells[c]) B.grid_refs.push_back((uint16_t)i); off += cells[c].size(); }
    B.grid_start[g.n_cells] = (uint16_t)off;
    B.grid_ok = true;
}

static void build(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, HostBvh& B) {
    B = HostBvh();
    B.leaf_size = leaf_size;
    if (n == 0) { B.nodes.resize(4); return; }
    std::vector<f4> llo(n), lhi(n);
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = 0; i < n; i++) {
        const float c[3] = {s[i].cx, s[i].cy, s[i].cz};
        for (int a = 0; a < 3; a++) { clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]); }
    }
    float cinv[3];
    for (int a = 0; a < 3; a++) cinv[a] = chi[a] > clo[a] ? 1.0f / (chi[a] - clo[a]) : 0.0f;
    std::vector<uint32_t> codes(n), idx(n);
    for (uint32_t i = 0; i < n; i++) codes[i] = morton30(s[i].cx, s[i].cy, s[i].cz, clo, cinv);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
    B.codes.resize(n); B.orig = idx; B.geom.resize(n); B.mat.resize(n); B.type.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        const hh_sphere& p = s[idx[i]];
        B.codes[i] = codes[idx[i]];
        B.geom[i] = node_f4{p.cx, p.cy, p.cz, p.r};
        B.mat[i] = p.type == 2 ? dielectric_mat(p.fuzz_or_ir) : node_f4{p.ax, p.ay, p.az, p.fuzz_or_ir};
        B.type[i] = (uint8_t)p.type;
        const float pad = vn::leaf_pad(p.cx, p.cy, p.cz, p.r, pad_rel);
        llo[i] = f4{p.cx - pad, p.cy - pad, p.cz - pad, 0};
        lhi[i] = f4{p.cx + pad, p.cy + pad, p.cz + pad, 0};
    }
    const uint32_t ni = n - 1;
    B.kn.resize(ni);
    bool use_sah = ni > 0 && n <= g_sah_max;
    std::vector<uint32_t> perm_final;
    if (use_sah) {
        // same guard as lbvh_build: a SAH tree taller than 48 levels falls back to Karras
        std::vector<f4> cen0(n);
        for (uint32_t i = 0; i < n; i++) cen0[i] = f4{B.geom[i].x, B.geom[i].y, B.geom[i].z, B.geom[i].w};
        sah_small_host(n, cen0.data(), llo.data(), lhi.data(), B.kn, perm_final);
        std::vector<uint32_t> h(ni, 0);
        std::vector<std::pair<uint32_t, int>> st{{0u, 0}};
        // post-order height
        std::vector<uint32_t> order2; std::vector<uint32_t> stk{0};
        while (!stk.empty()) { uint32_t i = stk.back(); stk.pop_back(); order2.push_back(i); if (!(B.kn[i].left & kChildLeaf)) stk.push_back(B.kn[i].left); if (!(B.kn[i].right & kChildLeaf)) stk.push_back(B.kn[i].right); }
        for (auto it = order2.rbegin(); it != order2.rend(); ++it) { const KarrasNode& k = B.kn[*it]; uint32_t hl = (k.left & kChildLeaf) ? 0 : h[k.left], hr = (k.right & kChildLeaf) ? 0 : h[k.right]; h[*it] = 1 + std::max(hl, hr); }
        if (h[0] > 48) use_sah = false;
    }
    if (use_sah) {
        // small scene: SAH splits, then everything re-gathered in the final primitive order (codes stay Morton-sorted)
        std::vector<uint32_t> final_idx(n);
        for (uint32_t i = 0; i < n; i++) final_idx[i] = idx[perm_final[i]];
        B.orig = final_idx;
        for (uint32_t i = 0; i < n; i++) {
            const hh_sphere& p = s[final_idx[i]];
            B.geom[i] = node_f4{p.cx, p.cy, p.cz, p.r};
            B.mat[i] = p.type == 2 ? dielectric_mat(p.fuzz_or_ir) : node_f4{p.ax, p.ay, p.az, p.fuzz_or_ir};
            B.type[i] = (uint8_t)p.type;
            const float pad = vn::leaf_pad(p.cx, p.cy, p.cz, p.r, pad_rel);
            llo[i] = f4{p.cx - pad, p.cy - pad, p.cz - pad, 0};
            lhi[i] = f4{p.cx + pad, p.cy + pad, p.cz + pad, 0};
        }
    } else {
        for (uint32_t i = 0; i < ni; i++) B.kn[i] = karras_node(B.codes.data(), (int)n, (int)i);
    }
    std::vector<f4> ilo(ni), ihi(ni);
    // refit: children before parents == process internal nodes in post-order (iterative)
    if (ni) {
        std::vector<uint32_t> order; order.reserve(ni);
        std::vector<uint32_t> st{0};
        while (!st.empty()) {
            uint32_t i = st.back(); st.pop_back(); order.push_back(i);
            if (!(B.kn[i].left & kChildLeaf)) st.push_back(B.kn[i].left);
            if (!(B.kn[i].right & kChildLeaf)) st.push_back(B.kn[i].right);
        }
        for (auto it = order.rbegin(); it != order.rend(); ++it) {
            const KarrasNode& k = B.kn[*it];
            auto lo = [&](uint32_t c) { return (c & kChildLeaf) ? llo[c & ~kChildLeaf] : ilo[c]; };
            auto hi = [&](uint32_t c) { return (c & kChildLeaf) ? lhi[c & ~kChildLeaf] : ihi[c]; };
            f4 a = lo(k.left), b = lo(k.right), c = hi(k.left), d = hi(k.right);
            ilo[*it] = f4{std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z), 0};
            ihi[*it] = f4{std::max(c.x, d.x), std::max(c.y, d.y), std::max(c.z, d.z), 0};
        }
    }
    std::vector<uint32_t> rank(ni ? ni : 1, 0);
    uint32_t kept = 0;
    for (uint32_t i = 0; i < ni; i++) { rank[i] = kept; kept += (B.kn[i].last - B.kn[i].first + 1u) > leaf_size; }
    B.nodes.assign(2 * (2 + 2 * (size_t)kept), node_f4{0, 0, 0, 0});
    // root = node 1
    if (ni == 0) {
        B.nodes[2] = node_f4{llo[0].x, llo[0].y, llo[0].z, u2f(leaf_link(0, 1))};
        B.nodes[3] = node_f4{lhi[0].x, lhi[0].y, lhi[0].z, u2f(1u)};
    } else {
        const bool root_kept = n > leaf_size;
        B.nodes[2] = node_f4{ilo[0].x, ilo[0].y, ilo[0].z, u2f(root_kept ? 2u : leaf_link(0, n))};
        B.nodes[3] = node_f4{ihi[0].x, ihi[0].y, ihi[0].z, u2f(n)};
        for (uint32_t i = 0; i < ni; i++) {
            if ((B.kn[i].last - B.kn[i].first + 1u) <= leaf_size) continue;
            const uint32_t base = 2u + 2u * rank[i];
            PackedNode L = pack_child(B.kn[i].left, B.kn.data(), rank.data(), ilo.data(), ihi.data(), llo.data(), lhi.data(), leaf_size);
            PackedNode R = pack_child(B.kn[i].right, B.kn.data(), rank.data(), ilo.data(), ihi.data(), llo.data(), lhi.data(), leaf_size);
            B.nodes[2 * base + 0] = node_f4{L.a.x, L.a.y, L.a.z, L.a.w};
            B.nodes[2 * base + 1] = node_f4{L.b.x, L.b.y, L.b.z, L.b.w};
            B.nodes[2 * base + 2] = node_f4{R.a.x, R.a.y, R.a.z, R.a.w};
            B.nodes[2 * base + 3] = node_f4{R.b.x, R.b.y, R.b.z, R.b.w};
        }
    }
    B.root_link = f2u(B.nodes[2].w);
    // 8 copies in near/far-plane form, one per direction octant (same construction as the kernel prologue)
    const size_t nn = B.nodes.size() / 2;
    B.nodes_oct.resize(8 * B.nodes.size());
    for (uint32_t k = 0; k < 8; k++)
        for (size_t j = 0; j < nn; j++) {
            const node_f4 lo = B.nodes[2 * j], hi = B.nodes[2 * j + 1];
            node_f4 nr = lo, fr = hi;
            if (k & 1) { nr.x = hi.x; fr.x = lo.x; }
            if (k & 2) { nr.y = hi.y; fr.y = lo.y; }
            if (k & 4) { nr.z = hi.z; fr.z = lo.z; }
            B.nodes_oct[k * B.nodes.size() + 2 * j] = nr;
            B.nodes_oct[k * B.nodes.size() + 2 * j + 1] = fr;
        }
    build_wide(B);
    build_grid(B);
}

template <bool kCount>
static inline void hh_closest(const HostBvh& B, f3 o, f3 d, float& t, int& prim, TraceCounters& cnt) {
    if (g_use_grid && B.grid_ok) closest_hit_grid<kCount>(B.grid, B.grid_start.data(), B.grid_refs.data(), B.geom.data(), o, d, t, prim, cnt);
    else if (g_use_wide == 2 && B.wide_levels <= kWideGlobalMaxLevels && B.huge.n == 0) closest_hit_wide_global<kCount>(B.wide.data(), B.geom.data(), B.wide_root, o, d, t, prim, cnt, kTMax, -1, g_gate != 0);
    else if (g_use_wide && B.wide_levels <= kWideMaxLevels && !B.wide_oct.empty()) {
        float t0 = kTMax; int prim0 = -1;
        const float a = dot(d, d), inv_a = rcp(a);
        for (uint32_t i = 0; i < B.huge.n; i++) {
            const node_f4 g = B.geom[B.huge.idx[i]];
            if (kCount) cnt.spheres += 1;
            const float th = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, t0);
            if (th >= 0.0f) { t0 = th; prim0 = (int)B.huge.idx[i]; }
        }
        closest_hit_wide<kCount>(B.wide_oct.data(), (uint32_t)(B.wide_oct.size() / 8), B.geom.data(), B.wide_root, o, d, t, prim, cnt, t0, prim0);
    }
    else if (g_use_oct) closest_hit<kCount, true>(B.nodes_oct.data(), B.geom.data(), B.root_link, o, d, t, prim, cnt, (uint32_t)B.nodes.size(), g_gate != 0);
    else closest_hit<kCount, false>(B.nodes.data(), B.geom.data(), B.root_link, o, d, t, prim, cnt, 0u, g_gate != 0);
}

extern "C" {

void hh_set_oct(int on) { g_use_oct = on; }
void hh_set_gate(int on) { g_gate = on; }
void hh_set_wide(int on) { g_use_wide = on; }
void hh_set_grid(int on) { g_use_grid = on; }
// huge list of the host build (sorted sphere indices)
uint32_t hh_huge_list(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, uint32_t* idx8) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    for (uint32_t i = 0; i < B.huge.n; i++) idx8[i] = B.huge.idx[i];
    return B.huge.n;
}
// grid of the host build: header (80 bytes), start[n_cells + 1], refs[n_refs]; returns 1 when the scene suits the structure
int hh_build_grid(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, void* header_out, uint16_t* start_out, uint64_t cap_start,
                  uint16_t* refs_out, uint64_t cap_refs) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    if (!B.grid_ok) return 0;
    if (header_out) memcpy(header_out, &B.grid, sizeof(GridHeader));
    if (start_out) memcpy(start_out, B.grid_start.data(), std::min<uint64_t>(cap_start, B.grid_start.size()) * 2);
    if (refs_out) memcpy(refs_out, B.grid_refs.data(), std::min<uint64_t>(cap_refs, B.grid_refs.size()) * 2);
    return 1;
}
// canonical 4-wide nodes of the host build (8 float4 each); returns their number, *levels = breadth-first levels
uint64_t hh_build_wide(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, float* wide_out, uint64_t cap_nodes, uint32_t* levels) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    const uint64_t W = B.wide.size() / 8;
    if (wide_out) memcpy(wide_out, B.wide.data(), std::min<uint64_t>(W, cap_nodes) * 128);
    if (levels) *levels = B.wide_levels;
    return W;
}
void hh_set_sah_max(uint32_t n) { g_sah_max = n; }
uint32_t hh_tea4(uint32_t a, uint32_t b) { return tea4(a, b); }
uint32_t hh_lcg(uint32_t* s) { return lcg(*s); }
float hh_rnd(uint32_t* s) { return rnd(*s); }
uint32_t hh_morton30(float x, float y, float z, const float* clo, const float* cinv) { return morton30(x, y, z, clo, cinv); }
void hh_make_color(const float* rgb, uint8_t* out) { uint32_t c = make_color_u32(mk3(rgb[0], rgb[1], rgb[2])); memcpy(out, &c, 4); }

// Builds the packed LBVH on the CPU with the product's per-element functions.  Returns the node count; copies out
// up to cap nodes (32 B each) and the sorted->original permutation.
uint64_t hh_build_bvh(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, void* nodes_out, uint64_t cap,
                      uint32_t* orig_out, uint32_t* codes_out) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    uint64_t nn = B.nodes.size() / 2;
    if (nodes_out) memcpy(nodes_out, B.nodes.data(), std::min<uint64_t>(nn, cap) * 32);
    if (orig_out) memcpy(orig_out, B.orig.data(), 4ull * n);
    if (codes_out) memcpy(codes_out, B.codes.data(), 4ull * n);
    return nn;
}

void hh_closest_hit(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, const float* o, const float* d,
                    uint64_t nrays, float* t_out, int32_t* prim_out, uint64_t* node_visits, uint64_t* sphere_tests) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    TraceCounters cnt{0, 0};
    uint64_t nv = 0, st = 0;
    for (uint64_t i = 0; i < nrays; i++) {
        float t; int prim;
        cnt.nodes = cnt.spheres = 0;
        hh_closest<true>(B, mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), t, prim, cnt);
        nv += cnt.nodes; st += cnt.spheres;
        t_out[i] = prim >= 0 ? t : -1.0f;
        prim_out[i] = prim >= 0 ? (int32_t)B.orig[prim] : -1;
    }
    if (node_visits) *node_visits = nv;
    if (sphere_tests) *sphere_tests = st;
}

// pixel_color / spp for every pixel, forward throughput order (what the kernels compute).
void hh_render_mean(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, const hh_params* P,
                    float* mean_rgba, uint64_t* segments, uint64_t* node_visits, uint64_t* sphere_tests) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    SceneView sc{B.nodes.data(), B.geom.data(), B.mat.data(), B.type.data(), B.root_link};
    Camera cam;
    cam.origin = mk3(P->origin[0], P->origin[1], P->origin[2]);
    cam.u = mk3(P->u[0], P->u[1], P->u[2]);
    cam.v = mk3(P->v[0], P->v[1], P->v[2]);
    cam.w = mk3(P->w[0], P->w[1], P->w[2]);
    cam.u_unit = normalize(cam.u);
    cam.v_unit = normalize(cam.v);
    cam.lens_radius = P->lens_radius;
    cam.wm1 = (float)(P->width - 1); cam.hm1 = (float)(P->height - 1);
    cam.inv_wm1 = 1.0f / cam.wm1; cam.inv_hm1 = 1.0f / cam.hm1; cam.div_exact = (div_by_const_ok(cam.wm1) && div_by_const_ok(cam.hm1)) ? 1u : 0u;
    uint64_t segs = 0, nv = 0, stt = 0;
    const float inv_spp = 1.0f / (float)P->spp;
    for (uint32_t px = 0; px < P->width * P->height; px++) {
        uint32_t seed = tea4(px, P->subframe_index);
        f3 sum = mk3(0.0f);
        for (uint32_t k = 0; k < P->spp; k++) {
            PathState st;
            camera_ray(cam, px % P->width, px / P->width, seed, st.o, st.d);
            st.thr = mk3(1.0f);
            st.seed = seed;
            st.depth = (int)P->max_depth - 1;
            f3 result;
            while (true) {
                float t; int prim;
                TraceCounters cnt{0, 0};
                hh_closest<true>(B, st.o, st.d, t, prim, cnt);
                segs++; nv += cnt.nodes; stt += cnt.spheres;
                if (!shade_segment(sc, st, t, prim, result)) break;
            }
            sum = sum + result;
        }
        f3 m = sum * inv_spp;
        mean_rgba[4 * px] = m.x; mean_rgba[4 * px + 1] = m.y; mean_rgba[4 * px + 2] = m.z; mean_rgba[4 * px + 3] = 1.0f;
    }
    if (segments) *segments = segs;
    if (node_visits) *node_visits = nv;
    if (sphere_tests) *sphere_tests = stt;
}

}  // extern "C"

// One scatter event with the product's exact math on the CPU (same packing as k_scatter in path_kernels.cu).
extern "C" void hh_scatter(uint32_t type, const float* mat4, const float* dirs, const float* normals, const uint8_t* front,
                           const uint32_t* seeds, uint64_t n, float* dirs_out, uint8_t* scattered, uint32_t* seeds_out) {
    for (uint64_t i = 0; i < n; i++) {
        const f3 d = mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        const f3 nrm = mk3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        uint32_t seed = seeds[i];
        f3 out = mk3(0.0f);
        bool ok = true;
        if (type == 0u) out = scatter_lambertian(nrm, random_in_unit_sphere(seed));
        else if (type == 1u) ok = scatter_metal(normalize(d), nrm, mat4[3], random_in_unit_sphere(seed), out);
        else out = scatter_dielectric(normalize(d), nrm, front[i] != 0, mat4[3], seed);
        dirs_out[3 * i] = out.x; dirs_out[3 * i + 1] = out.y; dirs_out[3 * i + 2] = out.z;
        scattered[i] = ok ? 1 : 0;
        seeds_out[i] = seed;
    }
}

// Histogram of BVH node-pair visits and sphere tests per ray segment (for scheduling studies; tools/simt_model.py).
extern "C" void hh_visit_histogram(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, const hh_params* P,
                                   uint64_t* node_hist, uint64_t* sphere_hist, uint32_t bins) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    SceneView sc{B.nodes.data(), B.geom.data(), B.mat.data(), B.type.data(), B.root_link};
    Camera cam;
    cam.origin = mk3(P->origin[0], P->origin[1], P->origin[2]);
    cam.u = mk3(P->u[0], P->u[1], P->u[2]); cam.v = mk3(P->v[0], P->v[1], P->v[2]); cam.w = mk3(P->w[0], P->w[1], P->w[2]);
    cam.u_unit = normalize(cam.u); cam.v_unit = normalize(cam.v);
    cam.lens_radius = P->lens_radius;
    cam.wm1 = (float)(P->width - 1); cam.hm1 = (float)(P->height - 1);
    cam.inv_wm1 = 1.0f / cam.wm1; cam.inv_hm1 = 1.0f / cam.hm1; cam.div_exact = (div_by_const_ok(cam.wm1) && div_by_const_ok(cam.hm1)) ? 1u : 0u;
    for (uint32_t px = 0; px < P->width * P->height; px++) {
        uint32_t seed = tea4(px, P->subframe_index);
        for (uint32_t k = 0; k < P->spp; k++) {
            PathState st;
            camera_ray(cam, px % P->width, px / P->width, seed, st.o, st.d);
            st.thr = mk3(1.0f); st.seed = seed; st.depth = (int)P->max_depth - 1;
            f3 result;
            while (true) {
                float t; int prim;
                TraceCounters cnt{0, 0};
                closest_hit<true>(sc.nodes, sc.geom, sc.root_link, st.o, st.d, t, prim, cnt);
                node_hist[std::min(cnt.nodes, bins - 1)]++;
                sphere_hist[std::min(cnt.spheres, bins - 1)]++;
                if (!shade_segment(sc, st, t, prim, result)) break;
            }
        }
    }
}

// Per-ray step sequences (0 = node-pair step, k>0 = leaf step testing k spheres, 255 = end of ray), for the SIMT
// scheduling model in tools/simt_model.py.  Same visiting order as closest_hit().
static void trace_sequence(const SceneView& sc, f3 o, f3 d, std::vector<uint8_t>& out) {
    float tbest = kTMax;
    const f3 idir = slab_idir(d);
    const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const float a = dot(d, d), inv_a = rcp(a);
    uint32_t stack[kStackSize]; int sp = 0; uint32_t cur = sc.root_link;
    while (cur != kEmptyScene) {
        if (!(cur & kLeafFlag)) {
            out.push_back(0);
            const node_f4 l0 = sc.nodes[2 * cur], l1 = sc.nodes[2 * cur + 1], r0 = sc.nodes[2 * cur + 2], r1 = sc.nodes[2 * cur + 3];
            float tl, tr;
            const bool hl = box_hit(l0, l1, idir, ood, tbest, tl), hr = box_hit(r0, r1, idir, ood, tbest, tr);
            const uint32_t ll = f2u(l0.w), lr = f2u(r0.w);
            if (hl && hr) { const bool lf = tl <= tr; cur = lf ? ll : lr; stack[sp++] = lf ? lr : ll; }
            else if (hl) cur = ll; else if (hr) cur = lr; else cur = sp ? stack[--sp] : kEmptyScene;
        } else {
            const uint32_t first = (cur & 0x7FFFFFFFu) >> 3, count = (cur & 7u) + 1u;
            out.push_back((uint8_t)count);
            for (uint32_t k = 0; k < count; k++) {
                const node_f4 g = sc.geom[first + k];
                const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                if (t >= 0.0f) tbest = t;
            }
            cur = sp ? stack[--sp] : kEmptyScene;
        }
    }
    out.push_back(255);
}

static void trace_sequence_pp(const SceneView& sc, f3 o, f3 d, std::vector<uint8_t>& out);
// Same for the 4-wide octant-sorted nodes (closest_hit_wide): token 0 = one wide-node step.
static void trace_sequence_wide(const HostBvh& B, f3 o, f3 d, std::vector<uint8_t>& out) {
    float tbest = kTMax;
    const f3 idir = slab_idir(d);
    const node_f4* wn = B.wide_oct.data() + ray_octant(d) * (B.wide_oct.size() / 8);
    const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const float a = dot(d, d), inv_a = rcp(a);
    uint32_t stack[kStackSize]; int sp = 0; uint32_t cur = B.wide_root;
    while (cur != kEmptyScene) {
        if (!(cur & kLeafFlag)) {
            out.push_back(0);
            const node_f4* p = wn + kWideNodeF4 * cur;
            const node_f4 nx = p[0], ny = p[1], nz = p[2], fx = p[3], fy = p[4], fz = p[5], lk = p[6];
            if (slab_hit(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, idir, ood, tbest)) stack[sp++] = f2u(lk.w);
            if (slab_hit(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, idir, ood, tbest)) stack[sp++] = f2u(lk.z);
            if (slab_hit(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, idir, ood, tbest)) stack[sp++] = f2u(lk.y);
            if (slab_hit(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, idir, ood, tbest)) stack[sp++] = f2u(lk.x);
            cur = sp ? stack[--sp] : kEmptyScene;
        } else {
            const uint32_t first = (cur & 0x7FFFFFFFu) >> 3, count = (cur & 7u) + 1u;
            out.push_back((uint8_t)count);
            for (uint32_t k = 0; k < count; k++) {
                const node_f4 g = B.geom[first + k];
                const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
                if (t >= 0.0f) tbest = t;
            }
            cur = sp ? stack[--sp] : kEmptyScene;
        }
    }
    out.push_back(255);
}
extern "C" uint64_t hh_step_sequences(const hh_sphere* s, uint32_t n, uint32_t leaf_size, float pad_rel, const hh_params* P,
                                      uint8_t* out, uint64_t cap) {
    HostBvh B;
    build(s, n, leaf_size, pad_rel, B);
    SceneView sc{B.nodes.data(), B.geom.data(), B.mat.data(), B.type.data(), B.root_link};
    Camera cam;
    cam.origin = mk3(P->origin[0], P->origin[1], P->origin[2]);
    cam.u = mk3(P->u[0], P->u[1], P->u[2]); cam.v = mk3(P->v[0], P->v[1], P->v[2]); cam.w = mk3(P->w[0], P->w[1], P->w[2]);
    cam.u_unit = normalize(cam.u); cam.v_unit = normalize(cam.v);
    cam.lens_radius = P->lens_radius;
    cam.wm1 = (float)(P->width - 1); cam.hm1 = (float)(P->height - 1);
    cam.inv_wm1 = 1.0f / cam.wm1; cam.inv_hm1 = 1.0f / cam.hm1; cam.div_exact = (div_by_const_ok(cam.wm1) && div_by_const_ok(cam.hm1)) ? 1u : 0u;
    std::vector<uint8_t> seq;
    // pixel-major like the persistent kernel: each pixel's samples/segments are consecutive; 254 separates pixels
    for (uint32_t px = 0; px < P->width * P->height; px++) {
        uint32_t seed = tea4(px, P->subframe_index);
        for (uint32_t k = 0; k < P->spp; k++) {
            PathState st;
            seq.push_back(253);   // a new path starts (camera ray)
            camera_ray(cam, px % P->width, px / P->width, seed, st.o, st.d);
            st.thr = mk3(1.0f); st.seed = seed; st.depth = (int)P->max_depth - 1;
            f3 result;
            while (true) {
                if (g_use_wide == 1 && B.wide_levels <= kWideMaxLevels && !B.wide_oct.empty()) trace_sequence_wide(B, st.o, st.d, seq);
                else if (g_seq_postpone) trace_sequence_pp(sc, st.o, st.d, seq); else trace_sequence(sc, st.o, st.d, seq);
                float t; int prim; TraceCounters cnt{0, 0};
                closest_hit<false>(sc.nodes, sc.geom, sc.root_link, st.o, st.d, t, prim, cnt);
                if (!shade_segment(sc, st, t, prim, result)) break;
            }
        }
        seq.push_back(254);
    }
    const uint64_t m = std::min<uint64_t>(cap, seq.size());
    if (out) memcpy(out, seq.data(), m);
    return seq.size();
}

// Step sequence under the "postponed leaf" schedule: the first leaf a lane meets is parked and the lane keeps
// descending; leaves are tested when a second one turns up or the stack runs dry (model input only).
static void trace_sequence_pp(const SceneView& sc, f3 o, f3 d, std::vector<uint8_t>& out) {
    float tbest = kTMax;
    const f3 idir = slab_idir(d);
    const f3 ood = mk3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    const float a = dot(d, d), inv_a = rcp(a);
    uint32_t stack[kStackSize]; int sp = 0; uint32_t cur = sc.root_link, pend = kEmptyScene;
    auto leaf = [&](uint32_t l) {
        const uint32_t first = (l & 0x7FFFFFFFu) >> 3, count = (l & 7u) + 1u;
        out.push_back((uint8_t)count);
        for (uint32_t k = 0; k < count; k++) {
            const node_f4 g = sc.geom[first + k];
            const float t = sphere_root(o, d, a, inv_a, g.x, g.y, g.z, g.w, kTMin, tbest);
            if (t >= 0.0f) tbest = t;
        }
    };
    for (;;) {
        while (!(cur & kLeafFlag)) {
            out.push_back(0);
            const node_f4 l0 = sc.nodes[2 * cur], l1 = sc.nodes[2 * cur + 1], r0 = sc.nodes[2 * cur + 2], r1 = sc.nodes[2 * cur + 3];
            float tl, tr;
            const bool hl = box_hit(l0, l1, idir, ood, tbest, tl), hr = box_hit(r0, r1, idir, ood, tbest, tr);
            const uint32_t ll = f2u(l0.w), lr = f2u(r0.w);
            if (hl && hr) { const bool lf = tl <= tr; cur = lf ? ll : lr; stack[sp++] = lf ? lr : ll; }
            else if (hl) cur = ll; else if (hr) cur = lr; else cur = sp ? stack[--sp] : kEmptyScene;
            if ((cur & kLeafFlag) && cur != kEmptyScene && pend == kEmptyScene) { pend = cur; cur = sp ? stack[--sp] : kEmptyScene; }
        }
        if (pend != kEmptyScene) { leaf(pend); pend = kEmptyScene; }
        if (cur != kEmptyScene) { leaf(cur); cur = sp ? stack[--sp] : kEmptyScene; }
        else break;
    }
    out.push_back(255);
}
extern "C" void hh_set_seq_postpone(int on) { g_seq_postpone = on; }

// div_by_const() against the IEEE quotient for the jitter expression 2 * (px + k * 2^-24) / b, every px in [0, b] and `per_px`
// jitters each (edge values + a LCG sweep); returns the number of mismatches.  For divisors that fail div_by_const_ok() the
// caller expects to see mismatches -- that is what the guard is for.
extern "C" uint64_t hh_check_div_by_const(uint32_t b_int, uint32_t per_px) {
    const float b = (float)b_int, rb = 1.0f / b;
    uint64_t bad = 0;
    uint32_t s = 12345u + b_int;
    for (uint32_t px = 0; px <= b_int; px++)
        for (uint32_t j = 0; j < per_px; j++) {
            uint32_t k;
            if (j == 0) k = 0; else if (j == 1) k = 1; else if (j == 2) k = 0xFFFFFFu; else if (j == 3) k = 0x800000u; else k = lcg(s);
            const float a = 2.0f * ((float)px + (float)k * 5.9604644775390625e-8f);
            bad += f2u(div_by_const(a, b, rb)) != f2u(a / b);
        }
    return bad;
}

// rnd_pm1() of vn_math.cuh (three instructions on the GPU) against the literal -1 + 2*rnd form of RayTracer.cu:93-97, for every
// 24-bit LCG output and a spread of upper bytes; returns the number of mismatching bit patterns (must be 0).
extern "C" uint64_t hh_check_rnd_pm1() {
    uint64_t bad = 0;
    for (uint32_t hi = 0; hi < 256u; hi += 51u)
        for (uint32_t n = 0; n < (1u << 24); n++) {
            const uint32_t state = (hi << 24) | n;                      // the state AFTER the LCG step
            const uint32_t prev = (state - 1013904223u) * 4276115653u;    // 1664525^-1 mod 2^32
            uint32_t s1 = prev, s2 = prev;
            const float a = rnd_pm1(s1), b = rnd_pm1_literal(s2);
            bad += (f2u(a) != f2u(b)) || (s1 != s2) || (s1 != state);
        }
    return bad;
}

// Node / sphere visit counts for an externally built tree in the packed layout (BVH-quality studies).
extern "C" void hh_visits_custom(const node_f4* nodes, uint32_t n_nodes, const hh_sphere* sorted, uint32_t n, const hh_params* P,
                                 uint64_t* segs_out, uint64_t* nodes_out, uint64_t* spheres_out) {
    std::vector<node_f4> geom(n), mat(n);
    std::vector<uint8_t> type(n);
    for (uint32_t i = 0; i < n; i++) {
        geom[i] = node_f4{sorted[i].cx, sorted[i].cy, sorted[i].cz, sorted[i].r};
        mat[i] = sorted[i].type == 2 ? dielectric_mat(sorted[i].fuzz_or_ir) : node_f4{sorted[i].ax, sorted[i].ay, sorted[i].az, sorted[i].fuzz_or_ir};
        type[i] = (uint8_t)sorted[i].type;
    }
    (void)n_nodes;
    SceneView sc{nodes, geom.data(), mat.data(), type.data(), f2u(nodes[2].w)};
    Camera cam;
    cam.origin = mk3(P->origin[0], P->origin[1], P->origin[2]);
    cam.u = mk3(P->u[0], P->u[1], P->u[2]); cam.v = mk3(P->v[0], P->v[1], P->v[2]); cam.w = mk3(P->w[0], P->w[1], P->w[2]);
    cam.u_unit = normalize(cam.u); cam.v_unit = normalize(cam.v);
    cam.lens_radius = P->lens_radius;
    cam.wm1 = (float)(P->width - 1); cam.hm1 = (float)(P->height - 1);
    cam.inv_wm1 = 1.0f / cam.wm1; cam.inv_hm1 = 1.0f / cam.hm1; cam.div_exact = (div_by_const_ok(cam.wm1) && div_by_const_ok(cam.hm1)) ? 1u : 0u;
    uint64_t segs = 0, nv = 0, sv = 0;
    for (uint32_t px = 0; px < P->width * P->height; px++) {
        uint32_t seed = tea4(px, P->subframe_index);
        for (uint32_t k = 0; k < P->spp; k++) {
            PathState st;
            camera_ray(cam, px % P->width, px / P->width, seed, st.o, st.d);
            st.thr = mk3(1.0f); st.seed = seed; st.depth = (int)P->max_depth - 1;
            f3 result;
            while (true) {
                float t; int prim; TraceCounters cnt{0, 0};
                closest_hit<true>(sc.nodes, sc.geom, sc.root_link, st.o, st.d, t, prim, cnt);
                segs++; nv += cnt.nodes; sv += cnt.spheres;
                if (!shade_segment(sc, st, t, prim, result)) break;
            }
        }
    }
    *segs_out = segs; *nodes_out = nv; *spheres_out = sv;
}
