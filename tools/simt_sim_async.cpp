// tools/simt_sim_async.cpp -- SIMT schedule simulator with REAL rays (model input for kernel design, not product code).
// Unlike simt_sim_phases.cpp (which replays recorded token streams) this one traces the RTIOW workload itself with the host
// build of the product's traversal / shading code (tests/host_harness.cpp) warp by warp, lane by lane, so schedules that
// change WHEN a lane shades or traverses (asynchronous shading, postponed stragglers) can be costed before any CUDA is written.
// Cost = warp instructions of each block that at least one lane executes (SASS counts of the shipped kernel, see the constants).
//
//   g++ -O2 -std=c++17 -ffp-contract=off -w -Ivenusaur_b200/csrc -o /tmp/simt/sim_async tools/simt_sim_async.cpp
//   /tmp/simt/sim_async /tmp/simt/rtiow.bin /tmp/simt/cam_1920.bin 1920 1080 16 <n_warps> <policy> [Tn Tl Td]
// (scene = the hh_sphere array of the RTIOW scene, camera = 13 floats origin|u|v|w|lens: dump them with tests/oracle_lib.py.)
// Policies: 0 = the persistent kernel's rounds; 1 = one voted operation per quantum; 2 = node bursts (Tn) + leaf votes (Tl) + early
// shading (Td); Tn 1, Tl 33 = the phase form of k_render_async; 3 / 4 = greedy fine-grained operations with one ray per lane.
// Environment: PHASES=1 per-phase node / leaf statistics; HIST=1 node steps per ray; TNSTACK=1 stack entries carry their entry
// distance and are culled at the pop; C_NODE / C_SCHED / THR override costs and thresholds; FULL=1 the whole frame with n_warps
// resident warps (4736 = 148 SMs x 32) instead of a strided sample; DUMP_COST=file per-tile ray segments (sum | max | dielectric
// hits); TILE_ORDER=file a uint32 tile order for the tickets (how the cost-ordered schedule of vn_api.cu::prepare_tile_order was
// checked before it was built: row-major 54.1 -> cost-sorted 49.2 model warp-instructions per segment; measured 5.72 -> 5.24 G).
#include "../tests/host_harness.cpp"

#include <cstdio>
#include <cstdlib>

static int C_FETCH = 60, C_CAM = 85, C_DISK = 18, C_SETUP = 35, C_HUGE = 65, C_NODE = 83, C_LEAF = 78, C_NORM = 32, C_MISS = 16,
           C_HIT = 42, C_TRIAL = 27, C_LAMB = 50, C_METAL = 34, C_DIEL = 120, C_LOOP = 3, C_SCHED = 10, C_SHADE_FIX = 20;

static int g_thr[9] = {0, 1, 1, 1, 0, 1, 1, 1, 1};
enum Phase { NEED = 0, NODE = 1, LEAF = 2, DONE = 3, RETIRED = 4, P_CAM = 5, P_OPQ = 6, P_DIEL = 7, P_START = 8 };

static int g_tnstack = 0;
static inline float slab_tn(float nx, float ny, float nz, float fx, float fy, float fz, f3 idir, f3 ood, float tbest, bool& hit) {
    const float tn = fmaxf(fmaxf(fmaf(nx, idir.x, -ood.x), fmaf(ny, idir.y, -ood.y)), fmaxf(fmaf(nz, idir.z, -ood.z), 0.0f));
    const float tf = fminf(fminf(fmaf(fx, idir.x, -ood.x), fmaf(fy, idir.y, -ood.y)), fminf(fmaf(fz, idir.z, -ood.z), tbest));
    hit = tn <= tf; return tn;
}
struct Lane {
    Phase ph = NEED;
    bool has_pixel = false, active = false, first = true;
    uint32_t px = 0, py = 0, pix = 0, cam_seed = 0, s_left = 0;
    PathState st;
    uint32_t cur = 0, stack[kStackSize];
    int sp = 0;
    float tbest = 0, a = 0;
    int prim = -1;
    f3 idir, ood;
    const node_f4* wn = nullptr;
    int nsteps = 0; bool primary = false;
    uint32_t cur_tile = 0, seg_pixel = 0, diel_pixel = 0; bool cur_tile_valid = false;
    float tstack[kStackSize];
};

struct Sim {
    HostBvh B;
    Camera cam;
    uint32_t W, H, spp, subframe = 1, max_depth = 50;
    uint32_t tiles_x, tiles_y, n_tiles, next_ticket = 0, tile_stride = 1, max_tickets;
    SceneView sc;
    double cost = 0, c_node = 0, c_leaf = 0, c_shade = 0, c_sched = 0;
    uint64_t hist[2][64] = {{0}};
    uint64_t ph_node[8] = {0}, ph_lanes[8] = {0}, ph_leaf[8] = {0}, ph_leaf_lanes[8] = {0}, rounds = 0;
    uint64_t segs = 0, n_node_ops = 0, n_leaf_ops = 0, n_shade_ops = 0, lanes_node = 0, lanes_leaf = 0, lanes_shade = 0;

    std::vector<uint32_t> tile_order, tile_cost, tile_max, tile_diel;
    void close_pixel(Lane& L) {
        if (!L.cur_tile_valid) return;
        if (tile_cost.size()) { tile_cost[L.cur_tile] += L.seg_pixel; tile_max[L.cur_tile] = std::max(tile_max[L.cur_tile], L.seg_pixel); tile_diel[L.cur_tile] += L.diel_pixel; }
        L.cur_tile_valid = false;
    }
    // WARPTILE=1: a warp takes whole tiles (one global ticket per tile) and its lanes take the tile's pixels in order, so the lanes
    // of a warp always hold pixels of the same one or two tiles
    int warptile = 0;
    std::vector<uint32_t> w_tile, w_cursor;
    uint32_t next_tile = 0;
    bool fetch_warptile(Lane& L, int w) {
        close_pixel(L);
        for (;;) {
            if (w_cursor[w] >= 32u) {
                if (next_tile >= n_tiles) return false;
                const uint32_t t = next_tile++;
                w_tile[w] = tile_order.size() ? tile_order[t] : t;
                w_cursor[w] = 0;
            }
            const uint32_t tile = w_tile[w], in = w_cursor[w]++;
            L.cur_tile = tile; L.cur_tile_valid = true; L.seg_pixel = 0; L.diel_pixel = 0;
            const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
            L.px = tx * tile_w + (in % tile_w);
            L.py = ty * (32u / tile_w) + (in / tile_w);
            if (L.px < W && L.py < H) return true;
        }
    }
    uint32_t tile_w = 8;        // TILE_W: tile shape tile_w x (32 / tile_w), WARPTILE mode only
    int cur_warp = 0;
    bool fetch(Lane& L) {
        if (warptile) return fetch_warptile(L, cur_warp);
        close_pixel(L);
        for (;;) {
            if (next_ticket >= max_tickets) return false;
            const uint32_t w = next_ticket++;
            uint32_t tile = (uint32_t)(((uint64_t)(w >> 5) * tile_stride) % n_tiles);
            const uint32_t in = w & 31u;
            if (tile_order.size()) tile = tile_order[w >> 5];
            L.cur_tile = tile; L.cur_tile_valid = true; L.seg_pixel = 0; L.diel_pixel = 0;
            const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
            L.px = tx * 8u + (in & 7u);
            L.py = ty * 4u + (in >> 3);
            if (L.px < W && L.py < H) return true;
        }
    }
    static uint32_t draws_between(uint32_t a, uint32_t b) {
        uint32_t n = 0;
        while (a != b && n < 4096) { lcg(a); n++; }
        return n;
    }
    void start_ray(Lane& L) {
        const f3 o = L.st.o, d = L.st.d;
        L.tbest = kTMax; L.prim = -1;
        L.a = dot(d, d);
        const float inv_a = rcp(L.a);
        for (uint32_t i = 0; i < B.huge.n; i++) {
            const node_f4 g = B.geom[B.huge.idx[i]];
            const float th = sphere_root(o, d, L.a, inv_a, g.x, g.y, g.z, g.w, kTMin, L.tbest);
            if (th >= 0.0f) { L.tbest = th; L.prim = (int)B.huge.idx[i]; }
        }
        L.idir = slab_idir(d);
        L.wn = B.wide_oct.data() + ray_octant(d) * (B.wide_oct.size() / 8);
        L.ood = mk3(o.x * L.idir.x, o.y * L.idir.y, o.z * L.idir.z);
        if (L.nsteps >= 0 && !L.first) { hist[L.primary ? 0 : 1][std::min(L.nsteps, 63)]++; }
        L.first = false; L.nsteps = 0; L.primary = L.st.depth == (int)max_depth - 1;
        L.sp = 0;
        L.cur = B.wide_root;
        L.ph = (L.cur == kEmptyScene) ? DONE : ((L.cur & kLeafFlag) ? LEAF : NODE);
    }
    void set_phase(Lane& L) { L.ph = (L.cur == kEmptyScene) ? DONE : ((L.cur & kLeafFlag) ? LEAF : NODE); }

    // shade + regen + start for the lanes in DONE / NEED; returns the warp cost of the phase
    double shade_phase(Lane* w, bool only_ready = true) {
        bool any = false, any_miss = false, any_hit = false, any_opq = false, any_lamb = false, any_metal = false, any_diel = false, any_regen = false,
             any_fetch = false, any_start = false;
        uint32_t max_trials = 0, max_disk = 0;
        int n = 0;
        for (int l = 0; l < 32; l++) {
            Lane& L = w[l];
            if (L.ph != DONE && L.ph != NEED) continue;
            any = true; n++;
            if (L.ph == DONE) {
                segs++; L.seg_pixel++;
                if (L.prim >= 0 && B.type[L.prim] == 2u) L.diel_pixel++;
                f3 result;
                const uint32_t seed0 = L.st.seed;
                const int prim = L.prim;
                const bool cont = shade_segment(sc, L.st, L.tbest, prim, result);
                if (prim < 0) any_miss = true;
                else if (L.st.depth >= 0 && (cont || true)) {
                    const uint32_t type = B.type[prim];
                    any_hit = true;
                    const uint32_t nd = draws_between(seed0, L.st.seed);
                    if (type != 2u) { any_opq = true; max_trials = std::max(max_trials, nd / 3u); if (type == 0u) any_lamb = true; else any_metal = true; }
                    else any_diel = true;
                }
                if (!cont) L.active = false;
            }
            if (!L.active) {
                any_regen = true;
                if (L.has_pixel && L.s_left == 0u) L.has_pixel = false;
                if (!L.has_pixel) {
                    any_fetch = true;
                    if (!fetch(L)) { L.ph = RETIRED; continue; }
                    L.pix = L.py * W + L.px;
                    L.cam_seed = tea4(L.pix, subframe);
                    L.s_left = spp; L.has_pixel = true;
                }
                const uint32_t s0 = L.cam_seed;
                camera_ray(cam, L.px, L.py, L.cam_seed, L.st.o, L.st.d);
                max_disk = std::max(max_disk, (draws_between(s0, L.cam_seed) - 2u) / 2u);
                L.st.thr = mk3(1.0f); L.st.seed = L.cam_seed; L.st.depth = (int)max_depth - 1;
                L.s_left -= 1u; L.active = true;
            }
            any_start = true;
            start_ray(L);
        }
        if (!any) return 0;
        double c = C_SHADE_FIX + C_NORM;
        if (any_miss) c += C_MISS;
        if (any_hit) c += C_HIT;
        if (any_opq) c += C_TRIAL * max_trials + 8;
        if (any_lamb) c += C_LAMB;
        if (any_metal) c += C_METAL;
        if (any_diel) c += C_DIEL;
        if (any_fetch) c += C_FETCH;
        if (any_regen) c += C_CAM + C_DISK * max_disk;
        if (any_start) c += C_SETUP + (B.huge.n ? C_HUGE * B.huge.n : 0);
        n_shade_ops++; lanes_shade += n; c_shade += c;
        return c;
    }
    double node_op(Lane* w) {
        int n = 0;
        for (int l = 0; l < 32; l++) {
            Lane& L = w[l];
            if (L.ph != NODE) continue;
            n++; L.nsteps++;
            if (!g_tnstack) L.cur = wide_node_step(L.wn, L.cur, L.idir, L.ood, L.tbest, L.stack, L.sp);
            else {
                const node_f4* p = L.wn + kWideNodeF4 * L.cur;
                const node_f4 nx = p[0], ny = p[1], nz = p[2], fx = p[3], fy = p[4], fz = p[5], lk = p[6];
                bool h0, h1, h2, h3;
                const float t0 = slab_tn(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, L.idir, L.ood, L.tbest, h0);
                const float t1 = slab_tn(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, L.idir, L.ood, L.tbest, h1);
                const float t2 = slab_tn(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, L.idir, L.ood, L.tbest, h2);
                const float t3 = slab_tn(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, L.idir, L.ood, L.tbest, h3);
                const bool c3 = h3 && (h0 || h1 || h2), c2 = h2 && (h0 || h1), c1 = h1 && h0;
                if (c3) { L.tstack[L.sp] = t3; L.stack[L.sp++] = f2u(lk.w); }
                if (c2) { L.tstack[L.sp] = t2; L.stack[L.sp++] = f2u(lk.z); }
                if (c1) { L.tstack[L.sp] = t1; L.stack[L.sp++] = f2u(lk.y); }
                uint32_t next = h0 ? f2u(lk.x) : (h1 ? f2u(lk.y) : (h2 ? f2u(lk.z) : f2u(lk.w)));
                if (!(h0 || h1 || h2 || h3)) { next = kEmptyScene; while (L.sp) { --L.sp; if (L.tstack[L.sp] <= L.tbest) { next = L.stack[L.sp]; break; } } }
                L.cur = next;
            }
            set_phase(L);
        }
        n_node_ops++; lanes_node += n; c_node += C_NODE + C_LOOP;
        return C_NODE + C_LOOP;
    }
    double leaf_op(Lane* w) {
        int n = 0;
        TraceCounters cnt{0, 0};
        for (int l = 0; l < 32; l++) {
            Lane& L = w[l];
            if (L.ph != LEAF) continue;
            n++;
            if (!g_tnstack) L.cur = leaf_step<false>(B.geom.data(), L.cur, L.st.o, L.st.d, L.a, 0.0f, L.tbest, L.prim, L.stack, L.sp, cnt);
            else {
                int sp0 = 0; uint32_t dummy[4];
                leaf_step<false>(B.geom.data(), L.cur, L.st.o, L.st.d, L.a, 0.0f, L.tbest, L.prim, dummy, sp0, cnt);
                uint32_t next = kEmptyScene; while (L.sp) { --L.sp; if (L.tstack[L.sp] <= L.tbest) { next = L.stack[L.sp]; break; } }
                L.cur = next;
            }
            set_phase(L);
        }
        n_leaf_ops++; lanes_leaf += n; c_leaf += C_LEAF + C_LOOP;
        return C_LEAF + C_LOOP;
    }
    static void count(Lane* w, int& nN, int& nL, int& nD, int& nLive) {
        nN = nL = nD = nLive = 0;
        for (int l = 0; l < 32; l++) {
            const Phase p = w[l].ph;
            if (p == RETIRED) continue;
            nLive++;
            if (p == NODE) nN++; else if (p == LEAF) nL++; else nD++;
        }
    }
    // ---- policy 3: fine-grained operations (resolve / camera / opaque / dielectric / start / node / leaf), greedy by lane count
    f3 hitp[32], hitn[32]; // unused placeholders (shading is done by shade_segment in two halves below)
    double c_ops[9] = {0}; uint64_t n_ops[9] = {0}, l_ops[9] = {0};
    bool fine_step(Lane* w, const int* thr, int mode) {
        int cnt[9] = {0}, nLive = 0;
        for (int l = 0; l < 32; l++) { if (w[l].ph == RETIRED) continue; nLive++; cnt[w[l].ph == NEED ? P_CAM : w[l].ph]++; }
        if (!nLive) return false;
        cost += C_SCHED; c_sched += C_SCHED;
        // choose: the operation with the most waiting lanes, among those at/above their threshold; if none reaches it, the overall max
        static const int order[7] = {NODE, LEAF, DONE, P_START, P_OPQ, P_CAM, P_DIEL};
        int best = -1, bestn = 0;
        for (int k = 0; k < 7; k++) { const int o = order[k]; if (cnt[o] >= thr[o] && cnt[o] > bestn) { best = o; bestn = cnt[o]; } }
        if (best < 0) for (int k = 0; k < 7; k++) { const int o = order[k]; if (cnt[o] > bestn) { best = o; bestn = cnt[o]; } }
        if (mode == 1 && cnt[NODE] >= thr[NODE]) { best = NODE; bestn = cnt[NODE]; }
        double c = 0;
        if (best == NODE) { c = node_op(w); }
        else if (best == LEAF) { c = leaf_op(w); }
        else if (best == DONE) {
            // resolve: normalize, miss -> radiance (path over), hit -> frame + material fetch
            c = C_NORM + C_MISS + C_HIT + C_LOOP;
            for (int l = 0; l < 32; l++) {
                Lane& L = w[l];
                if (L.ph != DONE) continue;
                segs++;
                if (L.prim < 0 || !(L.st.depth > 0)) { f3 r; shade_segment(sc, L.st, L.tbest, L.prim, r); L.active = false; L.ph = P_CAM; }
                else L.ph = B.type[L.prim] == 2u ? P_DIEL : P_OPQ;
            }
        } else if (best == P_OPQ || best == P_DIEL) {
            uint32_t max_trials = 0; bool lamb = false, metal = false;
            for (int l = 0; l < 32; l++) {
                Lane& L = w[l];
                if (L.ph != best) continue;
                f3 r; const uint32_t s0 = L.st.seed; const uint32_t type = B.type[L.prim];
                const bool cont = shade_segment(sc, L.st, L.tbest, L.prim, r);
                if (best == P_OPQ) { max_trials = std::max(max_trials, draws_between(s0, L.st.seed) / 3u); if (type == 0u) lamb = true; else metal = true; }
                if (cont) L.ph = P_START; else { L.active = false; L.ph = P_CAM; }
            }
            c = best == P_DIEL ? C_DIEL + C_LOOP : C_TRIAL * max_trials + 8 + (lamb ? C_LAMB : 0) + (metal ? C_METAL : 0) + C_LOOP;
        } else if (best == P_CAM) {
            uint32_t max_disk = 0; bool any_fetch = false;
            for (int l = 0; l < 32; l++) {
                Lane& L = w[l];
                if (!(L.ph == P_CAM || L.ph == NEED)) continue;
                if (L.has_pixel && L.s_left == 0u) L.has_pixel = false;
                if (!L.has_pixel) {
                    any_fetch = true;
                    if (!fetch(L)) { L.ph = RETIRED; continue; }
                    L.pix = L.py * W + L.px; L.cam_seed = tea4(L.pix, subframe); L.s_left = spp; L.has_pixel = true;
                }
                const uint32_t s0 = L.cam_seed;
                camera_ray(cam, L.px, L.py, L.cam_seed, L.st.o, L.st.d);
                max_disk = std::max(max_disk, (draws_between(s0, L.cam_seed) - 2u) / 2u);
                L.st.thr = mk3(1.0f); L.st.seed = L.cam_seed; L.st.depth = (int)max_depth - 1; L.s_left -= 1u; L.active = true;
                L.ph = P_START;
            }
            c = (any_fetch ? C_FETCH : 0) + C_CAM + C_DISK * max_disk + C_LOOP;
        } else if (best == P_START) {
            for (int l = 0; l < 32; l++) { Lane& L = w[l]; if (L.ph == P_START) start_ray(L); }
            c = C_SETUP + C_HUGE * B.huge.n + C_LOOP;
        }
        cost += c; c_ops[best] += c; n_ops[best]++; l_ops[best] += bestn;
        return true;
    }
    // one scheduling quantum of a warp; returns false when all lanes are retired
    bool step_warp(Lane* w, int policy, int Tn, int Tl, int Td) {
        int nN, nL, nD, nLive;
        count(w, nN, nL, nD, nLive);
        if (!nLive) return false;
        if (policy == 3 || policy == 4) return fine_step(w, g_thr, policy == 4);
        if (policy == 0) {
            // shipped kernel: shade/regen for everybody, then while-while to completion
            cost += shade_phase(w);
            int phase = 0;
            for (;;) {
                count(w, nN, nL, nD, nLive);
                if (nN == 0 && nL == 0) break;
                while (nN) { cost += node_op(w); ph_node[std::min(phase, 7)]++; ph_lanes[std::min(phase, 7)] += nN; count(w, nN, nL, nD, nLive); }
                if (nL) { cost += leaf_op(w); ph_leaf[std::min(phase, 7)]++; ph_leaf_lanes[std::min(phase, 7)] += nL; }
                phase++;
            }
            rounds++;
            return true;
        }
        // asynchronous: one operation per quantum, chosen by thresholds
        cost += C_SCHED; c_sched += C_SCHED;
        if ((nN + nL == 0) || nD >= Td) cost += shade_phase(w);
        else if (nL > 0 && (nN == 0 || nL >= Tl)) cost += leaf_op(w);
        else if (policy == 2) {
            // node burst: keep stepping while at least Tn lanes stand on nodes (no scheduler pass in between)
            cost += node_op(w);
            for (;;) { count(w, nN, nL, nD, nLive); if (nN < Tn || nN == 0) break; cost += node_op(w); }
        } else cost += node_op(w);
        return true;
    }
};

int main(int argc, char** argv) {
    if (argc < 8) { fprintf(stderr, "usage: scene cam W H spp n_warps policy [Tn Tl Td] [tickets_per_warp]\n"); return 1; }
    Sim S;
    std::vector<hh_sphere> sph;
    { FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); sph.resize(n / sizeof(hh_sphere)); fread(sph.data(), 1, n, f); fclose(f); }
    float c[13];
    { FILE* f = fopen(argv[2], "rb"); fread(c, 4, 13, f); fclose(f); }
    S.W = atoi(argv[3]); S.H = atoi(argv[4]); S.spp = atoi(argv[5]);
    const int n_warps = atoi(argv[6]), policy = atoi(argv[7]);
    const int Tn = argc > 8 ? atoi(argv[8]) : 1, Tl = argc > 9 ? atoi(argv[9]) : 1, Td = argc > 10 ? atoi(argv[10]) : 32;
    const int per_warp = argc > 11 ? atoi(argv[11]) : 12;
    if (getenv("THR")) sscanf(getenv("THR"), "%d,%d,%d,%d,%d,%d,%d", &g_thr[NODE], &g_thr[LEAF], &g_thr[DONE], &g_thr[P_START], &g_thr[P_OPQ], &g_thr[P_CAM], &g_thr[P_DIEL]);
    if (getenv("TNSTACK")) g_tnstack = atoi(getenv("TNSTACK"));
    if (getenv("C_SCHED")) C_SCHED = atoi(getenv("C_SCHED"));
    if (getenv("C_NODE")) C_NODE = atoi(getenv("C_NODE"));
    g_use_wide = 1;
    build(sph.data(), (uint32_t)sph.size(), 1, 0.01f, S.B);
    S.sc = SceneView{S.B.nodes.data(), S.B.geom.data(), S.B.mat.data(), S.B.type.data(), S.B.root_link};
    S.cam.origin = mk3(c[0], c[1], c[2]); S.cam.u = mk3(c[3], c[4], c[5]); S.cam.v = mk3(c[6], c[7], c[8]); S.cam.w = mk3(c[9], c[10], c[11]);
    S.cam.u_unit = normalize(S.cam.u); S.cam.v_unit = normalize(S.cam.v); S.cam.lens_radius = c[12];
    S.cam.wm1 = (float)(S.W - 1); S.cam.hm1 = (float)(S.H - 1); S.cam.inv_wm1 = 1.0f / S.cam.wm1; S.cam.inv_hm1 = 1.0f / S.cam.hm1;
    S.tiles_x = (S.W + 7) / 8; S.tiles_y = (S.H + 3) / 4; S.n_tiles = S.tiles_x * S.tiles_y;
    S.tile_stride = 7919;                                   // prime: the simulated tickets sample the whole frame
    S.max_tickets = (uint32_t)n_warps * 32u * (uint32_t)per_warp;
    if (getenv("TILE_W")) { S.tile_w = atoi(getenv("TILE_W")); S.tiles_x = (S.W + S.tile_w - 1) / S.tile_w; S.tiles_y = (S.H + 32 / S.tile_w - 1) / (32 / S.tile_w); S.n_tiles = S.tiles_x * S.tiles_y; }
    if (getenv("WARPTILE")) { S.warptile = 1; S.w_tile.assign(n_warps, 0); S.w_cursor.assign(n_warps, 32u); }
    if (getenv("FULL")) { S.tile_stride = 1; S.max_tickets = S.n_tiles * 32u; }           // the whole frame, row-major unless TILE_ORDER
    if (getenv("TILE_ORDER")) { FILE* f = fopen(getenv("TILE_ORDER"), "rb"); S.tile_order.resize(S.n_tiles); fread(S.tile_order.data(), 4, S.n_tiles, f); fclose(f); }
    if (getenv("DUMP_COST")) { S.tile_cost.assign(S.n_tiles, 0); S.tile_max.assign(S.n_tiles, 0); S.tile_diel.assign(S.n_tiles, 0); }
    std::vector<Lane> lanes((size_t)n_warps * 32);
    bool any = true;
    while (any) {
        any = false;
        for (int w = 0; w < n_warps; w++) { S.cur_warp = w; any |= S.step_warp(&lanes[(size_t)w * 32], policy, Tn, Tl, Td); }
    }
    printf("policy %d Tn %d Tl %d Td %d: huge %u wide nodes %zu | segments %llu  warp-inst/seg %.2f  (node %.2f leaf %.2f shade %.2f sched %.2f) | ops/seg node %.3f leaf %.3f shade %.3f | lanes node %.1f leaf %.1f shade %.1f\n",
           policy, Tn, Tl, Td, S.B.huge.n, S.B.wide_oct.size() / 56, (unsigned long long)S.segs, S.cost / S.segs, S.c_node / S.segs, S.c_leaf / S.segs,
           S.c_shade / S.segs, S.c_sched / S.segs, (double)S.n_node_ops / S.segs, (double)S.n_leaf_ops / S.segs, (double)S.n_shade_ops / S.segs,
           (double)S.lanes_node / S.n_node_ops, (double)S.lanes_leaf / S.n_leaf_ops, (double)S.lanes_shade / S.n_shade_ops);
    if (getenv("DUMP_COST")) { for (auto& L : lanes) S.close_pixel(L); FILE* f = fopen(getenv("DUMP_COST"), "wb"); fwrite(S.tile_cost.data(), 4, S.n_tiles, f); fwrite(S.tile_max.data(), 4, S.n_tiles, f); fwrite(S.tile_diel.data(), 4, S.n_tiles, f); fclose(f); }
    if (getenv("PHASES")) for (int k = 0; k < 8; k++) printf("phase %d: node ops/round %.2f (lanes %.1f)  leaf ops/round %.2f (lanes %.1f)\n", k, (double)S.ph_node[k] / S.rounds, S.ph_node[k] ? (double)S.ph_lanes[k] / S.ph_node[k] : 0.0, (double)S.ph_leaf[k] / S.rounds, S.ph_leaf[k] ? (double)S.ph_leaf_lanes[k] / S.ph_leaf[k] : 0.0);
    if (getenv("HIST")) for (int k = 0; k < 2; k++) { uint64_t tot = 0, sum = 0; for (int i = 0; i < 64; i++) { tot += S.hist[k][i]; sum += i * S.hist[k][i]; } printf("%s rays %llu mean steps %.2f:", k ? "secondary" : "primary", (unsigned long long)tot, (double)sum / tot); double cum = 0; for (int i = 0; i < 40; i++) { cum += S.hist[k][i]; printf(" %d:%.3f", i, cum / tot); } printf("\n"); }
    if (policy >= 3) {
        const char* nm[9] = {"", "node", "leaf", "resolve", "", "camera", "opaque", "dielectric", "start"};
        for (int o = 1; o < 9; o++) if (S.n_ops[o]) printf("   %-10s cost/seg %6.2f  ops/seg %.3f  lanes %.1f\n", nm[o], S.c_ops[o] / S.segs, (double)S.n_ops[o] / S.segs, (double)S.l_ops[o] / S.n_ops[o]);
    }
    return 0;
}
