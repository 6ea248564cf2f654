"""CPU tests of the product's host side: the C ABI library loads and exports what the header declares, the drop-in
scene/camera match the oracle, and the product's device math + LBVH logic (compiled for the host by tests/host_harness.cpp)
is bit-identical to the oracle.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import venusaur_b200 as vb
from venusaur_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "venusaur_b200.h")).read()
    declared = set(re.findall(r"VN_API\s+[\w\s\*]+?\b(vn_\w+)\s*\(", header))
    assert len(declared) >= 35
    lib = vb.load()
    for name in declared:
        assert hasattr(lib, name), "libvenusaur_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    out = subprocess.run(["nm", "-D", "--defined-only", vb.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vn_\w+)", out))
    assert declared <= exported
    assert lib.vn_version().decode().startswith("venusaur_b200")


def test_abi_struct_sizes():
    assert C.sizeof(_lib.vn_sphere) == 36 and C.sizeof(_lib.vn_node32) == 32
    # the C side static_asserts the same numbers (vn_api.cu)
    assert C.sizeof(_lib.vn_params) == 96 and C.sizeof(_lib.vn_stats) == 64 and C.sizeof(_lib.vn_bvh_info) == 48
    assert _lib.vn_params.origin.offset == 32 and _lib.vn_params.flags.offset == 92


def test_no_gpu_fails_loudly():
    lib = vb.load()
    if lib.vn_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(vb.Exception, match="no CPU path"):
        vb.Context(0)
    with pytest.raises(vb.Exception):
        r = vb.Renderer()
        r.Init(vb.Scene())


def test_scene_and_camera_match_oracle(oracle_mod, rtiow):
    assert np.array_equal(vb.rtiow_final_scene().view(np.uint8), rtiow.view(np.uint8))
    sc = vb.Scene()
    assert len(sc.m_spheres) == 486 and sc.m_aabbs.shape == (486, 6) and list(sc.m_indices[:3]) == [0, 1, 2]
    for (w, h) in [(400, 225), (1920, 1080), (1200, 800)]:
        a = oracle_mod.rtiow_camera(w, h)
        b = vb.rtiow_camera(w, h).frame()
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    for n, seed, S, mix in [(1000, 0x5EED0001, 100.0, 0), (1000, 0x5EED0002, 250.0, 1)]:
        a = oracle_mod.random_scene(n, seed, S, mix)
        b = vb.random_scene(n, seed, S, mix)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    mixes = np.bincount(vb.random_scene(20000, 0x5EED0002, 250.0, 1)["type"], minlength=3) / 20000.0
    assert abs(mixes[2] - 0.5) < 0.02 and abs(mixes[0] - 0.4) < 0.02


def test_camera_dirty_flag():
    cam = vb.rtiow_camera(400, 225)
    assert cam.Changed() is True and cam.Changed() is False
    cam.SetFocalLength(10.0)
    assert cam.Changed() is False
    cam.SetFocalLength(9.0)
    assert cam.Changed() is True
    assert cam.GetLensRadius() == pytest.approx(0.05)


def _hh_params(cam, W, H, spp, sub, depth):
    class HP(C.Structure):
        _fields_ = [(n, C.c_uint32) for n in "width height spp subframe max_depth".split()] + \
                   [("origin", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3), ("lens", C.c_float)]
    p = HP()
    p.width, p.height, p.spp, p.subframe, p.max_depth = W, H, spp, sub, depth
    p.origin, p.u, p.v, p.w, p.lens = (C.c_float * 3)(*cam[0]), (C.c_float * 3)(*cam[1]), (C.c_float * 3)(*cam[2]), (C.c_float * 3)(*cam[3]), float(cam[4])
    return p


@pytest.mark.parametrize("leaf,depth,sub", [(1, 50, 1), (2, 4, 0), (2, 50, 7), (4, 50, 2), (8, 12, 3)])
def test_product_math_equals_oracle_on_cpu(host_harness, oracle_mod, rtiow, leaf, depth, sub):
    """vn_math.cuh (exact build) + the LBVH bodies of lbvh_core.cuh, run on the CPU, vs the oracle with brute-force
    closest hit: every pixel bit-identical, same number of ray segments."""
    W, H, spp = 48, 27, 3
    cam = oracle_mod.rtiow_camera(W, H)
    hp = _hh_params(cam, W, H, spp, sub, depth)
    mean = np.zeros((H, W, 4), np.float32)
    segs, nv, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
    host_harness.hh_render_mean(rtiow.ctypes.data_as(C.c_void_p), len(rtiow), leaf, C.c_float(0.01), C.byref(hp),
                                mean.ctypes.data_as(C.c_void_p), C.byref(segs), C.byref(nv), C.byref(st))
    orc = oracle_mod.Oracle(rtiow)
    want, stats = orc.render_mean(orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BRUTE))
    assert segs.value == stats.segments
    assert np.array_equal(mean, want)
    assert nv.value / segs.value < 20      # the LBVH actually prunes


@pytest.mark.parametrize("mix,depth,wide", [(1, 64, 1), (0, 50, 1), (1, 64, 0)])
def test_product_math_equals_oracle_on_glass_heavy_random_scene(host_harness, oracle_mod, mix, depth, wide):
    """The same bit-for-bit comparison away from RTIOW: 400 random spheres (mix 1 = 50 % glass, configs[4]'s material mix), seen from
    inside the cloud with odd frame sizes -- long dielectric chains (ratio > 1 exits, total internal reflection), the per-sphere
    dielectric constants, the quotient shortcuts of sphere_root and the jitter division by width - 1 = 36, height - 1 = 22."""
    W, H, spp, sub = 37, 23, 4, 5
    spheres = np.ascontiguousarray(oracle_mod.random_scene(400, 0x5EED0002, 1.2, mix))
    cam = oracle_mod.camera((0.1, 0.05, 2.6), (0.0, 0.0, -1.0), 40.0, W / H, 0.05, 2.5)
    hp = _hh_params(cam, W, H, spp, sub, depth)
    mean = np.zeros((H, W, 4), np.float32)
    segs, nv, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
    host_harness.hh_set_wide(wide)
    try:
        host_harness.hh_render_mean(spheres.ctypes.data_as(C.c_void_p), len(spheres), 1 if wide else 2, C.c_float(0.01), C.byref(hp),
                                    mean.ctypes.data_as(C.c_void_p), C.byref(segs), C.byref(nv), C.byref(st))
    finally:
        host_harness.hh_set_wide(0)
    orc = oracle_mod.Oracle(spheres)
    want, stats = orc.render_mean(orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BRUTE))
    assert segs.value == stats.segments
    assert np.array_equal(mean.view(np.uint32), want.view(np.uint32))
    assert stats.segments > 3 * W * H * spp                 # paths really bounce around in there


def _check_bvh(nodes, order, spheres, leaf_size, pad_rel=0.01):
    """Structural invariants of the packed 32-byte-node LBVH."""
    n = len(spheres)
    assert sorted(order.tolist()) == list(range(n))
    seen = np.zeros(n, np.int32)
    LEAF = 0x80000000

    def walk(idx, depth):
        nd = nodes[idx]
        link = int(nd["link"])
        if link & LEAF:
            first, cnt = (link & 0x7FFFFFFF) >> 3, (link & 7) + 1
            assert cnt <= max(leaf_size, 1) and int(nd["aux"]) == cnt
            for k in range(first, first + cnt):
                seen[k] += 1
                s = spheres[order[k]]
                r = abs(float(s["r"]))
                c = np.array([s["cx"], s["cy"], s["cz"]], np.float64)
                assert (nd["lo"] <= c - r).all() and (nd["hi"] >= c + r).all()
            return cnt, depth
        assert link % 2 == 0 and link >= 2
        tot, dmax = 0, depth
        for ch in (link, link + 1):
            c = nodes[ch]
            assert (c["lo"] >= nd["lo"]).all() and (c["hi"] <= nd["hi"]).all(), "child box outside parent"
            k, d = walk(ch, depth + 1)
            tot += k
            dmax = max(dmax, d)
        assert int(nd["aux"]) == tot
        return tot, dmax

    total, depth = walk(1, 0)
    assert total == n and (seen == 1).all()
    assert depth <= 62
    return depth


@pytest.mark.parametrize("leaf", [1, 2, 4])
def test_lbvh_logic_invariants_on_cpu(host_harness, oracle_mod, rtiow, leaf):
    import sys
    sys.setrecursionlimit(10000)
    for spheres in (rtiow, oracle_mod.random_scene(3000, 0x5EED0001, 30.0, 0), rtiow[:1], rtiow[:2], rtiow[:3],
                    np.repeat(rtiow[5:6], 40)):      # 40 identical spheres: equal Morton codes
        n = len(spheres)
        spheres = np.ascontiguousarray(spheres)
        nodes = np.zeros(2 * n + 4, vb.api.NODE_DTYPE)
        order = np.zeros(n, np.uint32)
        codes = np.zeros(n, np.uint32)
        nn = host_harness.hh_build_bvh(spheres.ctypes.data_as(C.c_void_p), n, leaf, C.c_float(0.01), nodes.ctypes.data_as(C.c_void_p), len(nodes),
                                       order.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p))
        assert nn <= 2 * n + 2
        assert (np.diff(codes.astype(np.int64)) >= 0).all() and codes.max() < (1 << 30)
        _check_bvh(nodes[:nn], order, spheres, leaf)


def test_traversal_equals_brute_force_on_cpu(host_harness, oracle_mod):
    spheres = oracle_mod.random_scene(2000, 0x5EED0001, 20.0, 0)
    rng = np.random.RandomState(11)
    o = (rng.rand(20000, 3).astype(np.float32) - 0.5) * np.float32(50.0)
    d = rng.randn(20000, 3).astype(np.float32)
    orc = oracle_mod.Oracle(spheres)
    t0, p0 = orc.closest_hit(o, d, use_bvh=False)
    t1 = np.zeros(len(o), np.float32)
    p1 = np.zeros(len(o), np.int32)
    host_harness.hh_closest_hit(spheres.ctypes.data_as(C.c_void_p), len(spheres), 2, C.c_float(0.01), o.ctypes.data_as(C.c_void_p),
                                d.ctypes.data_as(C.c_void_p), len(o), t1.ctypes.data_as(C.c_void_p), p1.ctypes.data_as(C.c_void_p), None, None)
    assert (p0 >= 0).sum() > 1000
    assert np.array_equal(t0, t1) and np.array_equal(p0, p1)


def test_cpp_dropin_headers_compile():
    """The header-only C++17 shim compiles without CUDA or glm and uses the reference's class and method names."""
    src = r'''
    #include "venusaur/Renderer.h"
    int main() {
        Scene scene;                                   // Core.cpp:30
        Camera camera(venusaur::vec3(13, 2, 3), 20.0f, 3.0f / 2.0f, 0.1f, 10.0f);   // Core.cpp:29
        camera.SetForward(venusaur::vec3(0 - 13, 0 - 2, 0 - 3));                    // Core.cpp:355
        Renderer renderer;
        if (scene.m_spheres.size() != 486 || scene.m_aabbs.size() != 486 || scene.m_indices.size() != 486) return 1;
        try { renderer.Init(scene, "ptx is ignored"); } catch (const Exception& e) { return 0; }   // no GPU here -> throws
        CUDAOutputBuffer<uchar4> buf(CUDAOutputBufferType::CUDA_DEVICE, 64, 36);
        renderer.Draw(camera, buf);
        uchar4* px = buf.getHostPointer();
        renderer.Cleanup();
        return px ? 0 : 2;
    }'''
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    cpp = os.path.join(out, "dropin.cpp")
    open(cpp, "w").write(src)
    exe = os.path.join(out, "dropin")
    libdir = os.path.dirname(vb.lib_path())
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), cpp, "-o", exe,
                    "-L" + libdir, "-lvenusaur_b200", "-Wl,-rpath," + libdir], check=True)
    assert subprocess.run([exe]).returncode == 0


def test_near_zero_float_threshold_equals_double_compare():
    """vn_math.cuh replaces near_zero's double compare (RayTracer.cu:8-13) by a float compare; they agree for every float
    around the threshold."""
    c = np.float32(9.99999993922529e-09)
    assert float(c) < 1e-8 <= float(np.nextafter(c, np.float32(1)))
    xs = np.float32(1e-8) + np.arange(-2000, 2000, dtype=np.float32) * np.float32(1e-15)
    xs = np.concatenate([xs, np.array([0, 1e-9, 1e-7, c, np.nextafter(c, np.float32(1)), np.nextafter(c, np.float32(0))], np.float32)])
    assert np.array_equal(xs.astype(np.float64) < 1e-8, xs <= c)


def test_library_staleness_is_decided_by_content_not_by_file_times(tmp_path):
    """venusaur_b200/build.py: a copy of the tree arrives with fresh file times (that made eight ranks rebuild into the same objects at
    once); whether the library is current is decided by a hash of the sources kept next to it."""
    import venusaur_b200.build as b
    assert os.path.exists(b.LIB) and os.path.exists(b.STAMP), "build the library first (python -m venusaur_b200.build)"
    assert b.up_to_date()
    src = os.path.join(b.CSRC, "vn_math.cuh")
    st = os.stat(src)
    try:
        os.utime(src, None)                                   # newer than the library: still current
        assert b.up_to_date()
        with open(b.STAMP) as f:
            good = f.read()
        with open(b.STAMP, "w") as f:
            f.write("0" * 64)                                 # a different content hash: stale
        assert not b.up_to_date()
        with open(b.STAMP, "w") as f:
            f.write(good)
        assert b.up_to_date()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))


def test_tile_index_by_multiply_high():
    """kernels.h::tile_row_col: tile / tiles_x through floor(2^32 / tiles_x) and at most two corrections, for any 32-bit tile index."""
    rng = np.random.default_rng(7)
    for d in (1, 2, 3, 5, 7, 50, 240, 480, 2048, 65535, 100000):
        inv = (2 ** 32 // d) if d > 1 else 0xFFFFFFFF
        t = np.concatenate([rng.integers(0, 2 ** 32, 200000, dtype=np.uint64), np.arange(0, 5000, dtype=np.uint64),
                            np.uint64(2 ** 32 - 1) - np.arange(0, 5000, dtype=np.uint64)])
        q = (t * np.uint64(inv)) >> np.uint64(32)
        r = t - q * np.uint64(d)
        for _ in range(2):
            m = r >= d
            q = q + m
            r = r - m * np.uint64(d)
        assert (q == t // d).all() and (r == t % d).all(), d


def test_warp_owned_tile_protocol_hands_out_every_pixel_once():
    """Model of the pixel fetch of k_render_async<.., kWarpTile> (path_kernels.cu): a warp takes whole 8x4 tiles from one global
    counter and hands their 32 slots to the asking lanes by ballot rank and a warp-uniform cursor; slots outside a ragged frame are
    skipped (the lane keeps asking); when the tiles run out the asking lanes retire.  Whatever the pattern of asking lanes, every
    in-frame pixel must be handed out exactly once and every lane must end up retired."""
    rng = np.random.default_rng(3)
    for (W, H, row_begin, row_end, n_warps) in ((33, 17, 0, 17, 3), (64, 40, 4, 29, 5), (8, 4, 0, 4, 2), (100, 7, 0, 7, 40), (5, 3, 0, 3, 1)):
        tiles_x, rows = (W + 7) // 8, row_end - row_begin
        n_tiles = tiles_x * ((rows + 3) // 4)
        counter = 0
        owner = {}
        w_tile, w_cursor = [0] * n_warps, [32] * n_warps
        busy = np.zeros((n_warps, 32), np.int64)                 # rounds a lane still works on its pixel
        retired = np.zeros((n_warps, 32), bool)
        for step in range(100000):
            if retired.all():
                break
            for w in range(n_warps):
                busy[w] = np.maximum(busy[w] - 1, 0)
                need = (busy[w] == 0) & ~retired[w]
                while need.any():                                # while (m)
                    m = need.copy()                              # the ballot the ranks and the cursor advance are taken from
                    if w_cursor[w] >= 32:
                        t = counter
                        counter += 1
                        if t >= n_tiles:
                            retired[w] |= need
                            break
                        w_tile[w], w_cursor[w] = t, 0
                    slot = w_cursor[w] + np.cumsum(m) - m        # w_cursor + popc(m & lanemask_lt)
                    for lane in np.nonzero(m & (slot < 32))[0]:
                        ty, tx = divmod(w_tile[w], tiles_x)
                        px, py = tx * 8 + (int(slot[lane]) & 7), row_begin + ty * 4 + (int(slot[lane]) >> 3)
                        if px < W and py < row_end:
                            assert (px, py) not in owner
                            owner[(px, py)] = (w, int(lane))
                            need[lane] = False
                            busy[w, lane] = int(rng.integers(1, 6))
                    w_cursor[w] = min(32, w_cursor[w] + int(m.sum()))
        assert retired.all()
        assert len(owner) == W * rows and all(row_begin <= y < row_end and 0 <= x < W for x, y in owner)


def test_rnd_pm1_short_form_equals_literal_form(host_harness):
    """vn_math.cuh::rnd_pm1 forms random_float(seed, -1, 1) (RayTracer.cu:93-97) as float(int32((state << 8) ^ 2^31)) * 2^-31; it must
    give the bits of -1 + 2 * (float(state & 0xFFFFFF) / 2^24) for every 24-bit output (and leave the same LCG state)."""
    host_harness.hh_check_rnd_pm1.restype = C.c_uint64
    assert host_harness.hh_check_rnd_pm1() == 0


def test_div_by_const_equals_ieee_division(host_harness):
    """camera_ray divides by float(width - 1) / float(height - 1) (RayTracer.cu:173-174); vn_math.cuh::div_by_const gives the same bits
    with one multiply and two fmas whenever div_by_const_ok() accepts the divisor (significand not all ones)."""
    host_harness.hh_check_div_by_const.restype = C.c_uint64
    host_harness.hh_check_div_by_const.argtypes = [C.c_uint32, C.c_uint32]
    for b in (1919, 1079, 3839, 2159, 399, 224, 1199, 799, 199, 119, 32, 16, 4, 1, 2, 5, 1000, 4096, 47, 99):
        assert host_harness.hh_check_div_by_const(b, 4096 if b > 300 else 65536) == 0, b


def test_octant_mirrored_traversal_equals_plain_on_cpu(host_harness, oracle_mod, rtiow):
    """k_render_persistent's octant-specialised node copies (near/far-plane form, no per-axis min/max) find exactly
    the same closest hits, and visit exactly as many nodes, as the plain lo/hi slab test."""
    rng = np.random.RandomState(23)
    n = 30000
    o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(30.0)
    o[:, 1] = np.abs(o[:, 1]) * np.float32(0.2) + np.float32(0.01)
    d = rng.randn(n, 3).astype(np.float32)
    d[:50, 0] = 0.0                                   # axis-parallel rays: 1/0 = inf planes
    d[50:100, 1] = -0.0
    d[100:150, 2] = 0.0
    out = {}
    try:
        for oct_ in (0, 1):
            host_harness.hh_set_oct(oct_)
            t = np.zeros(n, np.float32)
            p = np.zeros(n, np.int32)
            nv, st = C.c_uint64(), C.c_uint64()
            host_harness.hh_closest_hit(rtiow.ctypes.data_as(C.c_void_p), len(rtiow), 2, C.c_float(0.01), o.ctypes.data_as(C.c_void_p),
                                        d.ctypes.data_as(C.c_void_p), n, t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                        C.byref(nv), C.byref(st))
            out[oct_] = (t, p, nv.value, st.value)
    finally:
        host_harness.hh_set_oct(0)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert abs(out[0][2] - out[1][2]) <= 0.001 * out[0][2]      # NaN planes of axis-parallel rays may add a visit or two
    orc = oracle_mod.Oracle(rtiow)
    t0, p0 = orc.closest_hit(o, d, use_bvh=False)
    assert np.array_equal(t0, out[1][0]) and np.array_equal(p0, out[1][1])


def _host_wide(host_harness, spheres, leaf):
    n = len(spheres)
    wide = np.zeros((max(n, 1), 4, 2, 4), np.float32)
    lev = C.c_uint32()
    W = host_harness.hh_build_wide(spheres.ctypes.data_as(C.c_void_p), n, leaf, C.c_float(0.01), wide.ctypes.data_as(C.c_void_p), len(wide), C.byref(lev))
    return wide[:W], lev.value


@pytest.mark.parametrize("leaf", [1, 2, 3, 8])
def test_wide_nodes_invariants_on_cpu(host_harness, oracle_mod, rtiow, leaf):
    """The 4-wide nodes derived from the packed pairs (lbvh_core.cuh::wide_collapse, breadth-first): every sphere is in
    exactly one leaf, every wide node is referenced exactly once, child boxes are the pairs' boxes, empty slots are
    inverted boxes, and the level count is what the traversal stack was sized for."""
    for spheres in (rtiow, np.ascontiguousarray(oracle_mod.random_scene(3000, 0x5EED0001, 30.0, 0)), np.ascontiguousarray(rtiow[:5])):
        wide, levels = _host_wide(host_harness, spheres, leaf)
        n = len(spheres)
        idx8 = (C.c_uint32 * 8)()
        n_huge = host_harness.hh_huge_list(spheres.ctypes.data_as(C.c_void_p), n, leaf, C.c_float(0.01), idx8)
        huge = [int(idx8[i]) for i in range(n_huge)]
        assert n_huge == (1 if (spheres is rtiow and leaf == 1) else 0)      # the radius-1000 ground, when leaves hold one sphere
        if n <= leaf:
            assert len(wide) == 0
            continue
        links = wide[:, :, 0, 3].copy().view(np.uint32)
        seen = np.zeros(n, np.int32)
        refs = np.zeros(len(wide), np.int32)
        refs[0] = 1
        for w in range(len(wide)):
            kids = 0
            for c in range(4):
                l = int(links[w, c])
                lo, hi = wide[w, c, 0, :3], wide[w, c, 1, :3]
                if l == 0xFFFFFFFF:
                    assert (lo > 1e38).all() and (hi < -1e38).all()
                    continue
                kids += 1
                assert (lo <= hi).all()
                if l & 0x80000000:
                    first, cnt = (l & 0x7FFFFFFF) >> 3, (l & 7) + 1
                    assert cnt <= leaf
                    seen[first:first + cnt] += 1
                else:
                    assert w < l < len(wide)          # breadth-first: children come later
                    refs[l] += 1
            assert kids >= (1 if huge else 2)
        want_seen = np.ones(n, np.int32)
        want_seen[huge] = 0                   # huge spheres are tested before the traversal, not through the wide nodes
        assert np.array_equal(seen, want_seen) and (refs == 1).all()
        assert 1 <= levels <= 20
        assert len(wide) <= (n + 1) // 2 + 1


def test_wide_traversal_equals_plain_on_cpu(host_harness, oracle_mod, rtiow):
    """closest_hit_wide over the octant-sorted 4-wide nodes finds exactly the closest hits of the pair traversal and of
    brute force, with fewer than half the node steps; a whole render through it is bit-identical to the oracle."""
    rng = np.random.RandomState(29)
    n = 30000
    o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(30.0)
    o[:, 1] = np.abs(o[:, 1]) * np.float32(0.2) + np.float32(0.01)
    d = rng.randn(n, 3).astype(np.float32)
    d[:50, 0] = 0.0
    d[50:100, 1] = -0.0
    d[100:150, 2] = 0.0
    out = {}
    try:
        for wide in (0, 1):
            host_harness.hh_set_wide(wide)
            t = np.zeros(n, np.float32)
            p = np.zeros(n, np.int32)
            nv, st = C.c_uint64(), C.c_uint64()
            host_harness.hh_closest_hit(rtiow.ctypes.data_as(C.c_void_p), len(rtiow), 2, C.c_float(0.01), o.ctypes.data_as(C.c_void_p),
                                        d.ctypes.data_as(C.c_void_p), n, t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                        C.byref(nv), C.byref(st))
            out[wide] = (t, p, nv.value, st.value)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        assert out[1][2] < 0.6 * out[0][2]
        orc = oracle_mod.Oracle(rtiow)
        t0, p0 = orc.closest_hit(o, d, use_bvh=False)
        assert np.array_equal(t0, out[1][0]) and np.array_equal(p0, out[1][1])
        W, H, spp, sub, depth = 48, 27, 3, 5, 50
        cam = oracle_mod.rtiow_camera(W, H)
        hp = _hh_params(cam, W, H, spp, sub, depth)
        want, stats = orc.render_mean(orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BRUTE))
        for leaf in (1, 3):
            mean = np.zeros((H, W, 4), np.float32)
            segs, nv, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
            host_harness.hh_render_mean(rtiow.ctypes.data_as(C.c_void_p), len(rtiow), leaf, C.c_float(0.01), C.byref(hp),
                                        mean.ctypes.data_as(C.c_void_p), C.byref(segs), C.byref(nv), C.byref(st))
            assert segs.value == stats.segments and np.array_equal(mean, want)
            assert nv.value / segs.value < 8
    finally:
        host_harness.hh_set_wide(0)


def test_wide_global_traversal_equals_pairs_on_cpu(host_harness, oracle_mod):
    """closest_hit_wide_global (canonical 128-byte wide nodes, hit children sorted by entry distance -- the form large
    scenes are traversed in from L2/HBM) finds the closest hits of the pair traversal with about half the node steps."""
    spheres = np.ascontiguousarray(oracle_mod.random_scene(20000, 0x5EED0001, 30.0, 0))
    rng = np.random.RandomState(31)
    n = 20000
    o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(60.0)
    d = rng.randn(n, 3).astype(np.float32)
    d[:30, 1] = 0.0
    out = {}
    try:
        for wide in (0, 2):
            host_harness.hh_set_wide(wide)
            t = np.zeros(n, np.float32)
            p = np.zeros(n, np.int32)
            nv, st = C.c_uint64(), C.c_uint64()
            host_harness.hh_closest_hit(spheres.ctypes.data_as(C.c_void_p), len(spheres), 2, C.c_float(0.01), o.ctypes.data_as(C.c_void_p),
                                        d.ctypes.data_as(C.c_void_p), n, t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                        C.byref(nv), C.byref(st))
            out[wide] = (t, p, nv.value, st.value)
    finally:
        host_harness.hh_set_wide(0)
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
    assert (out[0][1] >= 0).sum() > 500
    assert out[2][2] < 0.6 * out[0][2] and out[2][3] < 1.05 * out[0][3]


def _host_grid(host_harness, spheres):
    """(header[26] uint32, start uint16, refs uint16) of the host emulation's grid, or None."""
    hdr = np.zeros(26, np.uint32)
    a = spheres.ctypes.data_as(C.c_void_p)
    if not host_harness.hh_build_grid(a, len(spheres), 2, C.c_float(0.01), hdr.ctypes.data_as(C.c_void_p), None, 0, None, 0):
        return None
    start = np.zeros(int(hdr[15]) + 1, np.uint16)
    refs = np.zeros(max(int(hdr[16]), 1), np.uint16)
    host_harness.hh_build_grid(a, len(spheres), 2, C.c_float(0.01), hdr.ctypes.data_as(C.c_void_p), start.ctypes.data_as(C.c_void_p), len(start),
                               refs.ctypes.data_as(C.c_void_p), len(refs))
    return hdr, start, refs[:int(hdr[16])]


def test_grid_invariants_and_traversal_on_cpu(host_harness, oracle_mod, rtiow):
    """Uniform grid + oversize list (grid_core.cuh): RTIOW puts the ground and the three radius-1 spheres in the oversize
    list and the 482 small ones in a one-layer grid; every small sphere is referenced by every cell its box touches, cell
    lists are sorted; closest hits equal brute force (axis-parallel and -0 directions included); a whole render through
    the grid is bit-identical to the oracle and needs fewer than half the traversal steps of the wide BVH."""
    hdr, start, refs = _host_grid(host_harness, rtiow)
    f = hdr.view(np.float32)
    res, n_cells, n_refs, n_big = hdr[12:15], int(hdr[15]), int(hdr[16]), int(hdr[17])
    assert n_big == 4 and int(res[1]) == 1 and n_cells == int(res[0]) * int(res[2]) and start[-1] == n_refs == len(refs)
    assert (np.diff(start.astype(np.int64)) >= 0).all()
    for c in range(n_cells):
        lst = refs[start[c]:start[c + 1]]
        assert (np.diff(lst.astype(np.int64)) > 0).all()
    assert len(np.unique(refs)) == len(rtiow) - n_big
    rng = np.random.RandomState(37)
    for spheres, S in ((rtiow, 30.0), (np.ascontiguousarray(oracle_mod.random_scene(3000, 0x5EED0001, 30.0, 0)), 60.0)):
        n = 30000
        o = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(S)
        d = rng.randn(n, 3).astype(np.float32)
        d[:100, 0] = 0.0
        d[100:200, 1] = -0.0
        d[200:300, 2] = 0.0
        d[300:350, :2] = -0.0
        out = {}
        try:
            for grid in (0, 1):
                host_harness.hh_set_grid(grid)
                t = np.zeros(n, np.float32)
                p = np.zeros(n, np.int32)
                nv, st = C.c_uint64(), C.c_uint64()
                host_harness.hh_closest_hit(spheres.ctypes.data_as(C.c_void_p), len(spheres), 2, C.c_float(0.01), o.ctypes.data_as(C.c_void_p),
                                            d.ctypes.data_as(C.c_void_p), n, t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p),
                                            C.byref(nv), C.byref(st))
                out[grid] = (t, p)
        finally:
            host_harness.hh_set_grid(0)
        assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
        t0, p0 = oracle_mod.Oracle(spheres).closest_hit(o, d, use_bvh=False)
        assert np.array_equal(t0, out[1][0]) and np.array_equal(p0, out[1][1])
    W, H, spp, sub, depth = 48, 27, 3, 5, 50
    cam = oracle_mod.rtiow_camera(W, H)
    hp = _hh_params(cam, W, H, spp, sub, depth)
    orc = oracle_mod.Oracle(rtiow)
    want, stats = orc.render_mean(orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BRUTE))
    steps = {}
    try:
        for grid in (0, 1):
            host_harness.hh_set_grid(grid)
            host_harness.hh_set_wide(1)
            mean = np.zeros((H, W, 4), np.float32)
            segs, nv, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
            host_harness.hh_render_mean(rtiow.ctypes.data_as(C.c_void_p), len(rtiow), 1, C.c_float(0.01), C.byref(hp),
                                        mean.ctypes.data_as(C.c_void_p), C.byref(segs), C.byref(nv), C.byref(st))
            assert segs.value == stats.segments and np.array_equal(mean, want)
            steps[grid] = nv.value / segs.value
    finally:
        host_harness.hh_set_grid(0)
        host_harness.hh_set_wide(0)
    assert steps[1] < 0.5 * steps[0]


def test_hit_gate_makes_the_closest_hit_independent_of_the_bvh(host_harness, oracle_mod):
    """The hit-point gate.  Rays that start 250-750 units from spheres of radius 0.1-0.3 -- configs[4]'s geometry, where
    the float quadratic of RayTracer.cu:239-253 is mostly rounding noise: WITHOUT the gate brute force, the oracle's BVH and the product's
    LBVH (run here on the CPU, exact math) disagree on which phantom hits they report; WITH it all three agree bit for bit, for pair
    nodes, octant-mirrored pairs and the canonical wide nodes, with the default pad and with 10 % boxes."""
    spheres = np.ascontiguousarray(oracle_mod.random_scene(3000, 0x5EED0002, 4.0, 1))
    rng = np.random.RandomState(5)
    n = 60000
    o = np.zeros((n, 3), np.float32)
    o[:, 2] = np.float32(500.0)
    o[:, :2] = (rng.rand(n, 2).astype(np.float32) - np.float32(0.5)) * np.float32(2.0)
    tgt = (rng.rand(n, 3).astype(np.float32) - np.float32(0.5)) * np.float32(8.0)
    d = (tgt - o).astype(np.float32)
    orc = oracle_mod.Oracle(spheres)
    tb, pb = orc.closest_hit(o, d, use_bvh=False, gate=True)
    tv, pv = orc.closest_hit(o, d, use_bvh=True, gate=True)
    assert np.array_equal(tb, tv) and np.array_equal(pb, pv)
    tu, pu = orc.closest_hit(o, d, use_bvh=False, gate=False)
    phantom = int(((pu != pb) | (tu != tb)).sum())
    print("far rays: %d of %d hits change when the gate is applied" % (phantom, int((pu >= 0).sum())))
    assert phantom > 50 and (pb >= 0).sum() > 1000          # the regime really is noise-dominated, and there still are hits
    sp, op, dp = spheres.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p)
    host_harness.hh_set_gate(1)
    try:
        for oct_, wide, pad in ((0, 0, 0.01), (1, 0, 0.01), (0, 2, 0.01), (0, 0, 0.10)):
            host_harness.hh_set_oct(oct_)
            host_harness.hh_set_wide(wide)
            t1, p1 = np.zeros(n, np.float32), np.zeros(n, np.int32)
            host_harness.hh_closest_hit(sp, len(spheres), 2, C.c_float(pad), op, dp, n, t1.ctypes.data_as(C.c_void_p), p1.ctypes.data_as(C.c_void_p), None, None)
            assert np.array_equal(tb, t1) and np.array_equal(pb, p1), (oct_, wide, pad, int((pb != p1).sum()))
        host_harness.hh_set_oct(0)
        host_harness.hh_set_wide(0)
        host_harness.hh_set_gate(0)
        t2, p2 = np.zeros(n, np.float32), np.zeros(n, np.int32)
        host_harness.hh_closest_hit(sp, len(spheres), 2, C.c_float(0.01), op, dp, n, t2.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p), None, None)
        assert not (np.array_equal(tu, t2) and np.array_equal(pu, p2))     # ... which is what the gate is for
    finally:
        host_harness.hh_set_gate(0)
        host_harness.hh_set_oct(0)
        host_harness.hh_set_wide(0)


def test_hit_gate_never_fires_on_the_rtiow_scene(oracle_mod, rtiow):
    """The headline kernel (shared-memory wide nodes) does not spend instructions on the gate: on scenes of that scale it is a no-op.
    Whole C1 frame (400x225, 10 spp, depth 50 = 2.6 M segments): gated and ungated oracle are bit-identical, and so are the golden
    frames produced by the reference's own RayTracer.cu (tests/test_oracle.py runs ungated)."""
    W, H = 400, 225
    cam = oracle_mod.rtiow_camera(W, H)
    orc = oracle_mod.Oracle(rtiow)
    a, sa = orc.render_mean(orc.params(cam, W, H, 10, 1, 50, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BVH))
    b, sb = orc.render_mean(orc.params(cam, W, H, 10, 1, 50, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BVH | oracle_mod.CLOSEST_GATE))
    assert sa.segments == sb.segments and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("wide,leaf", [(0, 2), (1, 1), (0, 1)])
def test_negative_radius_hollow_glass_on_cpu(host_harness, oracle_mod, rtiow, wide, leaf):
    """a NEGATIVE radius (sphere.h:17-28; RTIOW's hollow glass sphere).  The reference's own programs pin the oracle on
    this scene (tests/golden/ref_render_hollow_*.npz); here the product's math and LBVH (|r| boxes, signed r in the normal
    (p - c) / r of RayTracer.cu:257), run on the CPU, reproduce the oracle bit for bit -- pair nodes and the 4-wide nodes with the
    huge list, two coincident centres included."""
    W, H, spp, sub, depth = 48, 27, 4, 2, 50
    spheres = oracle_mod.hollow_glass_scene(rtiow)
    cam = oracle_mod.camera((3.0, 1.6, 4.0), (-3.0, -0.6, -4.0), 25.0, W / H, 0.02, 5.0)
    hp = _hh_params(cam, W, H, spp, sub, depth)
    mean = np.zeros((H, W, 4), np.float32)
    segs, nv, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
    host_harness.hh_set_wide(wide)
    try:
        host_harness.hh_render_mean(spheres.ctypes.data_as(C.c_void_p), len(spheres), leaf, C.c_float(0.01), C.byref(hp),
                                    mean.ctypes.data_as(C.c_void_p), C.byref(segs), C.byref(nv), C.byref(st))
    finally:
        host_harness.hh_set_wide(0)
    orc = oracle_mod.Oracle(spheres)
    want, stats = orc.render_mean(orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BRUTE))
    assert segs.value == stats.segments
    assert np.array_equal(mean.view(np.uint32), want.view(np.uint32))
    # the inner surface really takes part: without the extra sphere the picture is another one
    plain = oracle_mod.Oracle(rtiow)
    other, _ = plain.render_mean(orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_FORWARD, closest=oracle_mod.CLOSEST_BRUTE))
    assert (other.view(np.uint32) != want.view(np.uint32)).any(axis=-1).mean() > 0.05
