"""CPU tests: the oracle against the golden vectors generated from the reference's own sources."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLD, "ref_kat.json")) as f:
        return json.load(f)


def f32(x):
    return float(np.float32(x))


def test_rng_kats(oracle_mod, kat):
    o = oracle_mod.load()
    for a, b, want in kat["tea4"]:
        assert o.orc_tea(4, a, b) == want
    for a, b, want in kat["tea1"]:
        assert o.orc_tea(1, a, b) == want
    for a, b, want in kat["tea16"]:
        assert o.orc_tea(16, a, b) == want
    s = C.c_uint32(0)
    assert [o.orc_lcg(C.byref(s)) for _ in range(3)] + [s.value] == kat["lcg_from_0"]
    s = C.c_uint32(o.orc_tea(4, 0, 1))
    got = [f32(o.orc_rnd(C.byref(s))) for _ in range(8)] + [s.value]
    assert got == kat["rnd_from_tea4_0_1"]
    # SURVEY section 4 table, independently transcribed
    assert o.orc_tea(4, 0, 0) == 1576399551 and o.orc_tea(4, 2073599, 64) == 3513108779 and o.orc_tea(4, 12345, 7) == 1964180806


def _vec(fn, *args):
    out = np.zeros(3, np.float32)
    cargs = []
    keep = []
    for a in args:
        if isinstance(a, (list, tuple)):
            arr = np.array(a, np.float32)
            keep.append(arr)
            cargs.append(arr.ctypes.data_as(C.c_void_p))
        else:
            cargs.append(C.c_float(a))
    fn(*cargs, out.ctypes.data_as(C.c_void_p))
    return [f32(x) for x in out]


def test_float3_kats(oracle_mod, kat):
    o = oracle_mod.load()
    v, n = kat["normalize"]
    assert _vec(o.orc_normalize, v) == n
    i, nn, want = kat["reflect"]
    assert _vec(o.orc_reflect, i, nn) == want
    i, nn, eta, want = kat["refract"]
    assert _vec(o.orc_refract, i, nn, eta) == want
    a, b, t, want = kat["lerp"]
    assert _vec(o.orc_lerp, a, b, t) == want
    for c, i, want in kat["reflectance"]:
        assert f32(o.orc_reflectance(c, i)) == want


def test_make_color_kats(oracle_mod, kat):
    cols = np.array([c for c, _ in kat["make_color"]], np.float32)
    want = np.array([w for _, w in kat["make_color"]], np.uint8)
    assert np.array_equal(oracle_mod.make_color(cols), want)


def test_scene_matches_probe(rtiow):
    # libstdc++ + default mt19937, left-to-right: 486 spheres, 392 L / 71 M / 23 D
    assert len(rtiow) == 486
    assert [int((rtiow["type"] == t).sum()) for t in (0, 1, 2)] == [392, 71, 23]
    assert np.array_equal(rtiow.view(np.uint8), np.load(os.path.join(GOLD, "rtiow_final_scene.npy")).view(np.uint8))
    assert tuple(rtiow[0][["cx", "cy", "cz", "r"]]) == (0.0, -1000.0, 0.0, 1000.0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ref_render_*.npz"))))
def test_oracle_equals_reference_programs(oracle_mod, rtiow, path):
    """Whole-path pin: the oracle (reference multiplication order, GCC's draw order) reproduces, bit for bit, what the
    reference's unmodified OptiX programs computed on the CPU (tests/golden/gen_golden.py)."""
    g = np.load(path)
    W, H, spp, sub, depth = int(g["width"]), int(g["height"]), int(g["spp"]), int(g["subframe"]), int(g["max_depth"])
    cam = (g["origin"], g["u"], g["v"], g["w"], g["lens"])
    # ref_render_hollow_*: the scene with a negative-radius sphere (hollow glass), see gen_golden.py
    orc = oracle_mod.Oracle(oracle_mod.hollow_glass_scene(rtiow) if "hollow" in os.path.basename(path) else rtiow)
    for closest in (oracle_mod.CLOSEST_BRUTE, oracle_mod.CLOSEST_BVH, oracle_mod.CLOSEST_BVH | oracle_mod.CLOSEST_GATE):
        p = orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_UNWIND, draw=oracle_mod.DRAW_ZYX, closest=closest)
        mean, st = orc.render_mean(p)
        assert st.segments == int(g["segments"])
        # RayTracer.cu:208-216: blend with the previous accum when subframe_index > 0
        acc, img = oracle_mod.accumulate_tonemap(g["prev"], mean, sub > 0, np.float32(1.0) / np.float32(sub + 1))
        assert np.array_equal(acc, g["accum"])
        assert np.array_equal(img, g["image"])


def test_oracle_against_live_reference(oracle_mod, rtiow):
    """Same pin on a fresh configuration, when oracle/_ref is available (built here or shipped prebuilt)."""
    r = oracle_mod.load_ref()
    if r is None:
        pytest.skip("oracle/_ref not available")
    W, H, spp, sub = 40, 30, 3, 5
    cam = oracle_mod.rtiow_camera(W, H)
    P = oracle_mod.ref_params()
    P.width, P.height, P.samples_per_pixel, P.subframe_index = W, H, spp, sub
    P.origin, P.u, P.v, P.w, P.lens_radius = (oracle_mod.c_float3(*cam[0]), oracle_mod.c_float3(*cam[1]), oracle_mod.c_float3(*cam[2]),
                                              oracle_mod.c_float3(*cam[3]), float(cam[4]))
    for depth in (4, 12):
        acc = np.zeros((H, W, 4), np.float32)
        img = np.zeros((H, W, 4), np.uint8)
        seg = r.ref_render(rtiow.ctypes.data_as(C.c_void_p), len(rtiow), C.byref(P), None, 0, acc.ctypes.data_as(C.c_void_p),
                           img.ctypes.data_as(C.c_void_p), 0 if depth == 4 else depth, 4)
        orc = oracle_mod.Oracle(rtiow)
        p = orc.params(cam, W, H, spp, sub, depth, atten=oracle_mod.ATTEN_UNWIND, draw=oracle_mod.DRAW_ZYX, closest=oracle_mod.CLOSEST_BVH)
        mean, st = orc.render_mean(p)
        want, want_img = oracle_mod.accumulate_tonemap(np.zeros_like(mean), mean, True, np.float32(1.0) / np.float32(sub + 1))
        assert st.segments == seg
        assert np.array_equal(want, acc) and np.array_equal(want_img, img)


def test_forward_vs_unwind_order(oracle_mod, rtiow):
    """The kernels multiply albedos forward, the reference on recursion unwind: same factors, <= depth * 2^-24."""
    W, H = 64, 36
    orc = oracle_mod.Oracle(rtiow)
    cam = oracle_mod.rtiow_camera(W, H)
    a, sa = orc.render_mean(orc.params(cam, W, H, 4, 1, 50, atten=oracle_mod.ATTEN_UNWIND))
    b, sb = orc.render_mean(orc.params(cam, W, H, 4, 1, 50, atten=oracle_mod.ATTEN_FORWARD))
    assert sa.segments == sb.segments
    rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-6)
    assert rel.max() < 1e-5


def test_oracle_bvh_equals_brute_force(oracle_mod, rtiow):
    W, H = 64, 36
    orc = oracle_mod.Oracle(rtiow)
    cam = oracle_mod.rtiow_camera(W, H)
    a, sa = orc.render_mean(orc.params(cam, W, H, 2, 2, 50, closest=oracle_mod.CLOSEST_BRUTE))
    b, sb = orc.render_mean(orc.params(cam, W, H, 2, 2, 50, closest=oracle_mod.CLOSEST_BVH))
    assert sa.segments == sb.segments and np.array_equal(a, b)
    rng = np.random.RandomState(3)
    o = (rng.rand(20000, 3).astype(np.float32) - 0.5) * np.float32(30.0)
    o[:, 1] = np.abs(o[:, 1]) * np.float32(0.2) + np.float32(0.01)
    d = rng.randn(20000, 3).astype(np.float32)
    t0, p0 = orc.closest_hit(o, d, use_bvh=False)
    t1, p1 = orc.closest_hit(o, d, use_bvh=True)
    assert np.array_equal(t0, t1) and np.array_equal(p0, p1)


def test_depth_semantics(oracle_mod, rtiow):
    """max_depth = N means at most N segments per path (RayTracer.cu:172,184)."""
    W, H = 32, 18
    orc = oracle_mod.Oracle(rtiow)
    cam = oracle_mod.rtiow_camera(W, H)
    for depth in (1, 2, 4, 7):
        _, st = orc.render_mean(orc.params(cam, W, H, 4, 1, depth))
        assert st.max_segments_in_path <= depth and st.paths == W * H * 4
    m1, _ = orc.render_mean(orc.params(cam, W, H, 4, 1, 1))
    # depth 1: only sky pixels are non-black
    assert (m1[..., :3].sum(axis=-1) == 0).any() and (m1[..., :3].sum(axis=-1) > 0).any()
