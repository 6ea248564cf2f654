"""Generates the committed golden vectors under tests/golden/ from the REFERENCE's own sources.

Run in the build container (needs /root/reference):   python tests/golden/gen_golden.py

  ref_kat.json         known answers of the reference's random.cuh (tea<N>, lcg, rnd) and vec_math.h
                       (normalize, reflect, refract, lerp) + RayTracer.cu's make_color / reflectance / near_zero,
                       produced by oracle/_ref/libvenusaur_ref.so = those headers compiled for the host.
  ref_render_*.npz     accum (float4) + image (uchar4) + segment count written by the reference's unmodified
                       __raygen__rg / __intersection__hit_sphere / __closesthit__* / __miss__ms run on the CPU through
                       oracle/ref_shim/optix.h (brute-force optixTrace).  g++ evaluates the random_float() calls inside
                       make_float3(...) right-to-left, so these pin the oracle in its DRAW_ZYX mode; every other
                       aspect (draw schedule, depth semantics, operation order, unwind multiplication) is shared.
The GPU box has no /root/reference: tests only read the files written here.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def f32(x):
    return float(np.float32(x))


def main():
    ol.build()
    r = ol.load_ref()
    assert r is not None, "needs /root/reference to build oracle/_ref"
    kat = {"source": "oracle/_ref/libvenusaur_ref.so = /root/reference/Core/{random.cuh,vec_math.h,RayTracer.cu} compiled by g++ -ffp-contract=off"}
    kat["tea4"] = [[a, b, int(r.ref_tea4(a, b))] for a, b in
                   [(0, 0), (0, 1), (1, 1), (399, 1), (400, 1), (89999, 1), (2073599, 1), (2073599, 64), (8294399, 256), (12345, 7),
                    (0xFFFFFFFF, 0xFFFFFFFF), (123456789, 987654321)]]
    kat["tea1"] = [[1, 2, int(r.ref_tea1(1, 2))]]
    kat["tea16"] = [[1, 2, int(r.ref_tea16(1, 2))]]
    s = C.c_uint32(0)
    kat["lcg_from_0"] = [int(r.ref_lcg(C.byref(s))) for _ in range(3)] + [int(s.value)]
    s = C.c_uint32(r.ref_tea4(0, 1))
    kat["rnd_from_tea4_0_1"] = [f32(r.ref_rnd(C.byref(s))) for _ in range(8)] + [int(s.value)]
    s = C.c_uint32(0)
    mx = 0.0
    st = C.c_uint32(4294967295)
    # state that yields lcg == 0xFFFFFF: just scan a few thousand draws for the max
    for _ in range(200000):
        mx = max(mx, f32(r.ref_rnd(C.byref(st))))
    kat["rnd_max_seen"] = mx

    def vec(fn, *args):
        out = np.zeros(3, np.float32)
        arrs = [np.array(a, np.float32) if isinstance(a, (list, tuple)) else a for a in args]
        cargs = [a.ctypes.data_as(C.c_void_p) if isinstance(a, np.ndarray) else C.c_float(a) for a in arrs]
        fn(*cargs, out.ctypes.data_as(C.c_void_p))
        return [f32(x) for x in out]

    v = [0.3, -0.8, 0.52]
    n = vec(r.ref_normalize, v)
    kat["normalize"] = [v, n]
    kat["reflect"] = [n, [0, 1, 0], vec(r.ref_reflect, n, [0, 1, 0])]
    kat["refract"] = [n, [0, 1, 0], f32(1 / 1.5), vec(r.ref_refract, n, [0, 1, 0], f32(1 / 1.5))]
    kat["lerp"] = [[1, 1, 1], [0.5, 0.7, 1.0], 0.25, vec(r.ref_lerp, [1, 1, 1], [0.5, 0.7, 1.0], 0.25)]
    rng = np.random.RandomState(1234)
    cols = np.concatenate([rng.rand(64, 3).astype(np.float32) * 1.2 - 0.1,
                           np.array([[0, 0, 0], [1, 1, 1], [0.0031308, 0.0031307, 0.0031309], [0.5, 0.25, 0.75], [2, -1, 0.999999]], np.float32)])
    mc = []
    for c in cols:
        out = np.zeros(4, np.uint8)
        r.ref_make_color(c.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        mc.append([[f32(x) for x in c], [int(x) for x in out]])
    kat["make_color"] = mc
    kat["reflectance"] = [[f32(c), f32(i), f32(r.ref_reflectance(f32(c), f32(i)))] for c, i in
                          [(0.0, 1.5), (0.3, 1.5), (1.0, 1.5), (0.7, 1 / 1.5), (0.05, 0.6666667), (0.999, 2.4)]]
    kat["near_zero"] = [[list(map(f32, vv)), int(r.ref_near_zero(np.array(vv, np.float32).ctypes.data_as(C.c_void_p)))] for vv in
                        [(0, 0, 0), (1e-9, -1e-9, 1e-9), (9.99999993922529e-09, 0, 0), (1.0000001e-8, 0, 0), (1e-7, 0, 0)]]
    kat["sizeof_params"] = int(r.ref_sizeof_params())
    kat["sizeof_sphere_record"] = int(r.ref_sizeof_sphere_record())
    with open(os.path.join(HERE, "ref_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- whole-path goldens from the reference's programs
    spheres = ol.rtiow_final_scene()
    W, H, SPP = 48, 27, 4
    cam = ol.rtiow_camera(W, H)
    for sub, depth in [(0, 4), (3, 4), (1, 50)]:
        P = ol.ref_params()
        P.width, P.height, P.samples_per_pixel, P.subframe_index = W, H, SPP, sub
        P.origin, P.u, P.v, P.w, P.lens_radius = ol.c_float3(*cam[0]), ol.c_float3(*cam[1]), ol.c_float3(*cam[2]), ol.c_float3(*cam[3]), float(cam[4])
        prev = (np.random.RandomState(7).rand(H, W, 4).astype(np.float32))
        prev[..., 3] = 1.0
        acc = prev.copy()
        img = np.zeros((H, W, 4), np.uint8)
        seg = r.ref_render(spheres.ctypes.data_as(C.c_void_p), len(spheres), C.byref(P), None, 0, acc.ctypes.data_as(C.c_void_p),
                           img.ctypes.data_as(C.c_void_p), 0 if depth == 4 else depth, 8)
        np.savez_compressed(os.path.join(HERE, "ref_render_%dx%d_spp%d_sub%d_depth%d.npz" % (W, H, SPP, sub, depth)),
                            prev=prev, accum=acc, image=img, segments=np.uint64(seg), width=W, height=H, spp=SPP, subframe=sub,
                            max_depth=depth, origin=cam[0], u=cam[1], v=cam[2], w=cam[3], lens=cam[4])
        print("golden sub=%d depth=%d segments=%d" % (sub, depth, seg))
    # ---- negative radius (sphere.h:17-28): RTIOW's hollow-glass trick -- a second dielectric sphere of radius -0.9 inside
    # the big glass sphere, whose normal (p - c) / r (RayTracer.cu:257) points inwards.  Seen from close by, depth 50.
    hollow = ol.hollow_glass_scene(spheres)
    cam = ol.camera((3.0, 1.6, 4.0), (-3.0, -0.6, -4.0), 25.0, W / H, 0.02, 5.0)
    for sub, depth in [(2, 50)]:
        P = ol.ref_params()
        P.width, P.height, P.samples_per_pixel, P.subframe_index = W, H, SPP, sub
        P.origin, P.u, P.v, P.w, P.lens_radius = ol.c_float3(*cam[0]), ol.c_float3(*cam[1]), ol.c_float3(*cam[2]), ol.c_float3(*cam[3]), float(cam[4])
        prev = (np.random.RandomState(8).rand(H, W, 4).astype(np.float32))
        prev[..., 3] = 1.0
        acc = prev.copy()
        img = np.zeros((H, W, 4), np.uint8)
        seg = r.ref_render(hollow.ctypes.data_as(C.c_void_p), len(hollow), C.byref(P), None, 0, acc.ctypes.data_as(C.c_void_p),
                           img.ctypes.data_as(C.c_void_p), depth, 8)
        np.savez_compressed(os.path.join(HERE, "ref_render_hollow_%dx%d_spp%d_sub%d_depth%d.npz" % (W, H, SPP, sub, depth)),
                            prev=prev, accum=acc, image=img, segments=np.uint64(seg), width=W, height=H, spp=SPP, subframe=sub,
                            max_depth=depth, origin=cam[0], u=cam[1], v=cam[2], w=cam[3], lens=cam[4])
        print("golden hollow sub=%d depth=%d segments=%d" % (sub, depth, seg))
    np.save(os.path.join(HERE, "rtiow_final_scene.npy"), spheres)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
