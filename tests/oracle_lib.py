"""ctypes wrapper of the CPU oracle (oracle/liboracle.so) -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libvenusaur_ref.so")

SPHERE_DTYPE = np.dtype([("cx", "f4"), ("cy", "f4"), ("cz", "f4"), ("r", "f4"), ("ax", "f4"), ("ay", "f4"),
                         ("az", "f4"), ("fuzz_or_ir", "f4"), ("type", "u4")])
c_float3 = C.c_float * 3


class orc_params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples_per_pixel", C.c_uint32),
                ("subframe_index", C.c_uint32), ("max_depth", C.c_uint32),
                ("origin", c_float3), ("u", c_float3), ("v", c_float3), ("w", c_float3), ("lens_radius", C.c_float),
                ("atten_order", C.c_uint32), ("draw_order", C.c_uint32), ("closest", C.c_uint32), ("threads", C.c_uint32)]


class orc_stats(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("paths", C.c_uint64), ("node_visits", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("max_segments_in_path", C.c_uint64)]


class ref_params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples_per_pixel", C.c_uint32), ("subframe_index", C.c_uint32),
                ("origin", c_float3), ("u", c_float3), ("v", c_float3), ("w", c_float3), ("lens_radius", C.c_float)]


ATTEN_UNWIND, ATTEN_FORWARD = 0, 1
DRAW_XYZ, DRAW_ZYX = 0, 1
CLOSEST_BRUTE, CLOSEST_BVH = 0, 1
CLOSEST_GATE = 2        # flag: hit-point gate (oracle.cpp::hit_gate), what the product applies to scenes traversed from L2 / HBM

_lib = None
_ref = None


def build():
    """Compiles the oracle (and oracle/_ref when /root/reference is present) with the committed Makefile."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


def _stale(lib, srcs):
    return not os.path.exists(lib) or any(os.path.getmtime(os.path.join(ORACLE_DIR, s)) > os.path.getmtime(lib) for s in srcs)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if _stale(LIB, ["oracle.cpp", "oracle.h"]):
            build()
        o = C.CDLL(LIB)
        o.orc_tea.restype = C.c_uint32
        o.orc_tea.argtypes = [C.c_uint32] * 3
        o.orc_lcg.restype = C.c_uint32
        o.orc_lcg.argtypes = [C.POINTER(C.c_uint32)]
        o.orc_rnd.restype = C.c_float
        o.orc_rnd.argtypes = [C.POINTER(C.c_uint32)]
        o.orc_reflectance.restype = C.c_float
        o.orc_reflectance.argtypes = [C.c_float, C.c_float]
        o.orc_refract.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        o.orc_lerp.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        o.orc_scene_rtiow_final.restype = C.c_uint32
        o.orc_scene_rtiow_final.argtypes = [C.c_void_p, C.c_uint32]
        o.orc_scene_random.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_float, C.c_uint32]
        o.orc_camera.argtypes = [c_float3, c_float3, C.c_float, C.c_float, C.c_float, C.c_float, c_float3, c_float3, c_float3, c_float3,
                                 C.POINTER(C.c_float)]
        o.orc_scene_create.restype = C.c_void_p
        o.orc_scene_create.argtypes = [C.c_void_p, C.c_uint64]
        o.orc_scene_destroy.argtypes = [C.c_void_p]
        o.orc_render_mean.argtypes = [C.c_void_p, C.POINTER(orc_params), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(orc_stats)]
        o.orc_accumulate_tonemap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64]
        o.orc_closest_hit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib = o
    return _lib


def load_ref():
    """The reference's own RayTracer.cu compiled for the host (None when neither prebuilt nor buildable here)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB) and os.path.exists("/root/reference/Core/RayTracer.cu"):
            build()
        if not os.path.exists(REF_LIB):
            return None
        r = C.CDLL(REF_LIB)
        for name in ("ref_tea1", "ref_tea4", "ref_tea16"):
            getattr(r, name).restype = C.c_uint32
            getattr(r, name).argtypes = [C.c_uint32, C.c_uint32]
        r.ref_lcg.restype = C.c_uint32
        r.ref_lcg.argtypes = [C.POINTER(C.c_uint32)]
        r.ref_rnd.restype = C.c_float
        r.ref_rnd.argtypes = [C.POINTER(C.c_uint32)]
        r.ref_reflectance.restype = C.c_float
        r.ref_reflectance.argtypes = [C.c_float, C.c_float]
        r.ref_refract.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        r.ref_lerp.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        r.ref_sizeof_params.restype = C.c_uint32
        r.ref_sizeof_sphere_record.restype = C.c_uint32
        r.ref_render.restype = C.c_uint64
        r.ref_render.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(ref_params), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
        _ref = r
    return _ref


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def rtiow_final_scene() -> np.ndarray:
    o = load()
    n = o.orc_scene_rtiow_final(None, 0)
    s = np.zeros(n, SPHERE_DTYPE)
    o.orc_scene_rtiow_final(_p(s), n)
    return s


def hollow_glass_scene(rtiow: np.ndarray) -> np.ndarray:
    """The RTIOW final scene + a dielectric sphere of radius -0.9 inside the big glass sphere at (0, 1, 0): the book's hollow-glass
    trick, a NEGATIVE radius (sphere.h:17-28; the normal (p - c) / r of RayTracer.cu:257 then points inwards)."""
    glass = [i for i in range(len(rtiow)) if rtiow["type"][i] == 2 and rtiow["r"][i] == 1.0]
    assert len(glass) == 1
    extra = rtiow[glass[0]:glass[0] + 1].copy()
    extra["r"] = np.float32(-0.9)
    return np.ascontiguousarray(np.concatenate([rtiow, extra]))


def random_scene(n, seed, S, mix) -> np.ndarray:
    s = np.zeros(n, SPHERE_DTYPE)
    load().orc_scene_random(_p(s), n, seed, S, mix)
    return s


def camera(lookfrom, forward, vfov, aspect, aperture, focal):
    o, u, v, w = c_float3(), c_float3(), c_float3(), c_float3()
    lens = C.c_float()
    load().orc_camera(c_float3(*lookfrom), c_float3(*forward), vfov, aspect, aperture, focal, o, u, v, w, C.byref(lens))
    f = lambda a: np.array(list(a), np.float32)  # noqa: E731
    return f(o), f(u), f(v), f(w), np.float32(lens.value)


def rtiow_camera(width, height):
    return camera((13.0, 2.0, 3.0), (-13.0, -2.0, -3.0), 20.0, width / height, 0.1, 10.0)


class Oracle:
    """A scene + the oracle's render entry points."""

    def __init__(self, spheres: np.ndarray):
        self.o = load()
        self.spheres = np.ascontiguousarray(spheres, SPHERE_DTYPE)
        self.h = self.o.orc_scene_create(_p(self.spheres), len(self.spheres))

    def __del__(self):
        try:
            if self.h:
                self.o.orc_scene_destroy(self.h)
                self.h = None
        except BaseException:  # noqa: BLE001
            pass

    def params(self, cam, width, height, spp, subframe, max_depth, atten=ATTEN_FORWARD, draw=DRAW_XYZ, closest=CLOSEST_BVH, threads=0):
        o, u, v, w, lens = cam
        p = orc_params()
        p.width, p.height, p.samples_per_pixel, p.subframe_index, p.max_depth = width, height, spp, subframe, max_depth
        p.origin, p.u, p.v, p.w, p.lens_radius = c_float3(*o), c_float3(*u), c_float3(*v), c_float3(*w), float(lens)
        p.atten_order, p.draw_order, p.closest, p.threads = atten, draw, closest, threads
        return p

    def render_mean(self, p: orc_params, pixels=None, per_sample=False):
        """pixel_color / spp as (H, W, 4) float32, stats[, per-sample radiance (n_pixels, spp, 3)]."""
        mean = np.zeros((p.height, p.width, 4), np.float32)
        st = orc_stats()
        px = None if pixels is None else np.ascontiguousarray(pixels, np.uint32)
        n = 0 if px is None else len(px)
        ps = None
        if per_sample:
            ps = np.zeros((n if px is not None else p.width * p.height, p.samples_per_pixel, 3), np.float32)
        self.o.orc_render_mean(self.h, C.byref(p), _p(px), n, _p(mean), _p(ps), C.byref(st))
        return (mean, st, ps) if per_sample else (mean, st)

    def closest_hit(self, origins, dirs, use_bvh=False, gate=False):
        o = np.ascontiguousarray(origins, np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        t = np.zeros(len(o), np.float32)
        prim = np.zeros(len(o), np.int32)
        self.o.orc_closest_hit(self.h, int(bool(use_bvh)) | (CLOSEST_GATE if gate else 0), _p(o), _p(d), len(o), _p(t), _p(prim))
        return t, prim


def accumulate_tonemap(prev, mean, blend: bool, a: float):
    """RayTracer.cu:208-216 on (H, W, 4) float32 arrays -> (accum, uchar4 image)."""
    mean = np.ascontiguousarray(mean, np.float32)
    prev = np.ascontiguousarray(prev if prev is not None else np.zeros_like(mean), np.float32)
    out = np.zeros_like(mean)
    img = np.zeros(mean.shape[:-1] + (4,), np.uint8)
    load().orc_accumulate_tonemap(_p(prev), _p(mean), int(blend), float(a), _p(out), _p(img), mean.size // 4)
    return out, img


def make_color(rgb):
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
    out = np.zeros((len(rgb), 4), np.uint8)
    o = load()
    for i in range(len(rgb)):
        o.orc_make_color(_p(rgb[i:i + 1]), _p(out[i:i + 1]))
    return out
