"""ctypes binding of libvenusaur_b200.so -- one Python declaration per entry point of include/venusaur_b200.h.

There is no fallback: if the shared library is missing it is built with nvcc (venusaur_b200.build); if that fails,
or if no CUDA device is present when a handle is created, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_float3 = C.c_float * 3


class vn_sphere(C.Structure):
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("cz", C.c_float), ("r", C.c_float),
                ("ax", C.c_float), ("ay", C.c_float), ("az", C.c_float),
                ("fuzz_or_ir", C.c_float), ("type", C.c_uint32)]


class vn_params(C.Structure):
    _fields_ = [("image", C.c_void_p),
                ("width", C.c_uint32), ("height", C.c_uint32),
                ("samples_per_pixel", C.c_uint32), ("subframe_index", C.c_uint32),
                ("max_depth", C.c_uint32), ("accum_count", C.c_uint32),
                ("origin", c_float3), ("u", c_float3), ("v", c_float3), ("w", c_float3),
                ("lens_radius", C.c_float),
                ("row_begin", C.c_uint32), ("row_end", C.c_uint32),
                ("flags", C.c_uint32)]


class vn_stats(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("paths", C.c_uint64), ("node_visits", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("segments_total", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("kernel_launches_total", C.c_uint32),
                ("ms_render", C.c_float), ("ms_trace", C.c_float), ("ms_build", C.c_float), ("ms_upload", C.c_float)]


class vn_bvh_info(C.Structure):
    _fields_ = [("num_spheres", C.c_uint64), ("num_nodes", C.c_uint64), ("max_leaf_size", C.c_uint32),
                ("scene_in_smem", C.c_uint32), ("bounds_lo", c_float3), ("bounds_hi", c_float3)]


class vn_node32(C.Structure):
    _fields_ = [("lo", c_float3), ("link", C.c_uint32), ("hi", c_float3), ("aux", C.c_uint32)]


VN_OK = 0
VN_LAMBERTIAN, VN_METAL, VN_DIELECTRIC = 0, 1, 2
VN_EXACT, VN_IMAGE_HOST, VN_ACCUM_SUM, VN_NO_TONEMAP = 1 << 0, 1 << 1, 1 << 2, 1 << 3
VN_WAVEFRONT, VN_COUNTERS, VN_ASYNC, VN_FAST = 1 << 4, 1 << 5, 1 << 6, 1 << 7
VN_GRID = 1 << 11

# name -> (restype, argtypes); must list every VN_API symbol of include/venusaur_b200.h (checked by tests)
_P = C.POINTER
SIGNATURES = {
    "vn_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "vn_destroy": (None, [C.c_void_p]),
    "vn_last_error": (C.c_char_p, [C.c_void_p]),
    "vn_device_count": (C.c_int, []),
    "vn_version": (C.c_char_p, []),
    "vn_set_spheres": (C.c_int, [C.c_void_p, _P(vn_sphere), C.c_uint64]),
    "vn_update_spheres": (C.c_int, [C.c_void_p, _P(vn_sphere), C.c_uint64]),
    "vn_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "vn_build_bvh": (C.c_int, [C.c_void_p]),
    "vn_get_bvh_info": (C.c_int, [C.c_void_p, _P(vn_bvh_info)]),
    "vn_read_bvh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "vn_read_sched_counters": (C.c_int, [C.c_void_p, _P(C.c_uint64)]),
    "vn_read_timeline": (C.c_int, [C.c_void_p, _P(C.c_uint32)]),
    "vn_read_timeline_ex": (C.c_int, [C.c_void_p, _P(C.c_uint32)]),
    "vn_read_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "vn_last_accel": (C.c_int, [C.c_void_p]),
    "vn_read_huge": (C.c_int, [C.c_void_p, _P(C.c_uint32)]),
    "vn_read_wide_bvh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, _P(C.c_uint32), _P(C.c_uint32)]),
    "vn_resize": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "vn_reset_accum": (C.c_int, [C.c_void_p]),
    "vn_render": (C.c_int, [C.c_void_p, _P(vn_params)]),
    "vn_render_subframes": (C.c_int, [C.c_void_p, _P(vn_params), C.c_uint32]),
    "vn_render_subframes_strided": (C.c_int, [C.c_void_p, _P(vn_params), C.c_uint32, C.c_uint32]),
    "vn_tonemap": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_uint32]),
    "vn_synchronize": (C.c_int, [C.c_void_p]),
    "vn_get_stats": (C.c_int, [C.c_void_p, _P(vn_stats)]),
    "vn_reset_stats": (C.c_int, [C.c_void_p]),
    "vn_read_accum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vn_write_accum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vn_accum_device_ptr": (C.c_int, [C.c_void_p, _P(C.c_void_p)]),
    "vn_set_accum_external": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vn_reduce_tonemap_peers": (C.c_int, [C.c_void_p, _P(C.c_void_p), C.c_uint32, C.c_float, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_uint32]),
    "vn_reduce_tonemap_peers_to": (C.c_int, [C.c_void_p, _P(C.c_void_p), C.c_uint32, C.c_float, C.c_uint32, C.c_uint32,
                                             C.c_void_p, C.c_void_p, C.c_uint32]),
    "vn_sync_flags": (C.c_int, [C.c_void_p, _P(C.c_void_p)]),
    "vn_signal": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "vn_wait_flags": (C.c_int, [C.c_void_p, _P(C.c_void_p), C.c_uint32, C.c_uint32]),
    "vn_check_flags": (C.c_int, [C.c_void_p]),
    "vn_reduce_tonemap_peers_wait": (C.c_int, [C.c_void_p, _P(C.c_void_p), C.c_uint32, C.c_float, C.c_uint32, C.c_uint32,
                                               C.c_void_p, C.c_void_p, _P(C.c_void_p), C.c_uint32, C.c_uint32]),
    "vn_multi_create": (C.c_int, [_P(C.c_int), C.c_int, _P(C.c_void_p)]),
    "vn_multi_destroy": (None, [C.c_void_p]),
    "vn_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "vn_multi_device_count": (C.c_int, [C.c_void_p]),
    "vn_multi_device": (C.c_void_p, [C.c_void_p, C.c_int]),
    "vn_multi_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "vn_multi_set_spheres": (C.c_int, [C.c_void_p, _P(vn_sphere), C.c_uint64]),
    "vn_multi_build_bvh": (C.c_int, [C.c_void_p]),
    "vn_multi_render": (C.c_int, [C.c_void_p, _P(vn_params), C.c_uint32]),
    "vn_multi_synchronize": (C.c_int, [C.c_void_p]),
    "vn_multi_read_accum": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vn_multi_get_stats": (C.c_int, [C.c_void_p, _P(vn_stats)]),
    "vn_multi_subframes_accumulated": (C.c_uint32, [C.c_void_p]),
    "vn_buffer_alloc": (C.c_int, [C.c_int, C.c_uint64, C.c_int, _P(C.c_void_p), _P(C.c_void_p)]),
    "vn_buffer_free": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "vn_buffer_copy_to_host": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]),
    "vn_stream_synchronize": (C.c_int, [C.c_int, C.c_void_p]),
    "vn_stream": (C.c_void_p, [C.c_void_p]),
    "vn_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vn_ipc_export_at": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _P(C.c_uint64)]),
    "vn_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_void_p)]),
    "vn_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vn_scene_rtiow_final": (C.c_uint32, [_P(vn_sphere), C.c_uint32]),
    "vn_scene_random": (None, [_P(vn_sphere), C.c_uint64, C.c_uint32, C.c_float, C.c_uint32]),
    "vn_camera_frame": (None, [c_float3, c_float3, C.c_float, C.c_float, C.c_float, C.c_float,
                               c_float3, c_float3, c_float3, c_float3, _P(C.c_float)]),
    "vn_test_rng": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vn_trace_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]),
    "vn_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]),
    "vn_morton_codes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "vn_test_make_color": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]),
    "vn_test_scatter": (C.c_int, [C.c_void_p, C.c_uint32, C.c_float * 4, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Loads (building first if needed) libvenusaur_b200.so and declares every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    # rebuild only when the sources' CONTENT differs from what the library was built from (build.py keeps a hash next to the .so);
    # build() itself is serialised by a file lock, so one process per GPU may all land here at once
    if not os.path.exists(_build.LIB) or (os.path.isdir(_build.CSRC) and not _build.up_to_date() and _build.shutil.which("nvcc")):
        _build.build()
    path = _build.LIB
    if os.environ.get("VN_EXPERIMENT"):        # a compile-time variant built by `python -m venusaur_b200.build --exp N` (A/B measurements)
        path = os.path.join(_build.HERE, "libvenusaur_b200_exp%d.so" % int(os.environ["VN_EXPERIMENT"]))
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is not exported: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
