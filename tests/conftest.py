import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle_lib
    oracle_lib.load()
    return oracle_lib


@pytest.fixture(scope="session")
def rtiow(oracle_mod):
    return oracle_mod.rtiow_final_scene()


@pytest.fixture(scope="session")
def host_harness():
    """The product's exact device math + LBVH bodies compiled for the CPU (tests/host_harness.cpp). Test-only."""
    import ctypes as C
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libhost_harness.so")
    srcs = [os.path.join(ROOT, "tests", "host_harness.cpp"), os.path.join(ROOT, "venusaur_b200", "csrc", "vn_math.cuh"),
            os.path.join(ROOT, "venusaur_b200", "csrc", "lbvh_core.cuh"), os.path.join(ROOT, "venusaur_b200", "csrc", "grid_core.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "venusaur_b200", "csrc"),
                        "-o", lib, srcs[0]], check=True)
    h = C.CDLL(lib)
    h.hh_tea4.restype = C.c_uint32
    h.hh_tea4.argtypes = [C.c_uint32, C.c_uint32]
    h.hh_build_bvh.restype = C.c_uint64
    h.hh_build_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    h.hh_closest_hit.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    h.hh_render_mean.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    h.hh_set_oct.argtypes = [C.c_int]
    h.hh_set_gate.argtypes = [C.c_int]
    h.hh_set_wide.argtypes = [C.c_int]
    h.hh_set_grid.argtypes = [C.c_int]
    h.hh_huge_list.restype = C.c_uint32
    h.hh_huge_list.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p]
    h.hh_build_grid.restype = C.c_int
    h.hh_build_grid.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    h.hh_set_sah_max.argtypes = [C.c_uint32]
    h.hh_build_wide.restype = C.c_uint64
    h.hh_build_wide.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_uint64, C.c_void_p]
    h.hh_scatter.argtypes = [C.c_uint32, C.c_float * 4, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    h.hh_morton30.restype = C.c_uint32
    h.hh_morton30.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    return h


@pytest.fixture(scope="session")
def ctx():
    """A library handle on cuda:0; GPU tests only."""
    import venusaur_b200 as vb
    c = vb.Context(0)
    yield c
    c.close()
