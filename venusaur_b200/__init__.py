"""venusaur_b200 -- B200-native replacement for Venusaur's per-pixel Monte-Carlo hot path.

The product is libvenusaur_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/venusaur_b200.h) and
the header-only C++17 drop-in classes in include/venusaur/.  This package is the Python mirror of that host interface
(api.Renderer / Scene / Camera / CUDAOutputBuffer) used by tests and bench.py.  There is no CPU path.
"""
from .api import (Camera, Context, CUDAOutputBuffer, Exception, MultiContext, Renderer, Scene, random_scene, rtiow_camera,  # noqa: F401,A004
                  rtiow_final_scene)
from ._lib import (VN_ACCUM_SUM, VN_ASYNC, VN_COUNTERS, VN_EXACT, VN_FAST, VN_IMAGE_HOST, VN_GRID, VN_NO_TONEMAP, VN_WAVEFRONT, lib_path,  # noqa: F401
                   load)

__all__ = ["Camera", "Context", "MultiContext", "CUDAOutputBuffer", "Exception", "Renderer", "Scene", "random_scene", "rtiow_camera",
           "rtiow_final_scene", "load", "lib_path"]
