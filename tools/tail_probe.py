#!/usr/bin/env python
"""How much of a launch is drain?  Runs the instrumented path kernel (VN_COUNTERS) on the headline workload and reads the launch
timeline it leaves in the scheduler-statistics words: start, first lane that found the ticket counter exhausted, last warp's end.
The first launch of a view runs with row-major tickets (and counts the tiles' costs), the following ones with the cost-ordered
tiles (vn_api.cu::prepare_tile_order).  GPU only."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_COUNTERS, VN_NO_TONEMAP

def main():
    ctx = vb.Context(0)
    ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
    for (W, H) in ((1920, 1080), (3840, 2160)):
        cam = vb.rtiow_camera(W, H)
        for rep in range(3):
            ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, flags=VN_COUNTERS | VN_NO_TONEMAP))
            st = ctx.stats()
            start, exhaust, end = ctx.launch_timeline()
            print("%dx%d launch %d: ms_render %.3f | kernel %.3f ms, tickets exhausted at %.3f ms (%.1f %%), drain %.3f ms"
                  % (W, H, rep, st.ms_render, (end - start) / 1e6, (exhaust - start) / 1e6, 100.0 * (exhaust - start) / (end - start),
                     (end - exhaust) / 1e6))
    ctx.close()

if __name__ == "__main__":
    main()
