"""Tail / launch-length experiment: Mrays/s of one vn_render as a function of samples per launch and frame size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP

ctx = vb.Context(0)
ctx.set_spheres(vb.rtiow_final_scene())
ctx.build_bvh()
for (W, H) in ((1920, 1080), (3840, 2160), (960, 540)):
    cam = vb.rtiow_camera(W, H)
    for spp in (4, 16, 64):
        best = 0.0
        for rep in range(3):
            ctx.render(ctx.make_params(cam, W, H, spp, rep + 1, 50, flags=VN_NO_TONEMAP))
            st = ctx.stats()
            best = max(best, st.segments / (st.ms_render * 1e-3) / 1e6)
        print("%dx%d spp %3d: %.0f Mrays/s  (%.3f ms)" % (W, H, spp, best, st.ms_render), flush=True)
ctx.close()
