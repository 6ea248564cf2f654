"""One vn_render_subframes launch of 8 subframes on the headline workload, for `ncu -k regex:k_render_lean -s <launches before it>`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP
W, H = 1920, 1080
ctx = vb.Context(0)
ctx.set_option("split_tail", 0)
ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
cam = vb.rtiow_camera(W, H)
for rep in range(2):                                            # launches 0, 1: tile costs, tile order
    ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, accum_count=rep, flags=VN_NO_TONEMAP))
ctx.render_subframes(ctx.make_params(cam, W, H, 16, 3, 50, accum_count=2, flags=VN_NO_TONEMAP), 8)      # launch 2
print("ms", ctx.stats().ms_render, "segments", ctx.stats().segments)
