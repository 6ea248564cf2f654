"""Duration of a launch of the shared-memory path kernel that has (almost) nothing to render: what a CTA spends staging the scene (the 8 octant
copies of the wide nodes + spheres) before its first ray.  One 8x4-pixel tile, 1 spp, depth 1; vn_stats.ms_render = event time around the launch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb  # noqa: E402
from venusaur_b200 import VN_NO_TONEMAP  # noqa: E402

ctx = vb.Context(0)
ctx.set_spheres(vb.rtiow_final_scene())
ctx.build_bvh()
for (W, H) in ((8, 4), (1920, 1080)):
    cam = vb.rtiow_camera(W, H)
    ms = []
    for k in range(12):
        ctx.render(ctx.make_params(cam, W, H, 1, 1 + k, 1, flags=VN_NO_TONEMAP))
        ms.append(ctx.stats().ms_render)
    print("%dx%d, 1 spp, depth 1: ms_render min %.4f median %.4f (accel %d)" % (W, H, min(ms[2:]), float(np.median(ms[2:])), ctx.last_accel()))
ctx.close()
