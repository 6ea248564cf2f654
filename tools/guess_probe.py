"""An orbiting camera (a new view every second launch): what the collecting launch of a view costs with row-major tickets (tile_guess=0) and
with the previous view's tile order (tile_guess=1, the default).  GPU only."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP
ctx = vb.Context(0)
ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
W, H = 1920, 1080
for guess in (0, 1, 0, 1):
    ctx.set_option("tile_guess", guess)
    ctx.set_option("tile_order", 1)
    ms = []
    for rep in range(16):
        # an orbiting camera: a new view every second launch (collecting launch + one ordered launch per view)
        a = 0.002 * (rep // 2)
        cam = vb.Camera((13.0 * np.cos(a) - 3.0 * np.sin(a), 2.0, 13.0 * np.sin(a) + 3.0 * np.cos(a)), 20.0, W / H, 0.1, 10.0)
        cam.SetForward((-cam.m_position[0], -cam.m_position[1], -cam.m_position[2]))
        ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, flags=VN_NO_TONEMAP))
        ms.append(ctx.stats().ms_render)
    a = np.array(ms[2:])
    print("tile_guess=%d: collecting launches %.3f ms, ordered launches %.3f ms" % (guess, a[0::2].mean(), a[1::2].mean()))
