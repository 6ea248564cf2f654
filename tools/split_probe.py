"""ms_render of the headline frame for a list of option sets (one process, a dozen launches each; VN_DEBUG_SPLIT=1 prints the two launches of a
split frame).  usage: split_probe.py "k=v k=v" "k=v" ...   GPU only."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP
ctx = vb.Context(0)
ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
W, H = 1920, 1080
cam = vb.rtiow_camera(W, H)
defaults = {"split_tail": 0.5, "tile_guess": 1, "tile_order": 1, "steal_smem": 0, "steal": 1, "async_done": 26, "wide_threads": 1024, "units": 4}
for arg in sys.argv[1:] or [""]:
    opts = dict(defaults)
    for kv in arg.split():
        k, v = kv.split("="); opts[k] = float(v)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ms = []
    for rep in range(14):
        ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, flags=VN_NO_TONEMAP))
        ms.append(ctx.stats().ms_render)
    a = np.array(ms[2:])
    print("[%s]: ms_render mean %.3f min %.3f max %.3f  (first two launches %.3f %.3f)" % (arg, a.mean(), a.min(), a.max(), ms[0], ms[1]))
