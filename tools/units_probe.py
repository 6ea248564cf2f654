"""Calibration of "units_min_seg" (vn_api.cu): for random scenes of several sizes, traversed from L2 / HBM, the mean path length and the launch
time with whole pixels (units=1) and with forced sample-range units (2, 4).  usage: units_probe.py <n spheres> ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP
W, H = 1920, 1080
ctx = vb.Context(0)
ctx.set_option("units_min_seg", 0)
for n in [int(a) for a in sys.argv[1:]]:
    S = 10.0 * (n / 500.0) ** (1.0 / 3.0)
    ctx.set_spheres(vb.random_scene(n, 0x5EED0100 + n, S, 0)); ctx.build_bvh()
    cam = vb.Camera((0.0, 0.0, 2.0 * S), 40.0, W / H, 0.0, 2.0 * S)
    cam.SetForward((0.0, 0.0, -1.0))
    line = "n=%d" % n
    for units in (1, 2, 4):
        ctx.set_option("units", units)
        ms = []
        for rep in range(6):
            ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, flags=VN_NO_TONEMAP))
            st = ctx.stats(); ms.append(st.ms_render)
        if units == 1: line += " seg/path %.2f" % (st.segments / max(1, st.paths))
        line += " | units=%d %.2f ms" % (units, min(ms[2:]))
    print(line, flush=True)
