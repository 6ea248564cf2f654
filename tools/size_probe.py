"""Throughput against scene size: random scenes of n spheres (bench.py's generator, density kept), 1080p, 16 spp, depth 50.  Shows where the
closest-hit structure changes (4-wide nodes in shared memory -> pair nodes in shared memory -> quantised pairs from L2 / HBM).  GPU only.
usage: size_probe.py [n ...] [key=value options]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP
ns = [int(a) for a in sys.argv[1:] if a.isdigit()] or [300, 480, 600, 1000, 2000, 5000, 20000, 100000]
opts = [a for a in sys.argv[1:] if "=" in a]
ctx = vb.Context(0)
for kv in opts:
    k, v = kv.split("="); ctx.set_option(k, float(v))
W, H = 1920, 1080
for n in ns:
    S = 10.0 * (n / 500.0) ** (1.0 / 3.0)                       # half-extent: the density of a 500-sphere scene in a 20-unit cube
    ctx.set_spheres(vb.random_scene(n, 0x5EED0100 + n, S, 0)); ctx.build_bvh()
    cam = vb.Camera((0.0, 0.0, 2.0 * S), 40.0, W / H, 0.0, 2.0 * S)
    cam.SetForward((0.0, 0.0, -1.0))
    ms, segs = [], 0
    for rep in range(5):
        ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, flags=VN_NO_TONEMAP))
        st = ctx.stats(); ms.append(st.ms_render); segs = st.segments
    info = ctx.bvh_info()
    print("n=%7d: %.2f ms/launch, %6.0f Mrays/s (%.1f segments/path), accel %d (1 pairs, 2 wide nodes in shared memory), scene_in_smem %d"
          % (n, np.mean(ms[2:]), segs / (np.mean(ms[2:]) * 1e-3) / 1e6, segs / (W * H * 16.0), ctx.last_accel(), info.scene_in_smem))
