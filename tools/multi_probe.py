"""Subframes per launch (vn_render_subframes, "multi_subframes") on the headline workload: RTIOW 1920x1080, 16 spp per subframe, depth 50.
Wall time of 64 subframes with 1, 2, 4, 8, 16, 32, 64 subframes per launch (host synchronised at both ends).  usage: multi_probe.py [W H] [option=value ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_NO_TONEMAP, VN_ASYNC
args = [a for a in sys.argv[1:] if "=" not in a]
W, H = (int(args[0]), int(args[1])) if len(args) > 1 else (1920, 1080)
ctx = vb.Context(0)
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("="); ctx.set_option(k, float(v))
ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
cam = vb.rtiow_camera(W, H)
N = 64
for rep in range(3):                                            # the view's tile costs, clocks
    ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 50, accum_count=rep, flags=VN_NO_TONEMAP))
for per in ((1, 64, 1, 64) if any('=' in a for a in sys.argv[1:]) else (1, 64, 1, 2, 4, 8, 16, 32, 64)):
    ctx.set_option("multi_subframes", per)
    ctx.synchronize(); ctx.reset_stats()
    t0 = time.perf_counter()
    if per == 1:
        for k in range(N):
            ctx.render(ctx.make_params(cam, W, H, 16, 1 + k, 50, accum_count=k, flags=VN_NO_TONEMAP | VN_ASYNC))
    else:
        ctx.render_subframes(ctx.make_params(cam, W, H, 16, 1, 50, accum_count=0, flags=VN_NO_TONEMAP | VN_ASYNC), N)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    st = ctx.stats()
    print("%2d subframes per launch: %.2f ms for %d subframes = %.3f ms each, %.0f Mrays/s" % (per, dt * 1e3, N, dt * 1e3 / N, st.segments_total / dt / 1e6), flush=True)
