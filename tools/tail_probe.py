#!/usr/bin/env python
"""How much of a launch is drain, and what is in it?  Runs the instrumented path kernel (VN_COUNTERS) on the headline workload and reads
the launch timeline it leaves behind: start, first lane that found the tile tickets exhausted, last warp's end (vn_read_sched_counters),
and the histogram of lane retirements (vn_read_timeline: lanes that ran out of work per 8 us bin, and -- in the cost-collecting first
launch of a view -- the ray segments of the last pixel each of them finished).  The first launch of a view runs with row-major tickets
(and counts the tiles' costs), the following ones with the cost-ordered tiles (vn_api.cu::prepare_tile_order).  GPU only."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb
from venusaur_b200 import VN_COUNTERS, VN_NO_TONEMAP

def main():
    ctx = vb.Context(0)
    for kv in sys.argv[1:]:
        k, v = kv.split("=")
        if k in ("profile", "only", "scene"):
            continue
        ctx.set_option(k, float(v))
    big = [kv for kv in sys.argv[1:] if kv.startswith("scene=")]              # scene=1m | 16m: bench.py's synthetic scenes (L2 / HBM traversal)
    if big:
        n, seed, S, mix = {"1m": (1_000_000, 0x5EED0001, 100.0, 0), "16m": (16_000_000, 0x5EED0002, 250.0, 1)}[big[0].split("=")[1]]
        ctx.set_spheres(vb.random_scene(n, seed, S, mix)); ctx.build_bvh()
    else:
        ctx.set_spheres(vb.rtiow_final_scene()); ctx.build_bvh()
    for (W, H) in (((1920, 1080),) if ("only=1080" in sys.argv or big) else ((1920, 1080), (3840, 2160))):
        cam = vb.rtiow_camera(W, H)
        if big:
            cam = vb.Camera((0.0, 0.0, 2.0 * S), 40.0, W / H, 0.0, 2.0 * S)
            cam.SetForward((0.0, 0.0, -1.0))
        for rep in range(3):
            ctx.render(ctx.make_params(cam, W, H, 16, 1 + rep, 64 if big else 50, flags=VN_COUNTERS | VN_NO_TONEMAP))
            st = ctx.stats()
            start, exhaust, end = ctx.launch_timeline()
            print("%dx%d launch %d: ms_render %.3f | kernel %.3f ms, tickets exhausted at %.3f ms (%.1f %%), drain %.3f ms"
                  % (W, H, rep, st.ms_render, (end - start) / 1e6, (exhaust - start) / 1e6, 100.0 * (exhaust - start) / (end - start),
                     (end - exhaust) / 1e6))
            lanes, segs, tiles, shaded = ctx.timeline_ex()
            if lanes.sum():
                total = int(lanes.sum())
                nz = np.nonzero(lanes)[0]
                first, last = int(nz[0]), int(nz[-1])
                alive = total - np.cumsum(lanes)              # lanes still working after each bin
                idle_lane_us = float(alive[first:last + 1].sum()) * 8.192          # lane-time still busy after the first retirement
                waste = (float((last - first + 1) * total) * 8.192 - idle_lane_us)    # lane-time spent retired before the launch ends
                print("   lanes retire between %.3f and %.3f ms; retired lane-time %.1f %% of the launch's lane-time"
                      % (first * 8.192e-3, (last + 1) * 8.192e-3, 100.0 * waste / (total * (last + 1) * 8.192)))
                step = max(1, (last - first + 1) // 12)
                for b in range(first, last + 1, step):
                    n = int(lanes[b:b + step].sum())
                    sg = int(segs[b:b + step].sum())
                    print("   %.3f ms: %6d lanes retire (%5.1f %% still busy)%s" % (b * 8.192e-3, n, 100.0 * alive[min(b + step - 1, last)] / total,
                                                                                  ("  mean segments of their last pixel %.0f" % (sg / n)) if (n and sg) else ""))
            if shaded.sum() and "profile=1" in sys.argv:
                nzs = np.nonzero(shaded)[0]
                lastb = int(nzs[-1])
                stepb = max(1, (lastb + 1) // 40)
                print("   throughput over time (bin start ms: Msegments/s, tiles fetched)")
                for b in range(0, lastb + 1, stepb):
                    print("   %.3f ms: %8.0f Mseg/s  %6d tiles" % (b * 8.192e-3, shaded[b:b + stepb].sum() / (min(stepb, lastb + 1 - b) * 8.192), int(tiles[b:b + stepb].sum())))
                tail0 = max(0, int(np.nonzero(lanes)[0][0]) - 12) if lanes.sum() else 0
                print("   the end of the launch, bin by bin")
                for b in range(tail0, lastb + 1):
                    print("   %.3f ms: %8.0f Mseg/s  %6d tiles  %6d lanes retire" % (b * 8.192e-3, shaded[b] / 8.192, int(tiles[b]), int(lanes[b])))
    ctx.close()

if __name__ == "__main__":
    main()
