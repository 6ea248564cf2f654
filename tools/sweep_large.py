#!/usr/bin/env python
"""A/B sweep of path-kernel options on the scenes traversed from L2 / HBM (BASELINE configs[3] / configs[4]) inside one process:
one scene upload + BVH build, then per option set one warm-up frame and N timed frames at 1920x1080, 16 spp.  Prints Mrays/s and
whether the accumulation buffer equals the first set's bit for bit.  GPU only.

    python tools/sweep_large.py c4 "lean=0" "lean=1" "lean=1,async_done=8" ...
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import venusaur_b200 as vb  # noqa: E402
from venusaur_b200 import VN_COUNTERS, VN_NO_TONEMAP  # noqa: E402

REBUILD = {"leaf_size", "wide_max_prims", "aabb_pad"}


def main():
    workload = sys.argv[1]
    sets = sys.argv[2:] or ["lean=1"]
    frames = int(os.environ.get("SWEEP_FRAMES", "2"))
    W, H, SPP, DEPTH, scene = bench.WORKLOADS[workload]
    n, seed, S, mix = bench.SCENES[scene]
    ctx = vb.Context(0)
    ctx.set_spheres(vb.random_scene(n, seed, S, mix))
    ctx.build_bvh()
    cam = vb.Camera((0.0, 0.0, 2.0 * S), 40.0, W / H, 0.0, 2.0 * S)
    cam.SetForward((0.0, 0.0, -1.0))
    ref, current = None, {}
    for spec in sets:
        opts = dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in spec.split(",") if kv)
        rebuild = False
        for k, v in opts.items():
            if current.get(k) != v:
                ctx.set_option(k, v)
                current[k] = v
                rebuild |= k in REBUILD
        if rebuild:
            ctx.build_bvh()
        ms, seg = [], 0
        for f in range(1 + frames):
            ctx.render(ctx.make_params(cam, W, H, SPP, 1 + f, DEPTH, flags=VN_NO_TONEMAP))
            st = ctx.stats()
            if f >= 1:
                ms.append(st.ms_render)
                seg += st.segments
        ctx.render(ctx.make_params(cam, W, H, SPP, 1, DEPTH, flags=VN_NO_TONEMAP | VN_COUNTERS))
        st = ctx.stats()
        acc = ctx.read_accum()
        same = True
        if ref is None:
            ref = acc.copy()
        else:
            same = bool(np.array_equal(acc.view(np.uint32), ref.view(np.uint32)))
        print(json.dumps({"workload": workload, "opts": spec, "mrays_s": round(seg / (sum(ms) * 1e-3) / 1e6, 1), "ms_frame": round(float(np.mean(ms)), 2),
                          "accel": ctx.last_accel(), "nodes_per_seg": round(st.node_visits / st.segments, 2), "spheres_per_seg": round(st.sphere_tests / st.segments, 2),
                          "bit_identical_to_first": same}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
