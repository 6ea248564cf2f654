#!/usr/bin/env python
"""A/B sweep of path-kernel options inside ONE process (one context, one BVH build per accel setting): for every option set,
2 warm-up frames and N timed frames of the headline workload (RTIOW 1920x1080, 16 spp, depth 50), timed by the library's own CUDA
events around the render (vn_stats.ms_render).  Prints Mrays/s per set and checks that every set reproduces the first set's
accumulation buffer bit for bit.  GPU only.

    python tools/sweep_options.py "async_done=0" "async_done=24,async_node=8,async_leaf=8" ...
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb  # noqa: E402
from venusaur_b200 import VN_NO_TONEMAP  # noqa: E402

W, H, SPP, DEPTH, FRAMES = 1920, 1080, 16, 50, int(os.environ.get("SWEEP_FRAMES", "6"))
REBUILD = {"accel", "leaf_size", "huge_factor", "sah_max_prims", "wide_max_prims"}


def main():
    sets = sys.argv[1:] or ["async_done=0"]
    ctx = vb.Context(0)
    ctx.set_spheres(vb.rtiow_final_scene())
    ctx.build_bvh()
    cam = vb.rtiow_camera(W, H)
    ref = None
    current = {}
    out = []
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
        for f in range(2 + FRAMES):
            ctx.render(ctx.make_params(cam, W, H, SPP, 1 + f, DEPTH, flags=VN_NO_TONEMAP))
            st = ctx.stats()
            if f >= 2:
                ms.append(st.ms_render)
                seg += st.segments
        ctx.render(ctx.make_params(cam, W, H, SPP, 1, DEPTH, flags=VN_NO_TONEMAP))
        acc = ctx.read_accum()
        same = True
        if ref is None:
            ref = acc.copy()
        else:
            same = bool(np.array_equal(acc.view(np.uint32), ref.view(np.uint32)))
        rate = seg / (sum(ms) * 1e-3) / 1e6
        line = {"opts": spec, "mrays_s": round(rate, 1), "ms_frame": round(float(np.mean(ms)), 3), "ms_min": round(float(np.min(ms)), 3),
                "accel": ctx.last_accel(), "bit_identical_to_first": same}
        out.append(line)
        print(json.dumps(line), flush=True)
    ctx.close()
    best = max(out, key=lambda d: d["mrays_s"])
    print("BEST", json.dumps(best))


if __name__ == "__main__":
    main()
