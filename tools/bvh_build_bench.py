#!/usr/bin/env python
"""LBVH build timing on large synthetic scenes (BASELINE configs[3], [4]) + a sanity render of the 16 M-sphere scene."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import venusaur_b200 as vb  # noqa: E402
from venusaur_b200 import VN_COUNTERS, VN_NO_TONEMAP  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [1_000_000, 16_000_000]
ctx = vb.Context(0)
for n in sizes:
    mix, S, seed, depth = (0, 100.0, 0x5EED0001, 50) if n <= 2_000_000 else (1, 250.0, 0x5EED0002, 64)
    t0 = time.time()
    spheres = vb.random_scene(n, seed, S, mix)
    t_gen = time.time() - t0
    ctx.set_spheres(spheres)
    up = ctx.stats().ms_upload
    best = 1e9
    for _ in range(3):
        ctx.build_bvh()
        best = min(best, ctx.stats().ms_build)
    info = ctx.bvh_info()
    print("n=%d: scene gen %.1fs (host), upload %.2f ms, LBVH build %.2f ms (%.0f Mprims/s), %d nodes (%.1f MB), in_smem=%d" %
          (n, t_gen, up, best, n / best / 1e3, info.num_nodes, info.num_nodes * 32 / 1e6, info.scene_in_smem), flush=True)
    W, H = 480, 270
    cam = vb.Camera((0.0, 0.0, 2 * S), 40.0, W / H, 0.0, 2 * S)
    cam.SetForward((0.0, 0.0, -1.0))
    ctx.resize(W, H)
    ctx.render(ctx.make_params(cam, W, H, 4, 1, depth, flags=VN_NO_TONEMAP | VN_COUNTERS))
    st = ctx.stats()
    acc = ctx.read_accum()
    assert np.isfinite(acc).all()
    print("   render %dx%d 4spp depth %d: %.2f ms, %.0f Mrays/s, %.2f segments/path, %.1f nodes/seg, %.2f spheres/seg, mean radiance %.4f" %
          (W, H, depth, st.ms_render, st.segments / st.ms_render / 1e3, st.segments / st.paths, st.node_visits / st.segments,
           st.sphere_tests / st.segments, float(acc[..., :3].mean())), flush=True)
ctx.close()
