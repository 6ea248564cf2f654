"""Small renders through every kernel family of the library, meant to run under compute-sanitizer (memcheck):

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_probe.py

Frames are just large enough for the schedules that need a minimum size (cost-ordered tiles and split frames: >= 32 tiles per SM; sample-range
units: >= 8 tiles per SM).  Prints one line per case; results are checked for determinism only (the parity tests do the rest)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import venusaur_b200 as vb  # noqa: E402
from venusaur_b200 import VN_COUNTERS, VN_IMAGE_HOST, VN_NO_TONEMAP, VN_WAVEFRONT  # noqa: E402


def main():
    t0 = time.time()
    ctx = vb.Context(0)
    rt = vb.rtiow_final_scene()
    ctx.set_spheres(rt)
    ctx.build_bvh()                                                    # Morton, onesweep, Karras / SAH, refit, pack, wide build
    W, H, spp, depth = 512, 320, 2, 50
    cam = vb.rtiow_camera(W, H)
    img = np.zeros((H, W, 4), np.uint8)

    def frame(sub, count, flags=0):
        ctx.render(ctx.make_params(cam, W, H, spp, sub, depth, accum_count=count, image=img.ctypes.data, flags=flags | VN_IMAGE_HOST))
        return ctx.stats()

    for k in range(3):                                                  # collecting launch, sort, cost-ordered split frame
        st = frame(1 + k, k)
    a = ctx.read_accum().copy()
    print("shared-memory path kernel, 3 progressive frames (tile order + split frame): %d segments in the last, %d launches, accel %d" % (st.segments, st.kernel_launches, ctx.last_accel()))
    ctx.render_subframes(ctx.make_params(cam, W, H, spp, 1, depth, accum_count=0, image=img.ctypes.data, flags=VN_IMAGE_HOST), 3)
    b = ctx.read_accum()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "3 subframes in one call differ from 3 calls"
    print("vn_render_subframes (3 subframes in one launch): same buffer")
    ctx.set_option("steal_smem", 1)
    frame(1, 0); frame(2, 1); st = frame(3, 2)
    assert np.array_equal(a.view(np.uint32), ctx.read_accum().view(np.uint32)), "stealing drain changed the buffer"
    ctx.set_option("steal_smem", 0)
    print("stealing drain on the shared-memory kernel: same buffer")
    st = frame(4, 0, VN_COUNTERS)
    print("instrumented launch: %d node visits, %d sphere tests" % (st.node_visits, st.sphere_tests))
    st = frame(4, 0, VN_WAVEFRONT)
    print("wavefront kernel: %d segments" % st.segments)
    moved = rt.copy()
    moved.view(np.float32).reshape(len(rt), -1)[1:, 1] += 0.01
    ctx.update_spheres(moved)                                           # refit-only rebuild
    st = frame(5, 0)
    print("after vn_update_spheres (refit): %d segments" % st.segments)

    # a scene traversed from L2 / HBM: pair nodes (quantised and plain), stealing drain, sample-range units
    n, S = 20000, 30.0
    ctx.set_spheres(vb.random_scene(n, 0x5EED0001, S, 1))
    ctx.build_bvh()
    assert ctx.bvh_info().scene_in_smem == 0
    W, H, spp, depth = 256, 160, 4, 64
    cam = vb.Camera((0.0, 0.0, 2.0 * S), 40.0, W / H, 0.0, 2.0 * S)
    cam.SetForward((0.0, 0.0, -1.0))
    img = np.zeros((H, W, 4), np.uint8)
    ctx.set_option("units_min_seg", 0)
    outs = []
    for q, units, steal in ((1, 4, 1), (0, 4, 1), (1, 1, 0), (1, 4, 0)):
        ctx.set_option("qnodes", q); ctx.build_bvh()
        ctx.set_option("units", units); ctx.set_option("steal", steal)
        for k in range(3):
            st = frame(1 + k, k)
        outs.append(ctx.read_accum().copy())
        print("L2/HBM path kernel qnodes=%d units=%d steal=%d: %d segments in the last frame" % (q, units, steal, st.segments))
    for o in outs[1:]:
        assert np.array_equal(o.view(np.uint32), outs[0].view(np.uint32)), "L2/HBM variants disagree"
    st = frame(9, 0, VN_WAVEFRONT)
    print("wavefront kernel on pair nodes: %d segments" % st.segments)
    ctx.close()
    print("sanitize_probe done in %.1f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
