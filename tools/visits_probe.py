"""CPU probe of the traversal work per ray segment (V_node, V_sphere) of the shared-memory wide-node path on RTIOW, with the product's own
builder and traversal compiled for the host (tests/host_harness.cpp; the counts equal the GPU's: 5.00 wide-node visits and 2.01 sphere
tests per segment), plus a costing of the 4-wide collapse: sum of the wide nodes' box areas over the root's (the expected visits of a
random ray) for the shipped greedy policy (lbvh_core.cuh::wide_collapse, largest area first) against the optimum of a tree DP.

    python tools/visits_probe.py [-DNAME=value ...]        # extra flags reach the host build (compile-time variants)
    CSRC=/tmp/copy_of_csrc python tools/visits_probe.py      # headers from a modified copy

Round-2 findings (all on CPU, image hash unchanged): the octant order of a node's children (box centre / entry corner / exit corner along
the octant diagonal) does not move V_node at all (5.0030: stack entries are not re-tested at the pop, so the order only decides WHEN a
subtree is visited); the optimal collapse lowers the area sum by 0.7 % (4.431 against 4.461; 220 instead of 259 wide nodes without the
ground sphere); a huge list that also takes the three unit spheres (factor 4) trades -2.5 % node visits for +2.6 sphere tests per segment.
The tree is not where the remaining time is."""
import ctypes as C
import hashlib
import os
import subprocess
import sys
import time
from functools import lru_cache

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as om  # noqa: E402
import test_host_logic as thl  # noqa: E402
import venusaur_b200 as vb  # noqa: E402


def build(flags):
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    tag = "_".join(f.replace("-D", "").replace("=", "") for f in flags) or "base"
    lib = os.path.join(out, "libhh_probe_%s.so" % tag)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w",
                    "-I" + os.environ.get("CSRC", os.path.join(ROOT, "venusaur_b200", "csrc")), *flags, "-o", lib,
                    os.path.join(ROOT, "tests", "host_harness.cpp")], check=True)
    h = C.CDLL(lib)
    h.hh_render_mean.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    h.hh_build_bvh.restype = C.c_uint64
    h.hh_build_bvh.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    return h


def visits(h, spheres, W, H, spp):
    cam = om.rtiow_camera(W, H)
    hp = thl._hh_params(cam, W, H, spp, 1, 50)
    h.hh_set_wide(1)
    mean = np.zeros((H, W, 4), np.float32)
    segs, nv, st = C.c_uint64(), C.c_uint64(), C.c_uint64()
    t = time.time()
    h.hh_render_mean(spheres.ctypes.data_as(C.c_void_p), len(spheres), 1, C.c_float(0.01), C.byref(hp), mean.ctypes.data_as(C.c_void_p),
                     C.byref(segs), C.byref(nv), C.byref(st))
    h.hh_set_wide(0)
    print("segments %d  V_node %.4f  V_sphere %.4f  image md5 %s  (%.1f s)" % (segs.value, nv.value / segs.value, st.value / segs.value,
                                                                              hashlib.md5(mean.tobytes()).hexdigest()[:10], time.time() - t))


def collapse_costs(h, spheres):
    """Area sums of the wide nodes: shipped greedy collapse vs the DP optimum (every binary node is a wide node's root or absorbed by one;
    a root absorbs at most two connected descendants)."""
    n = len(spheres)
    nodes = np.zeros(2 * n + 4, vb.api.NODE_DTYPE)
    order = np.zeros(n, np.uint32)
    codes = np.zeros(n, np.uint32)
    nn = h.hh_build_bvh(spheres.ctypes.data_as(C.c_void_p), n, 1, C.c_float(0.01), nodes.ctypes.data_as(C.c_void_p), len(nodes),
                        order.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p))
    raw = nodes[:nn].view(np.uint32).reshape(nn, 8)
    f = nodes[:nn].view(np.float32).reshape(nn, 8)
    d = f[:, 4:7] - f[:, 0:3]
    area = (2.0 * (d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 0] * d[:, 2])).tolist()
    link = raw[:, 3].astype(np.int64).tolist()
    leaf = [(x & 0x80000000) != 0 for x in link]
    sys.setrecursionlimit(100000)

    def greedy_children(pair):
        out = [pair, pair + 1]
        while len(out) < 4:
            best, ba = -1, -1.0
            for k, c in enumerate(out):
                if not leaf[c] and area[c] > ba:
                    ba, best = area[c], k
            if best < 0:
                break
            out[best:best + 1] = [link[out[best]], link[out[best]] + 1]
        return out

    def greedy(v):
        tot, cnt = area[v], 1
        for c in greedy_children(link[v]):
            if not leaf[c]:
                t, k = greedy(c)
                tot, cnt = tot + t, cnt + k
        return tot, cnt

    @lru_cache(None)
    def best(v):
        g = below(v, 2)
        return area[v] + g[0], 1 + g[1]

    @lru_cache(None)
    def child(c, budget):
        if leaf[c]:
            return 0.0, 0
        if budget == 0:
            return best(c)
        return min(best(c), below(c, budget - 1))

    @lru_cache(None)
    def below(v, budget):
        L = link[v]
        return min((child(L, j)[0] + child(L + 1, budget - j)[0], child(L, j)[1] + child(L + 1, budget - j)[1]) for j in range(budget + 1))

    g, o = greedy(1), best(1)
    print("collapse, %d spheres: greedy area sum / root %.4f (%d wide nodes), optimum %.4f (%d wide nodes)" % (n, g[0] / area[1], g[1], o[0] / area[1], o[1]))


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    h = build(flags)
    rt = np.ascontiguousarray(om.rtiow_final_scene())
    visits(h, rt, int(os.environ.get("W", 480)), int(os.environ.get("H", 270)), int(os.environ.get("SPP", 2)))
    radius = rt.view(np.float32).reshape(len(rt), -1)[:, 3]
    collapse_costs(h, np.ascontiguousarray(rt[np.abs(radius) < 100.0]))        # without the ground sphere (it lives in the huge list)
