#!/usr/bin/env python
"""SIMT scheduling model: replays real per-ray BVH step sequences (tests/host_harness.cpp::hh_step_sequences) through
different warp-level schedules and reports warp-instructions per ray segment.  Used to choose the traversal loop
structure without spending GPU time; calibrated against ncu (persistent kernel, if/else loop: ~96-110 measured)."""
import sys
import numpy as np

N_NODE, LEAF_BASE, LEAF_PER, SHADE, CAMERA, FETCH, LOOP = 62, 14, 38, 230, 110, 30, 6


def run(seq, policy, K=0, vote_T=8, n_warps=64, seed=0):
    pix_starts = np.concatenate([[0], np.nonzero(seq == 254)[0][:-1] + 1])
    # 8x4-tile-ish locality is irrelevant here; hand pixels out in order
    n_pix = len(pix_starts)
    nxt = 0
    ptr = np.zeros((n_warps, 32), np.int64)
    alive = np.zeros((n_warps, 32), bool)
    for w in range(n_warps):
        for l in range(32):
            if nxt < n_pix:
                ptr[w, l] = pix_starts[nxt]; alive[w, l] = True; nxt += 1
    cost = 0
    segs = 0
    idle_lane_steps = 0
    # simulate each warp independently but share the pixel ticket
    warps = list(range(n_warps))
    while warps:
        still = []
        for w in warps:
            P = ptr[w]; A = alive[w]
            if not A.any():
                continue
            cur = np.where(A, seq[P], 250)
            # pixel ends -> fetch
            m = cur == 254
            if m.any():
                cost += FETCH
                for l in np.nonzero(m)[0]:
                    if nxt < n_pix:
                        P[l] = pix_starts[nxt]; nxt += 1
                    else:
                        A[l] = False
                cur = np.where(A, seq[P], 250)
            m = cur == 253
            if m.any():
                cost += CAMERA
                P[m] += 1
                cur = np.where(A, seq[P], 250)
            # traversal phase
            it = 0
            while True:
                node = cur == 0
                leaf = (cur >= 1) & (cur <= 8)
                if not (node.any() or leaf.any()):
                    break
                if K and it >= K:
                    break
                it += 1
                if policy == "ifelse":
                    c = LOOP
                    if node.any(): c += N_NODE
                    if leaf.any(): c += LEAF_BASE + LEAF_PER * int(cur[leaf].max())
                    cost += c
                    adv = node | leaf
                    P[adv] += 1
                elif policy == "vote":
                    nl = int(leaf.sum())
                    if node.any() and nl < vote_T:
                        cost += LOOP + 4 + N_NODE; P[node] += 1
                    else:
                        cost += LOOP + 4 + LEAF_BASE + LEAF_PER * int(cur[leaf].max()); P[leaf] += 1
                elif policy == "whilewhile":
                    # node phase until every lane is at a leaf / done
                    while node.any():
                        cost += N_NODE + 2; P[node] += 1
                        cur = np.where(A, seq[P], 250); node = cur == 0
                    # leaf phase: every lane tests all the leaves it holds (consecutive leaf tokens)
                    leaf = (cur >= 1) & (cur <= 8)
                    while leaf.any():
                        cost += LOOP + LEAF_BASE + LEAF_PER * int(cur[leaf].max()); P[leaf] += 1
                        cur = np.where(A, seq[P], 250); leaf = (cur >= 1) & (cur <= 8)
                cur = np.where(A, seq[P], 250)
            m = cur == 255
            if m.any():
                cost += SHADE
                segs += int(m.sum())
                P[m] += 1
            still.append(w)
        warps = still
    return cost / max(segs, 1), segs


if __name__ == "__main__":
    seq = np.load(sys.argv[1])
    lim = int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
    seq = seq[:np.nonzero(seq[:lim] == 254)[0][-1] + 1]
    nseg = int((seq == 255).sum())
    ideal = (N_NODE * (seq == 0).sum() + (LEAF_BASE * ((seq >= 1) & (seq <= 8)).sum() + LEAF_PER * seq[(seq >= 1) & (seq <= 8)].sum()) + SHADE * nseg + CAMERA * (seq == 253).sum()) / nseg / 32
    print("segments", nseg, "ideal (100%% lanes) warp-inst/segment %.1f" % ideal)
    for pol, kw in [("ifelse", {}), ("whilewhile", {}), ("vote", {"vote_T": 4}), ("vote", {"vote_T": 8}), ("vote", {"vote_T": 16}),
                    ("ifelse", {"K": 4}), ("ifelse", {"K": 8}), ("ifelse", {"K": 12}), ("ifelse", {"K": 16}), ("vote", {"K": 8, "vote_T": 8}), ("vote", {"K": 12, "vote_T": 8})]:
        c, s = run(seq, pol, n_warps=32, **kw)
        print("%-11s %-24s warp-inst/segment %6.1f  (lane efficiency %.0f%%)" % (pol, kw, c, 100 * ideal / c))
