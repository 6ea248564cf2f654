"""World-size-2 test of the multi-GPU host logic on CPU (gloo): sample-range sharding + one reduce per frame gives the same
image as one process accumulating every subframe (the per-rank renders come from the oracle here; on GPUs they come from
vn_render with VN_ACCUM_SUM, which tests/test_gpu_parity.py::test_row_tiles_and_partial_sums checks)."""
import os
import socket

import numpy as np
import pytest

from venusaur_b200 import sharding

W, H, SPP, DEPTH, STEPS, WORLD = 32, 18, 2, 8, 3, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = ol.Oracle(ol.rtiow_final_scene())
    cam = ol.rtiow_camera(W, H)
    acc = np.zeros((H, W, 4), np.float32)
    for sub in sharding.subframes_for_rank(rank, world, STEPS):
        mean, _ = orc.render_mean(orc.params(cam, W, H, SPP, sub, DEPTH, threads=1))
        acc += mean                                   # what VN_ACCUM_SUM does on the device
    t = torch.from_numpy(acc)
    mean = sharding.reduce_partial_sums(t, STEPS * world, dist, dst=0)
    if rank == 0:
        np.save(out_path, mean.numpy())
    else:
        assert mean is None
    dist.barrier()
    dist.destroy_process_group()


def test_subframe_dealing_covers_every_stream_once():
    for world in (1, 2, 4, 8):
        ids = sorted(s for r in range(world) for s in sharding.subframes_for_rank(r, world, 5))
        assert ids == list(range(1, 5 * world + 1))
        rows = [sharding.row_slice(r, world, 1080) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 1080 and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
    assert sharding.row_slice(3, 8, 2160) == (810, 1080)


@pytest.mark.timeout(300)
def test_two_rank_reduce_equals_single_process(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    out = str(tmp_path / "mean.npy")
    mp.spawn(_worker, args=(WORLD, _free_port(), out), nprocs=WORLD, join=True)
    got = np.load(out)
    orc = oracle_mod.Oracle(oracle_mod.rtiow_final_scene())
    cam = oracle_mod.rtiow_camera(W, H)
    want = np.zeros((H, W, 4), np.float32)
    for k in range(STEPS * WORLD):                    # the reference's sequential running mean over subframes 1..6
        mean, _ = orc.render_mean(orc.params(cam, W, H, SPP, k + 1, DEPTH, threads=1))
        want, _ = oracle_mod.accumulate_tonemap(want, mean, k > 0, np.float32(1.0) / np.float32(k + 1))
    rel = np.abs(got[..., :3] - want[..., :3]) / np.maximum(np.abs(want[..., :3]), 1e-6)
    assert rel.max() < 2e-6                           # float re-association only
