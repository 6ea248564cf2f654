#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun with N ranks): sample-range sharding + ONE fused peer reduce+tonemap (and the
NCCL variant) must reproduce the single-GPU running mean over the same subframes: accum within float re-association
(<= 2e-6 relative), uchar4 image within one code value.  Prints one line per rank-0 check and exits non-zero on failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import venusaur_b200 as vb  # noqa: E402
from venusaur_b200 import VN_ACCUM_SUM, VN_ASYNC, VN_NO_TONEMAP, sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H, spp, depth, K = 400, 225, 16, 50, 3
    ctx = vb.Context(local)
    ctx.set_spheres(vb.rtiow_final_scene())
    ctx.build_bvh()
    cam = vb.rtiow_camera(W, H)
    ctx.resize(W, H)
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    image = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
    ctx.set_accum_external(accum.data_ptr())
    for sub in sharding.subframes_for_rank(rank, world, K):
        ctx.render(ctx.make_params(cam, W, H, spp, sub, depth, flags=VN_ACCUM_SUM | VN_NO_TONEMAP | VN_ASYNC))
    ctx.synchronize()
    partial = accum.clone()

    # fused peer reduce + tonemap
    mine = torch.tensor(list(ctx.ipc_export(accum.data_ptr())) + list(ctx.ipc_export(image.data_ptr())), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    ptrs, img0 = [], None
    for r in range(world):
        hb = bytes(allh[r].cpu().tolist())
        ptrs.append(accum.data_ptr() if r == rank else ctx.ipc_open(hb[:64]))
        if r == 0:
            img0 = image.data_ptr() if rank == 0 else ctx.ipc_open(hb[64:])
    dist.barrier()
    torch.cuda.synchronize()
    ctx.reduce_tonemap_peers(ptrs, 1.0 / (K * world), sharding.row_slice(rank, world, H), img0, 0)
    ctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    # gather the reduced row slices of the float sum on rank 0 for the check
    rows = sharding.row_slice(rank, world, H)
    summed = accum.clone()
    full = [torch.zeros_like(summed) for _ in range(world)]
    dist.all_gather(full, summed)
    # NCCL variant on the untouched partial sums
    nccl_sum = partial.clone()
    dist.reduce(nccl_sum, dst=0, op=dist.ReduceOp.SUM)
    ok = True
    if rank == 0:
        peer_sum = torch.zeros_like(summed)
        for r in range(world):
            a, b = sharding.row_slice(r, world, H)
            peer_sum[a:b] = full[r][a:b]
        peer_mean = (peer_sum[..., :3] / float(K * world)).cpu().numpy()
        nccl_mean = (nccl_sum[..., :3] / float(K * world)).cpu().numpy()
        ref = vb.Context(local)
        ref.set_spheres(vb.rtiow_final_scene())
        ref.build_bvh()
        ref.resize(W, H)
        ref_img = np.zeros((H, W, 4), np.uint8)
        for k in range(K * world):
            p = ref.make_params(cam, W, H, spp, k + 1, depth, accum_count=k, image=ref_img.ctypes.data, flags=vb.VN_IMAGE_HOST)
            ref.render(p)
        want = ref.read_accum()[..., :3]
        for name, got in (("peer", peer_mean), ("nccl", nccl_mean)):
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
            print("check_multi N=%d %s: max rel err of accum vs single GPU %.3g" % (world, name, rel.max()))
            ok &= rel.max() < 2e-6
        d = np.abs(image.cpu().numpy().astype(np.int32) - ref_img.astype(np.int32))
        print("check_multi N=%d peer image: max code-value diff %d, differing pixels %.4f%%" % (world, d.max(), 100.0 * (d.max(axis=-1) > 0).mean()))
        ok &= d.max() <= 1
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if flag.item() == 1 else 1)



if __name__ == "__main__":
    main()
