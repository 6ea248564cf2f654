"""Multi-GPU partitioning of the hot path (host logic only; one process per GPU).

The seed of a pixel is a pure function of (pixel, subframe_index) (RayTracer.cu:169), so any rank can render any
(row tile, subframe) pair and obtain exactly the samples a single GPU would.  Scene and BVH are replicated.
  * sample-range sharding: rank r of N renders subframes r+1, r+1+N, ... into a per-rank partial SUM buffer
    (VN_ACCUM_SUM); there is no collective on the data path.
  * once per frame the partial sums are combined: either torch.distributed.reduce (NCCL over NVLink) followed by the
    tonemap kernel on rank 0, or the fused peer-memory kernel vn_reduce_tonemap_peers in which every rank reduces and
    tonemaps its own row slice.  sum-then-divide differs from the reference's sequential lerp (RayTracer.cu:208-213)
    by float re-association only.
"""
from __future__ import annotations


def subframes_for_rank(rank: int, world: int, steps: int) -> list[int]:
    """1-based RNG stream ids (Renderer.h:54 increments before the launch) rendered by `rank` in `steps` steps."""
    return [1 + rank + k * world for k in range(steps)]


def row_slice(rank: int, world: int, height: int) -> tuple[int, int]:
    """Rows [begin, end) of the final image that `rank` reduces + tonemaps in the fused peer kernel."""
    return (height * rank) // world, (height * (rank + 1)) // world


def reduce_partial_sums(accum_sum, total_subframes: int, dist=None, dst: int = 0):
    """Combines per-rank partial sums (a torch tensor, float32, H x W x 4) into the frame mean on rank `dst`.
    Returns the mean on dst and None elsewhere.  `dist` is torch.distributed (or None for a single process)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum_sum, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return None
    return accum_sum * (1.0 / float(total_subframes))
