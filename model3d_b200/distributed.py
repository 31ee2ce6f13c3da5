"""Multi-GPU partitioning of the ray-tracing path: one process per GPU (torch.distributed).

The reference treats every pixel as an independent work item scheduled over goroutines
(render3d/concurrency.go:17-43).  Across GPUs the same independence is used three ways
(SURVEY 8e), none of which needs a data-path collective except the final image reduce:

  raw ray batches   contiguous slices of the ray array per rank         (no collective)
  RayCaster         row bands [row_begin, row_end) per rank             (sum of disjoint bands)
  path tracers      sample-index shards of EVERY pixel per rank         (sum of per-pixel sums)

The only exchange step is one reduce (sum) of the W*H*3 float32 accumulators to rank 0 --
NCCL over NVLink on GPUs, gloo in the CPU tests of this host logic.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def split_even(n: int, world: int) -> List[Tuple[int, int]]:
    """[begin, end) of every rank: sizes differ by at most one, earlier ranks get the extra."""
    if world <= 0 or n < 0:
        raise ValueError("split_even needs world > 0 and n >= 0")
    base, extra = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def ray_slice(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Raw ray batch (BASELINE config 2): contiguous slice of the ray arrays."""
    return split_even(n_rays, world)[rank]


def row_band(height: int, rank: int, world: int) -> Tuple[int, int, int]:
    """RayCaster partition (row_begin, row_end, sample_begin=0) for m3d_partition."""
    b, e = split_even(height, world)[rank]
    return (b, e, 0)


def sample_shard(num_samples: int, rank: int, world: int) -> Tuple[Tuple[int, int, int], int]:
    """Path-tracer partition ((0, 0, sample_begin), sample_count): all rows, a contiguous
    range of absolute sample indices (the Philox stream is keyed by (pixel, sample), so the
    reduced image does not depend on `world`)."""
    b, e = split_even(num_samples, world)[rank]
    return (0, 0, b), e - b


def reduce_sums(acc: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Sum the per-rank accumulators into rank `dst` (in place).  No-op without a process
    group.  Row bands are disjoint and zero elsewhere, so the same sum gathers them."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)
    return acc


def finalize_mean(acc: torch.Tensor, total_samples: int) -> torch.Tensor:
    """colorSum.Scale(1/numSamples) (render3d/ray_renderer.go:150) after the reduce."""
    return acc / float(total_samples)
