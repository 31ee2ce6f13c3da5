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


# ---- fused flush + reduce: one accumulator in rank 0's memory ------------------------------------

class _CudaIpcBackend:
    """Accumulators are plain cudaMalloc blocks of rank 0 (m3d_device_alloc), mapped into the other
    ranks' address spaces with CUDA IPC (m3d_ipc_export / m3d_ipc_open); over NVLink on a B200 box."""

    def __init__(self, ctx):
        from . import _native as N
        self.N, self.ctx = N, ctx

    def alloc(self, nbytes):
        return self.N.device_alloc(self.ctx, nbytes)

    def export(self, ptr):
        return self.N.ipc_export(self.ctx, ptr)

    def open(self, handle):
        return self.N.ipc_open(self.ctx, handle)

    def close(self, ptr):
        self.N.ipc_close(self.ctx, ptr)

    def free(self, ptr):
        self.N.device_free(self.ctx, ptr)


class SharedAccumulator:
    """`nbuf` frame accumulators that live in rank 0's memory and that EVERY rank's flush kernel
    adds into (m3d_partition.flags = M3D_PART_ATOMIC): the cross-GPU reduce of the per-pixel sums
    happens inside path_flush (red.add over the NVLink mapping), tile by tile, instead of in a
    collective after the render.  Replaces `reduce_sums` for the one-process-per-GPU launch.

    Protocol per frame k (all on the ranks' streams, no host synchronisation):
        every rank     render + flush into ptr(k)
        every rank     barrier()                    (a 4-byte all_reduce: every flush has landed)
        rank 0         consume ptr(k), then clear it
    With nbuf >= 2 a rank that runs ahead flushes frame k+1 into the other buffer, which rank 0
    cleared before it entered barrier k, so one barrier per frame is enough."""

    def __init__(self, nbytes, rank, world, backend, nbuf=2, src=0):
        self.rank, self.world, self.backend, self.nbuf, self.src = rank, world, backend, nbuf, src
        self.nbytes = int(nbytes)
        self.owner = rank == src
        handles = [None] * nbuf
        self.ptrs = []
        err = None
        if self.owner:
            try:
                self.ptrs = [backend.alloc(self.nbytes) for _ in range(nbuf)]
                if world > 1:
                    handles = [backend.export(p) for p in self.ptrs]
            except Exception as e:  # the other ranks wait in the broadcast: tell them before raising
                err, handles = e, [None] * nbuf
        if world > 1:
            dist.broadcast_object_list(handles, src=src)
            if not self.owner:
                if handles[0] is None:
                    raise RuntimeError("rank %d could not export its accumulators" % src)
                self.ptrs = [backend.open(h) for h in handles]
        if err is not None:
            raise err

    def ptr(self, k):
        return self.ptrs[k % self.nbuf]

    def barrier(self, device=None):
        """Cross-rank ordering point on the current stream (NCCL) or the host (gloo)."""
        if self.world > 1:
            t = torch.zeros(1, dtype=torch.float32, device=device)
            dist.all_reduce(t)

    def close(self):
        for p in self.ptrs:
            (self.backend.free if self.owner else self.backend.close)(p)
        self.ptrs = []


class DevicePointer:
    """Wraps a raw device pointer for torch.as_tensor (the __cuda_array_interface__ protocol)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
