"""Per-shard render time of a path-tracing workload on ONE GPU: the whole frame, then the shard of rank
0 of `world` (plain flush and PART_ATOMIC flush into a local accumulator).  Checks that a shard costs
1 / world of the frame.   python scripts/shard_timing.py c3 256 2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from model3d_b200 import _native as N, distributed as D

wl, spp, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
ctx = N.default_context(0)
W = H = 1024
spec, psc, tr = bench.cornell_tracer(spp, wl, ctx=ctx)
if wl == "c4":
    W, H = spec["size"]
acc = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def timed(label, **kw):
    for _ in range(2):
        acc.zero_()
        tr.RenderSumsDevice(W, H, psc, acc.data_ptr(), stream=stream, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        acc.zero_()
        tr.RenderSumsDevice(W, H, psc, acc.data_ptr(), stream=stream, **kw)
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %8.2f ms per call" % (label, e0.elapsed_time(e1) / 3), flush=True)


timed("whole frame, %d spp" % spp)
part, my = D.sample_shard(spp, 0, world)
timed("shard 0 of %d (%d spp), plain flush" % (world, my), partition=part, sample_count=my)
timed("shard 0 of %d (%d spp), PART_ATOMIC flush" % (world, my), partition=part + (N.PART_ATOMIC,), sample_count=my)
part, my = D.sample_shard(spp, world - 1, world)
timed("shard %d of %d (%d spp), PART_ATOMIC flush" % (world - 1, world, my), partition=part + (N.PART_ATOMIC,), sample_count=my)
