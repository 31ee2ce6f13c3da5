"""Kernel time of one render with the plain flush vs the shared-accumulator flush
(M3D_PART_ATOMIC) on ONE GPU: isolates what the fused reduce costs inside the render call.
  python scripts/flush_mode_timing.py [c3|c4|c5] [spp]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import _native as N

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
spec, psc, tr = bench.cornell_tracer(spp, wl)
W, H = spec["size"] if wl == "c4" else (1024, 1024)
acc = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
for flags in (0, N.PART_ATOMIC, 0, N.PART_ATOMIC):
    import time
    for rep in range(3):
        acc.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = tr.RenderSumsDevice(W, H, psc, acc.data_ptr(), partition=(0, 0, 0, flags), sample_count=spp)
        wall = (time.perf_counter() - t0) * 1e3
    print("%s spp %d flags %d: kernel_ms %.3f wall_ms %.3f launches %d sum %.6e" % (wl, spp, flags, st["kernel_ms"], wall, st["launches"], float(acc.sum().item())), flush=True)
