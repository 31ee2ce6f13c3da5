"""Upper bound of what a ray-reordering pass can buy on the C2 batch: the same 2^24 rays are
permuted on the HOST by several candidate keys and the unchanged trace kernel is timed on each
order (device-resident buffers, kernel-only).  Decides whether a device binning pass is worth
building.
  python scripts/sorted_rays_experiment.py [steps]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import MeshCollider, _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
tris = bench.make_mesh()
col = MeshCollider(tris, ctx=N.Context(0))
n = 1 << 24
dev = torch.device("cuda", 0)
org, d = bench.make_rays(n, bench.SEED)


def part1by2(x):
    x = x.astype(np.uint32) & 0x3ff
    x = (x | (x << 16)) & 0x030000FF
    x = (x | (x << 8)) & 0x0300F00F
    x = (x | (x << 4)) & 0x030C30C3
    x = (x | (x << 2)) & 0x09249249
    return x


def morton(p, lo, hi, bits):
    q = np.clip((p - lo) / (hi - lo), 0, 0.999999) * (1 << bits)
    q = q.astype(np.uint32)
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


octant = ((d[:, 0] < 0).astype(np.uint32) << 2) | ((d[:, 1] < 0).astype(np.uint32) << 1) | (d[:, 2] < 0).astype(np.uint32)
# entry point of the ray into the mesh bounds [-1,1]^3 (clamped t >= 0); misses keep their origin
inv = 1.0 / np.where(np.abs(d) > 1e-20, d, 1e-20)
t0 = (-1.0 - org) * inv
t1 = (1.0 - org) * inv
tn = np.minimum(t0, t1).max(axis=1)
tf = np.maximum(t0, t1).min(axis=1)
tn = np.maximum(tn, 0)
enter = org + d * tn[:, None]
boxhit = tf >= tn
# direction quantised on an octahedral-ish grid: 3 bits per axis
keys = {
    "original": None,
    "octant": octant,
    "octant+morton(org,3b)": (octant << 9) | morton(org, -4, 4, 3),
    "octant+morton(org,4b)": (octant << 12) | morton(org, -4, 4, 4),
    "boxmiss|octant|morton(entry,3b)": ((~boxhit).astype(np.uint32) << 12) | (octant << 9) | morton(enter, -1.001, 1.001, 3),
    "boxmiss|octant|morton(entry,4b)": ((~boxhit).astype(np.uint32) << 15) | (octant << 12) | morton(enter, -1.001, 1.001, 4),
    "boxmiss|morton(entry,4b)|dir(3b)": ((~boxhit).astype(np.uint32) << 21) | (morton(enter, -1.001, 1.001, 4) << 9) | morton(d, -1.001, 1.001, 3),
    "boxmiss|morton(entry,5b)|octant": ((~boxhit).astype(np.uint32) << 18) | (morton(enter, -1.001, 1.001, 5) << 3) | octant,
}
o4 = torch.zeros((n, 4), dtype=torch.float32)
d4 = torch.full((n, 4), float("inf"), dtype=torch.float32)
h0 = torch.empty((n, 4), dtype=torch.float32, device=dev)
h1 = torch.empty((n, 4), dtype=torch.float32, device=dev)
ts = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
torch.cuda.set_stream(ts)
print("box-hit fraction %.3f" % boxhit.mean(), flush=True)
for name, key in keys.items():
    perm = np.arange(n) if key is None else np.argsort(key, kind="stable")
    o4[:, :3] = torch.from_numpy(org[perm])
    d4[:, :3] = torch.from_numpy(d[perm])
    od, dd = o4.to(dev), d4.to(dev)
    for flags, label in ((0, "trace+finish"), (N.TRACE_NO_REFINE, "no-refine")):
        for _ in range(3):
            col.FirstRayCollisionsDevice(od.data_ptr(), dd.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=ts.cuda_stream, refine=(flags == 0))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            col.FirstRayCollisionsDevice(od.data_ptr(), dd.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=ts.cuda_stream, refine=(flags == 0))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        nbins = 1 if key is None else len(np.unique(key))
        print("%-36s %-12s bins %7d  %.3f ms  %.2f Grays/s  hits %d" % (
            name, label, nbins, ms, n / ms / 1e6, int((h0[:, 3].contiguous().view(torch.int32) >= 0).sum().item())), flush=True)
    del od, dd
