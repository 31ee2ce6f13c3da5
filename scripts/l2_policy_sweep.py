"""C2 first-hit batch (device-resident SoA buffers) under different L2 residency settings of the
traversal kernel, all in one process (the M3D_L2_* knobs are read per launch):
  python scripts/l2_policy_sweep.py [steps]
M3D_LIB=variants/libm3dgpu_l2n.so etc. selects a build whose node / triangle loads carry an
evict_last cache policy (M3D_L2_HINT); with such a library only the first rows are of interest."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import MeshCollider, _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
tris = bench.make_mesh()
col = MeshCollider(tris, ctx=N.Context(0))
n = 1 << 24
dev = torch.device("cuda", 0)
org, d = bench.make_rays(n, bench.SEED)
o4 = torch.zeros((n, 4), dtype=torch.float32)
d4 = torch.full((n, 4), float("inf"), dtype=torch.float32)
o4[:, :3] = torch.from_numpy(org)
d4[:, :3] = torch.from_numpy(d)
o4, d4 = o4.to(dev), d4.to(dev)
h0 = torch.empty((n, 4), dtype=torch.float32, device=dev)
h1 = torch.empty((n, 4), dtype=torch.float32, device=dev)
ts = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
torch.cuda.set_stream(ts)


def run():
    col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=ts.cuda_stream)


run()
torch.cuda.synchronize()
ref = h0[:, 3].clone()

configs = [{}]
if not quick:
    for mb in (32, 64, 96, 1024):
        for what in ("nodes", "tris"):
            for ratio in ("1.0", "0.6"):
                configs.append({"M3D_L2_PERSIST_MB": str(mb), "M3D_L2_WINDOW": what, "M3D_L2_HITRATIO": ratio})
configs.append({})
for cfg in configs:
    for k in ("M3D_L2_PERSIST_MB", "M3D_L2_WINDOW", "M3D_L2_HITRATIO"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    for _ in range(4):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    same = bool(torch.equal(ref.view(torch.int32), h0[:, 3].contiguous().view(torch.int32)))
    print("lib %s cfg %s: %.3f ms/step %.2f Grays/s same_hits %s" % (
        os.environ.get("M3D_LIB", "default"), cfg, ms, n / ms / 1e6, same), flush=True)
