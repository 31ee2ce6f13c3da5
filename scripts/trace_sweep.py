"""Kernel-only timing of the C2 first-hit batch (device-resident SoA buffers) for tuning runs:
prints ms/step, Grays/s and nodes / triangle tests per ray for the current M3D_* environment.
  python scripts/trace_sweep.py [steps] [mix]     mix: A (incoherent, bench default) | B (camera)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import MeshCollider, _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
mix = sys.argv[2] if len(sys.argv) > 2 else "A"
tris = bench.make_mesh()
col = MeshCollider(tris, ctx=N.Context(0))
n = 1 << 24
dev = torch.device("cuda", 0)
if mix == "A":
    org, d = bench.make_rays(n, bench.SEED)
else:
    from model3d_b200 import render3d as R
    cam = R.NewCameraAt((0.0, -3.0, 0.0), (0.0, 0.0, 0.0), np.pi / 3.6)
    d = R.CasterRays(cam, 4096, 4096).astype(np.float32)
    org = np.tile(np.array([[0.0, -3.0, 0.0]], np.float32), (n, 1))
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
st = col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=ts.cuda_stream, counters=True)
torch.cuda.synchronize()
ref_prim = h0[:, 3].clone()
for _ in range(5):
    col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=ts.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    col.FirstRayCollisionsDevice(o4.data_ptr(), d4.data_ptr(), n, h0.data_ptr(), h1.data_ptr(), stream=ts.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
same = bool(torch.equal(ref_prim.view(torch.int32), h0[:, 3].contiguous().view(torch.int32)))
env = {k: v for k, v in os.environ.items() if k.startswith("M3D_")}
# checksum over every bit of both hit records (final outputs must not depend on tuning knobs)
bits = int(h0.view(torch.int32).to(torch.int64).sum().item()) ^ (int(h1.view(torch.int32).to(torch.int64).sum().item()) << 1)
print("env %s mix %s: %.3f ms/step %.2f Grays/s nodes/ray %.3f tris/ray %.3f prim_checksum %d all_bits %d stable %s" % (
    env, mix, ms, n / ms / 1e6, st["nodes_visited"] / n, st["tris_tested"] / n,
    int(h0[:, 3].contiguous().view(torch.int32).to(torch.int64).sum().item()), bits, same), flush=True)
