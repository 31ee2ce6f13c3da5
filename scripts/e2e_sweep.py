"""End-to-end timing of m3d_mesh_first_ray_collisions with pinned host buffers (C2 workload):
H2D + pack + trace + finish + unpack + D2H, for the current M3D_PIPE_* environment."""
import os, sys, time, ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import MeshCollider, _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tris = bench.make_mesh()
ctx = N.Context(0)
col = MeshCollider(tris, ctx=ctx)
n = 1 << 24
org, d = bench.make_rays(n, bench.SEED)
org_p, d_p = torch.from_numpy(org).pin_memory(), torch.from_numpy(d).pin_memory()
t_p = torch.empty(n, dtype=torch.float32).pin_memory()
prim_p = torch.empty(n, dtype=torch.int32).pin_memory()
nrm_p = torch.empty((n, 3), dtype=torch.float32).pin_memory()
f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)

def step():
    N.check(N.lib().m3d_mesh_first_ray_collisions(
        col.h, C.cast(org_p.data_ptr(), f32p), C.cast(d_p.data_ptr(), f32p), C.c_int64(n),
        C.cast(t_p.data_ptr(), f32p), C.cast(prim_p.data_ptr(), i32p), C.cast(nrm_p.data_ptr(), f32p), None,
        C.c_uint32(0), None))

for _ in range(2):
    step()
t0 = time.perf_counter()
for _ in range(steps):
    step()
ctx.synchronize()
dt = (time.perf_counter() - t0) / steps
env = {k: v for k, v in os.environ.items() if k.startswith("M3D_")}
print("env %s: e2e %.3f ms/step %.2f Grays/s (H2D %.1f GB/s, D2H %.1f GB/s) prim checksum %d" % (
    env, dt * 1e3, n / dt / 1e9, n * 24 / dt / 1e9, n * 20 / dt / 1e9, int(prim_p.to(torch.int64).sum())), flush=True)
