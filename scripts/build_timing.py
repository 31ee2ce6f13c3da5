"""BVH build times of the three builders on the C2 mesh (1,003,520 triangles) and the trace rate
on each: python scripts/build_timing.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import MeshCollider

tris = bench.make_mesh()
org, d = bench.make_rays(1 << 22, 5)
for name, kw in (("host SAH", {}), ("device LBVH + host collapse", dict(device_lbvh=True)), ("full device build", dict(device_build=True))):
    MeshCollider(tris[:1000], **kw)
    t0 = time.perf_counter()
    col = MeshCollider(tris, **kw)
    wall = (time.perf_counter() - t0) * 1e3
    info = col.Info()
    r = col.FirstRayCollisions(org, d, counters=True)
    r2 = col.FirstRayCollisions(org, d, want_stats=True)
    print("%s: build %.1f ms (wall incl. upload %.1f ms), %d nodes, depth %d, sah %.2f; nodes/ray %.2f tris/ray %.2f, kernels %.3f ms; hits checksum %d" % (
        name, info["build_ms"], wall, info["num_nodes"], info["max_depth"], info["sah_cost"], r.Stats["nodes_visited"] / org.shape[0],
        r.Stats["tris_tested"] / org.shape[0], r2.Stats["kernel_ms"], int(r.Triangle.astype(np.int64).sum())), flush=True)
for rep in range(3):
    t0 = time.perf_counter()
    col = MeshCollider(tris, device_build=True)
    print("full device build, repeat %d: wall %.1f ms, build %.1f ms" % (rep, (time.perf_counter() - t0) * 1e3, col.Info()["build_ms"]), flush=True)
import ctypes as C
from model3d_b200 import _native as N
t9 = np.ascontiguousarray(tris.reshape(-1, 9))
for flags in (2, 2, 1):
    h = C.c_void_p()
    t0 = time.perf_counter()
    N.check(N.lib().m3d_mesh_create(N.default_context().h, t9.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(t9.shape[0]), None, C.c_uint32(flags), C.byref(h)))
    print("m3d_mesh_create flags %d: %.1f ms" % (flags, (time.perf_counter() - t0) * 1e3), flush=True)
    N.lib().m3d_mesh_destroy(h)
