"""Scale check: a 5-million-triangle icosphere through the full device build and a 2^24-ray batch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from model3d_b200 import MeshCollider, meshes
n_sub = int(sys.argv[1]) if len(sys.argv) > 1 else 500
t0 = time.perf_counter()
tris = meshes.NewMeshIcosphere((0, 0, 0), 1.0, n_sub).astype(np.float32)
print("mesh: %d triangles (%.1f s on the host)" % (tris.shape[0], time.perf_counter() - t0), flush=True)
org, d = bench.make_rays(1 << 24, 9)
for name, kw in (("full device build", dict(device_build=True)), ("host SAH", {})):
    t0 = time.perf_counter()
    col = MeshCollider(tris, **kw)
    wall = time.perf_counter() - t0
    info = col.Info()
    r = col.FirstRayCollisions(org, d, counters=True)
    r2 = col.FirstRayCollisions(org, d, want_stats=True)
    print("%s: wall %.2f s, build %.1f ms, %d nodes, %.0f MB on device; nodes/ray %.2f tris/ray %.2f; kernels %.2f ms -> %.2f Grays/s; hits %.4f checksum %d" % (
        name, wall, info["build_ms"], info["num_nodes"], info["device_bytes"] / 1e6, r.Stats["nodes_visited"] / org.shape[0],
        r.Stats["tris_tested"] / org.shape[0], r2.Stats["kernel_ms"], org.shape[0] / r2.Stats["kernel_ms"] / 1e6,
        r.Collides.mean(), int(r.Triangle.astype(np.int64).sum())), flush=True)
    del col
