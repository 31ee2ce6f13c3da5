"""Times the nearest-triangle (MeshToSDF) kernel: C1 marching-cubes sphere (376,832 triangles) and
the C2 icosphere, random points ~N(0, I) and a regular grid (the marching-cubes access pattern)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from model3d_b200 import MeshCollider, meshes

def run(name, tris, pts):
    col = MeshCollider(tris)
    col.FaceSDF(pts[:1000])
    best = 1e9
    for _ in range(3):
        st = col.FaceSDF(pts, want_stats=True)[-1]
        best = min(best, st["kernel_ms"])
    n = pts.shape[0]
    print("%s: %d tris, %d points, kernel %.3f ms -> %.1f Mqueries/s; per query: nodes %.1f, f32 screens %.1f, f64 evals %.2f" % (
        name, tris.shape[0], n, best, n / best / 1e3, st["nodes_visited"] / n, st["tris_tested"] / n, st["hits"] / n), flush=True)

rng = np.random.default_rng(1)
n = 1 << 22
rand = rng.normal(size=(n, 3)).astype(np.float32)
g = np.linspace(-1.1, 1.1, 162, dtype=np.float32)
grid = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
u = rng.normal(size=(n, 3))
u /= np.linalg.norm(u, axis=1, keepdims=True)
shell = (u * (1.0 + 0.03 * rng.normal(size=(n, 1)))).astype(np.float32)
c1 = meshes.MarchingCubesSearch(meshes.SphereSolid((0, 0, 0), 1.0), 0.01, 8).astype(np.float32)
run("C1 mesh / random", c1, rand)
run("C1 mesh / grid", c1, grid)
run("C1 mesh / shell (|r-1| ~ 0.03)", c1, shell)
c2 = meshes.NewMeshIcosphere((0, 0, 0), 1.0, 224).astype(np.float32)
run("C2 mesh / random", c2, rand)
run("C2 mesh / shell (|r-1| ~ 0.03)", c2, shell)
