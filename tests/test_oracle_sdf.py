"""CPU tests that pin the oracle's nearest-triangle / SDF restatement (oracle/sdf.hpp) with the
reference's own property tests, restated:
  TestTriangleDist               model3d/primitives_test.go:48-76 (Closest vs SphereCollision bisection)
  TestMeshShapeCollisions/Sphere model3d/collisions_test.go:115-130 (BVH == brute force)
  TestMeshSDFVertices            model3d/sdf_test.go:27-37
  TestMeshPointSDF               model3d/sdf_test.go:39-57
  TestMeshSDFConsistency         model3d/sdf_test.go:10-25 (bisection of SphereCollision == exact SDF)
"""
import numpy as np

from model3d_b200 import meshes


class TorusSolid:
    """sdfTestingSolid (sdf_test.go:189-196): TorusSolid{axis (1,2,-0.5) normalised, inner 0.2, outer 0.7}."""

    def __init__(self):
        a = np.array([1.0, 2.0, -0.5])
        self.axis = a / np.linalg.norm(a)
        self.inner, self.outer = 0.2, 0.7

    def Min(self):
        return -np.full(3, self.inner + self.outer)

    def Max(self):
        return np.full(3, self.inner + self.outer)

    def Contains(self, pts):
        h = pts @ self.axis
        perp = pts - h[:, None] * self.axis
        ring = np.linalg.norm(perp, axis=1) - self.outer
        return np.sqrt(ring * ring + h * h) <= self.inner


def torus_mesh(delta=0.05):
    return meshes.MarchingCubesSearch(TorusSolid(), delta, 8)


def test_triangle_closest_matches_sphere_collision_bisection(oracle):
    rng = np.random.default_rng(3)
    for _ in range(100):
        tri = rng.normal(size=(3, 3))
        for _ in range(10):
            c = rng.normal(size=3)
            lo, hi = 0.0, float(np.linalg.norm(tri[0] - c))
            for _ in range(64):
                mid = (lo + hi) / 2
                if oracle.triangle_sphere_collision(tri, c, mid):
                    hi = mid
                else:
                    lo = mid
            d = np.linalg.norm(oracle.triangle_closest(tri, c) - c)
            assert abs((lo + hi) / 2 - d) < 1e-5


def test_mesh_sphere_collision_equals_brute_force(oracle):
    mesh = oracle.mesh_polar(0.5, 0.1, 10)
    col = oracle.Collider(mesh)
    rng = np.random.default_rng(4)
    centers = rng.normal(size=(1000, 3)).astype(np.float32)
    radii = rng.uniform(0.1, 1.1, size=1000)
    got = col.sphere_collisions(centers, radii)
    tri32 = mesh.astype(np.float32).astype(np.float64)
    for i in range(1000):
        exp = any(oracle.triangle_sphere_collision(t, centers[i].astype(np.float64), radii[i]) for t in tri32)
        assert got[i] == exp
    assert 0.1 < got.mean() < 0.9


def test_mesh_sdf_equals_brute_force_and_point_is_consistent(oracle):
    mesh = oracle.mesh_polar(0.5, 0.1, 8).astype(np.float32)
    col = oracle.Collider(mesh)
    rng = np.random.default_rng(5)
    pts = (rng.normal(size=(300, 3)) * 0.6).astype(np.float32)
    sdf, cp, face = col.sdf(pts)
    inside = col.contains_margin(pts, 0.0, solid=1)
    assert np.array_equal(sdf > 0, inside)
    m64 = mesh.astype(np.float64)
    for i in range(300):
        d = np.array([np.linalg.norm(oracle.triangle_closest(t, pts[i].astype(np.float64)) - pts[i]) for t in m64])
        assert abs(d.min() - abs(sdf[i])) < 1e-14
        assert d[face[i]] - d.min() < 1e-14
    # TestMeshPointSDF: |closest - c| == |sdf|
    assert np.abs(np.linalg.norm(cp - pts, axis=1) - np.abs(sdf)).max() < 1e-12


def test_mesh_sdf_is_zero_on_vertices_and_matches_bisection(oracle):
    mesh = torus_mesh().astype(np.float32)
    col = oracle.Collider(mesh)
    verts = np.unique(mesh.reshape(-1, 3), axis=0)
    sdf, _, _ = col.sdf(verts, threads=8)
    assert np.abs(sdf).max() < 1e-8  # TestMeshSDFVertices
    # TestMeshSDFConsistency: ColliderToSDF (bisection over SphereCollision, sdf.go:148-185) == exact
    rng = np.random.default_rng(6)
    pts = rng.normal(size=(100, 3)).astype(np.float32)
    exact, _, _ = col.sdf(pts, threads=8)
    lo, hi = np.zeros(100), np.full(100, 8.0)
    for _ in range(48):
        mid = (lo + hi) / 2
        hit = col.sphere_collisions(pts, mid, threads=8)
        hi = np.where(hit, mid, hi)
        lo = np.where(hit, lo, mid)
    assert np.abs((lo + hi) / 2 - np.abs(exact)).max() < 1e-5


def test_collider_contains_margin_rules(oracle):
    mesh = oracle.mesh_icosphere((0, 0, 0), 1.0, 6).astype(np.float32)
    col = oracle.Collider(mesh)
    pts = np.array([[0, 0, 0], [0.93, 0, 0], [1.05, 0, 0], [3, 0, 0]], np.float32)
    assert col.contains_margin(pts, 0.0).tolist() == [True, True, False, False]
    assert col.contains_margin(pts, 0.1).tolist() == [True, False, False, False]
    assert col.contains_margin(pts, -0.1).tolist() == [True, True, True, False]
    # ColliderSolid.Contains checks the bounds first (solid.go:293-295)
    assert col.contains_margin(pts, 0.0, solid=1).tolist() == [True, True, False, False]
