"""Marching cubes (BASELINE config C1's mesh: Sphere -> MarchingCubesSearch(0.01, 8)).
The oracle restates model3d/mc.go in C++ (oracle/mcubes.hpp); the product-side generator
(model3d_b200/meshes.py) is a vectorised numpy version and must be bit-identical.  The
reference's own tests for this code are properties (mc_test.go:9-42: table determinism,
MustValidateMesh = closed oriented manifold), restated here."""
import numpy as np

from model3d_b200 import meshes


def _directed_edges(tris):
    v, inv = np.unique(tris.reshape(-1, 3), axis=0, return_inverse=True)
    f = inv.reshape(-1, 3)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    return v, f, e


def assert_closed_oriented_manifold(tris):
    """MustValidateMesh(t, mesh, true) (model3d/testing.go): every edge is shared by exactly two
    triangles with opposite orientation, no repeated triangles, no degenerate ones."""
    v, f, e = _directed_edges(tris)
    assert (f[:, 0] != f[:, 1]).all() and (f[:, 1] != f[:, 2]).all() and (f[:, 0] != f[:, 2]).all()
    key = e[:, 0].astype(np.int64) * len(v) + e[:, 1]
    rkey = e[:, 1].astype(np.int64) * len(v) + e[:, 0]
    assert len(np.unique(key)) == len(key), "a directed edge is used twice"
    assert np.array_equal(np.sort(key), np.sort(rkey)), "an edge has no opposite partner"
    # Euler characteristic of a sphere-like closed surface
    return len(v) - len(key) // 2 + len(f)


def test_lookup_table_matches_oracle_and_is_complete(oracle):
    c, t = oracle.mc_table()
    c2, t2 = meshes._mc_lookup_table()
    assert np.array_equal(c, c2) and np.array_equal(t, t2)
    assert c[0] == 0 and c[255] == 0 and (c[1:255] > 0).all() and c.max() == 5
    # complementary masks have the same number of boundary edges crossed on unambiguous cases
    for bit in range(8):
        assert c[1 << bit] == 1 and c[255 ^ (1 << bit)] == 1


def test_sphere_mesh_bit_identical_to_oracle(oracle):
    for cen, r, delta, iters in [((0, 0, 0), 1.0, 0.05, 0), ((0.1, 0.3, -0.2), 1.0, 0.05, 8),
                                 ((0.1, -0.2, 0.3), 0.7, 0.03, 3)]:
        a = meshes.MarchingCubesSearch(meshes.SphereSolid(cen, r), delta, iters)
        b = oracle.mesh_mc_sphere(cen, r, delta, iters)
        assert a.shape == b.shape and np.array_equal(a, b)


def test_c1_mesh_size_and_surface(oracle):
    m = oracle.mesh_mc_sphere((0, 0, 0), 1.0, 0.01, 8)
    assert m.shape[0] == 376832  # C1 triangle count
    r = np.linalg.norm(m.reshape(-1, 3), axis=1)
    # 8 bisection steps of a 0.01 edge: |r - 1| <= 0.01 / 2^9
    assert np.abs(r - 1).max() <= 0.01 / 512 + 1e-12
    assert np.array_equal(m, meshes.MarchingCubesSearch(meshes.SphereSolid((0, 0, 0), 1.0), 0.01, 8))


def test_sphere_mesh_is_closed_oriented_manifold(oracle):
    m = oracle.mesh_mc_sphere((0.1, 0.3, -0.2), 1.0, 0.05, 8)
    assert assert_closed_oriented_manifold(m) == 2
    # outward orientation: normal . (centroid - centre) > 0 (mcTriangle is counter-clockwise from outside)
    n = np.cross(m[:, 1] - m[:, 0], m[:, 2] - m[:, 0])
    assert (np.einsum("ij,ij->i", n, m.mean(axis=1) - np.array([0.1, 0.3, -0.2])) > 0).all()


class _Blobs:
    """A random union of balls (stands in for mc_test.go's randomSolid)."""

    def __init__(self, rng, k=6):
        self.c = rng.uniform(-0.6, 0.6, size=(k, 3))
        self.r = rng.uniform(0.2, 0.5, size=k)

    def Min(self):
        return (self.c - self.r[:, None]).min(0)

    def Max(self):
        return (self.c + self.r[:, None]).max(0)

    def Contains(self, pts):
        d = np.linalg.norm(pts[:, None, :] - self.c[None], axis=2)
        return (d <= self.r[None]).any(axis=1)


def test_random_solids_give_manifolds():
    rng = np.random.default_rng(1337)
    for i in range(6):
        solid = _Blobs(rng)
        for iters in (0, 2):
            m = meshes.MarchingCubesSearch(solid, 0.1, iters)
            assert m.shape[0] > 100
            assert_closed_oriented_manifold(m)
