"""GPU parity tests for the nearest-triangle queries (SphereCollision,
ColliderContains with margins, ColliderSolid inset / hollow, MeshToSDF) through the C ABI
against the float64 oracle (oracle/sdf.hpp) on the same seeded inputs.

Contract: the kernel keeps its running minimum in float64 with the reference's own
Triangle.Closest arithmetic, so |sdf| and the nearest point agree with the oracle to float32
output rounding (1e-6 relative + 1e-7 absolute); the face id may differ only between faces that
are equidistant (shared edges / vertices); the sign is the same parity test as ColliderContains."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check_sdf(oracle, tris32, pts, col):
    face, cp, sdf, nrm = col.FaceSDF(pts)
    ocol = oracle.Collider(tris32)
    rs, rp, rf = ocol.sdf(pts, threads=8)
    tol = 1e-6 * np.abs(rs) + 2e-7 * max(1.0, float(np.abs(pts).max()))
    flip = np.nonzero((sdf > 0) != (rs > 0))[0]
    assert len(flip) == 0 or np.abs(rs[flip]).max() < 1e-6, "sign flips away from the surface"
    assert (np.abs(np.abs(sdf) - np.abs(rs)) <= tol).all(), np.abs(np.abs(sdf) - np.abs(rs)).max()
    # nearest point: same distance; same point unless two faces tie
    d_cp = np.linalg.norm(cp.astype(np.float64) - pts.astype(np.float64), axis=1)
    assert (np.abs(d_cp - np.abs(rs)) <= 4 * tol).all()
    same = face == rf
    assert np.abs(cp[same] - rp[same]).max(initial=0) < 1e-6 * max(1.0, float(np.abs(pts).max()))
    # a different face must be equidistant in float64
    t64 = tris32.astype(np.float64)
    for i in np.nonzero(~same)[0][:200]:
        d = np.linalg.norm(oracle.triangle_closest(t64[face[i]], pts[i].astype(np.float64)) - pts[i])
        assert abs(d - abs(rs[i])) <= 1e-12 * max(1.0, abs(rs[i])), "face %d is not a tie" % i
    # NormalSDF: flat normal of the reported face
    n_ref = np.cross(t64[face, 1] - t64[face, 0], t64[face, 2] - t64[face, 0])
    n_ref /= np.linalg.norm(n_ref, axis=1, keepdims=True)
    assert np.abs(nrm - n_ref).max() < 1e-6
    return sdf, rs, same


@pytest.mark.parametrize("mesh", ["polar10", "ico32", "rect", "torus"])
def test_mesh_sdf_matches_oracle(built, oracle, mesh):
    from model3d_b200 import MeshCollider
    from test_oracle_sdf import torus_mesh
    tris = {"polar10": lambda: oracle.mesh_polar(0.5, 0.1, 10),
            "ico32": lambda: oracle.mesh_icosphere((0.1, -0.2, 0.3), 1, 32),
            "rect": lambda: oracle.mesh_rect((-1, -2, -3), (1, 2, 3)),
            "torus": lambda: torus_mesh(0.04)}[mesh]().astype(np.float32)
    rng = np.random.default_rng(77)
    pts = np.concatenate([rng.normal(size=(20000, 3)), rng.normal(size=(2000, 3)) * 10,
                          tris.reshape(-1, 3)[:3000] + rng.normal(size=(min(3000, tris.shape[0] * 3), 3)) * 1e-3])
    pts = pts.astype(np.float32)
    col = MeshCollider(tris)
    sdf, rs, same = check_sdf(oracle, tris, pts, col)
    assert (sdf > 0).sum() > 100 and (sdf < 0).sum() > 100
    assert same.mean() > 0.3  # the rest are exact ties on shared edges / vertices (checked above)


def test_sdf_vertices_and_sphere_values(built, oracle):
    """TestMeshSDFVertices (sdf_test.go:27-37) at C1 size: the marching-cubes sphere."""
    from model3d_b200 import MeshToSDF, meshes
    tris = meshes.MarchingCubesSearch(meshes.SphereSolid((0, 0, 0), 1.0), 0.01, 8).astype(np.float32)
    sdf = MeshToSDF(tris)
    verts = np.unique(tris.reshape(-1, 3), axis=0)
    assert np.abs(sdf.SDF(verts)).max() < 1e-7
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(500000, 3)).astype(np.float32)
    cp, val = sdf.PointSDF(pts)
    r = np.linalg.norm(pts.astype(np.float64), axis=1)
    assert np.abs(val - (1 - r)).max() < 2e-4  # chordal error of 0.01 cells
    assert np.abs(np.linalg.norm(cp.astype(np.float64) - pts, axis=1) - np.abs(val)).max() < 1e-6 * max(1, r.max())
    face, cp2, val2 = sdf.FaceSDF(pts[:1000])
    assert np.array_equal(val2, val[:1000]) and face.min() >= 0 and face.max() < tris.shape[0]


def test_sphere_collisions_and_contains_margins(built, oracle):
    from model3d_b200 import (ColliderContains, MeshCollider, NewColliderSolid, NewColliderSolidHollow,
                              NewColliderSolidInset)
    tris = oracle.mesh_polar(0.5, 0.1, 24).astype(np.float32)
    col, ocol = MeshCollider(tris), oracle.Collider(tris)
    rng = np.random.default_rng(9)
    pts = (rng.normal(size=(50000, 3)) * 0.5).astype(np.float32)
    radii = rng.uniform(0.01, 0.6, size=pts.shape[0]).astype(np.float32)
    got = col.SphereCollisions(pts, radii)
    ref = ocol.sphere_collisions(pts, radii.astype(np.float64), threads=8)
    dist = np.abs(ocol.sdf(pts, threads=8)[0])
    bad = got != ref
    assert (np.abs(dist[bad] - radii[bad]) < 1e-6).all()  # only exact-boundary cases may differ
    assert 0.2 < got.mean() < 0.8
    assert col.SphereCollision((0.5, 0, 0), 0.2) == bool(ocol.sphere_collisions([(0.5, 0, 0)], 0.2)[0])
    assert not col.SphereCollision((5, 0, 0), 0.0) and not col.SphereCollision((5, 0, 0), 1.0)
    for margin in (0.05, -0.05, 0.2, -0.3):
        g = ColliderContains(col, pts, margin)
        r = ocol.contains_margin(pts, margin, threads=8)
        bad = g != r
        assert (np.abs(dist[bad] - abs(margin)) < 1e-6).all(), margin
        assert g.any() and not g.all()
    assert np.array_equal(NewColliderSolid(col).Contains(pts), ocol.contains_margin(pts, 0.0, solid=1, threads=8))
    g = NewColliderSolidInset(col, 0.05).Contains(pts)
    r = ocol.contains_margin(pts, 0.05, solid=2, threads=8)
    assert (np.abs(dist[g != r] - 0.05) < 1e-6).all()
    g = NewColliderSolidHollow(col, 0.07).Contains(pts)
    r = ocol.sphere_collisions(pts, 0.07, threads=8)  # every sample point is inside the padded bounds or far
    mn, mx = tris.reshape(-1, 3).min(0) - 0.07, tris.reshape(-1, 3).max(0) + 0.07
    r &= np.all((pts >= mn) & (pts <= mx), axis=1)
    assert (np.abs(dist[g != r] - 0.07) < 1e-6).all()


def test_sdf_edge_cases(built, oracle):
    from model3d_b200 import MeshCollider, MeshToSDF
    from model3d_b200 import _native as N
    with pytest.raises((ValueError, N.M3DError)):
        MeshToSDF(np.zeros((0, 3, 3), np.float32))
    one = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    col = MeshCollider(one)
    face, cp, sdf, nrm = col.FaceSDF([[0.25, 0.25, 2.0], [2, 0, 0], [-1, -1, 0], [0.2, 0.2, 0.0]])
    assert np.allclose(np.abs(sdf), [2.0, 1.0, np.sqrt(2), 0.0], atol=1e-7)
    assert np.allclose(cp, [[0.25, 0.25, 0], [1, 0, 0], [0, 0, 0], [0.2, 0.2, 0]], atol=1e-7)
    assert (face == 0).all() and np.allclose(nrm, [[0, 0, 1]] * 4)
    assert (sdf <= 0).all()  # an open surface has no inside
    # empty batch and empty collider
    f, c, s, n = col.FaceSDF(np.zeros((0, 3), np.float32))
    assert f.shape == (0,) and s.shape == (0,)
    empty = MeshCollider(np.zeros((0, 3, 3), np.float32))
    assert not empty.SphereCollisions([[0, 0, 0]], [10.0]).any()
    assert not empty.Contains([[0, 0, 0]], margin=-1.0).any()
