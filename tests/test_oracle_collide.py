"""CPU tests that pin the oracle with the reference's own property tests, restated:
  TestMeshRayCollisions            model3d/collisions_test.go:22-76
  TestMeshRayCollisionsConsistency model3d/collisions_test.go:78-106
  TestEmptyColliders               model3d/collisions_test.go:13-20
  testSolidColliderSDFRay (sphere/rect/cylinder surface + first-is-earliest)
                                   model3d/shapes_test.go:531-603
The reference has no golden vectors for this path."""
import numpy as np
import pytest


def rand_rays(rng, n):
    org = rng.normal(size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return org, d


def test_mesh_ray_collisions_bvh_equals_brute_force(oracle):
    mesh = oracle.mesh_polar(0.5, 0.1, 10)
    col = oracle.Collider(mesh)
    rng = np.random.default_rng(1)
    org, d = rand_rays(rng, 1000)
    for i in range(1000):
        n1, t1, p1 = col.all_hits(org[i], d[i])
        n2, t2, p2 = col.all_hits(org[i], d[i], brute=True)
        assert n1 == n2
        assert np.array_equal(t1, t2)
        assert sorted(p1.tolist()) == sorted(p2.tolist())


def test_barycentric_reconstruction(oracle):
    mesh = oracle.mesh_polar(0.5, 0.1, 10)
    col = oracle.Collider(mesh)
    rng = np.random.default_rng(2)
    org, d = rand_rays(rng, 2000)
    r = col.first_hits(org, d)
    hit = r["prim"] >= 0
    assert hit.sum() > 50
    tri = mesh[r["prim"][hit]]
    p_bary = (tri * r["bary"][hit][:, :, None]).sum(1)
    p_ray = org[hit] + d[hit] * r["t"][hit][:, None]
    assert np.abs(p_bary - p_ray).max() < 1e-8


def test_first_collision_is_min_of_all(oracle):
    mesh = oracle.mesh_polar(0.5, 0.1, 100)
    assert mesh.shape[0] == 19800
    col = oracle.Collider(mesh)
    rng = np.random.default_rng(3)
    org, d = rand_rays(rng, 1000)
    r = col.first_hits(org, d)
    for i in range(1000):
        n, t, _ = col.all_hits(org[i], d[i])
        assert (n > 0) == (r["prim"][i] >= 0)
        if n:
            assert abs(r["t"][i] - t[0]) <= 1e-8


def test_empty_collider(oracle):
    col = oracle.Collider(np.zeros((0, 3, 3), np.float32))
    r = col.first_hits(np.zeros((4, 3)), np.ones((4, 3)))
    assert (r["prim"] == -1).all()
    mn, mx = col.bounds()
    assert (mn == 0).all() and (mx == 0).all()


def test_triangle_edge_rules(oracle):
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    # hit at t == 0 is accepted (primitives.go:183)
    ok, t, n, b = oracle.triangle_first_hit(tri, (0.25, 0.25, 0.0), (0, 0, 1))
    assert ok and t == 0.0
    # behind the origin -> miss
    ok, *_ = oracle.triangle_first_hit(tri, (0.25, 0.25, 1.0), (0, 0, 1))
    assert not ok
    # normal is never flipped toward the ray (primitives.go:27-33)
    ok, t, n, b = oracle.triangle_first_hit(tri, (0.25, 0.25, 1.0), (0, 0, -1))
    assert ok and np.allclose(n, (0, 0, 1)) and t == 1.0
    ok, t, n, b = oracle.triangle_first_hit(tri, (0.25, 0.25, -1.0), (0, 0, 2))
    assert ok and np.allclose(n, (0, 0, 1)) and t == 0.5  # t in units of |d|
    # inclusive barycentrics: on an edge / vertex counts (primitives.go:232,238)
    ok, *_ = oracle.triangle_first_hit(tri, (0.5, 0.0, 1.0), (0, 0, -1))
    assert ok
    ok, *_ = oracle.triangle_first_hit(tri, (0.0, 0.0, 1.0), (0, 0, -1))
    assert ok
    # parallel ray -> miss (primitives.go:208)
    ok, *_ = oracle.triangle_first_hit(tri, (0.25, 0.25, 0.0), (1, 0, 0))
    assert not ok


def test_icosphere_counts_and_radius(oracle):
    m = oracle.mesh_icosphere((1, 2, 3), 2.0, 5)
    assert m.shape == (20 * 25, 3, 3)
    r = np.linalg.norm(m.reshape(-1, 3) - np.array([1, 2, 3]), axis=1)
    assert np.abs(r - 2.0).max() < 1e-12
    # closed surface: every ray from the centre hits exactly once
    col = oracle.Collider(m)
    rng = np.random.default_rng(5)
    _, d = rand_rays(rng, 200)
    for i in range(200):
        n, _, _ = col.all_hits((1, 2, 3), d[i])
        assert n >= 1  # ==1 except exactly through shared edges


def sdf_sphere(p, c, r):
    return r - np.linalg.norm(p - c)


@pytest.mark.parametrize("kind", ["sphere", "rect", "cylinder"])
def test_shape_first_hit_on_surface(oracle, kind):
    rng = np.random.default_rng(11)
    org, d = rand_rays(rng, 3000)
    org *= 2.0
    nhit = 0
    for i in range(3000):
        if kind == "sphere":
            c, r = np.array([0.3, -0.2, 0.1]), 1.1
            ok, t, n = oracle.shape_first_hit(oracle.SPHERE, c, c, r, org[i], d[i])
            if ok:
                p = org[i] + d[i] * t
                assert abs(sdf_sphere(p, c, r)) < 1e-8
                assert np.allclose(n, (p - c) / np.linalg.norm(p - c), atol=1e-9)
        elif kind == "rect":
            mn, mx = np.array([-1, -0.5, -0.25]), np.array([0.7, 0.9, 1.3])
            ok, t, n = oracle.shape_first_hit(oracle.RECT, mn, mx, 0, org[i], d[i])
            if ok:
                p = org[i] + d[i] * t
                dist = np.minimum(np.abs(p - mn), np.abs(p - mx)).min()
                assert dist < 1e-8
                assert (p >= mn - 1e-8).all() and (p <= mx + 1e-8).all()
                assert np.abs(n).sum() == 1.0
        else:
            p1, p2, r = np.array([0, 0, -1.0]), np.array([0.5, 0.2, 1.0]), 0.6
            ok, t, n = oracle.shape_first_hit(oracle.CYLINDER, p1, p2, r, org[i], d[i])
            if ok:
                p = org[i] + d[i] * t
                v = (p2 - p1) / np.linalg.norm(p2 - p1)
                frac = (p - p1) @ v
                radial = np.linalg.norm((p - p1) - v * frac)
                on_side = abs(radial - r) < 1e-8 and -1e-8 <= frac <= np.linalg.norm(p2 - p1) + 1e-8
                on_cap = (abs(frac) < 1e-8 or abs(frac - np.linalg.norm(p2 - p1)) < 1e-8) and radial <= r + 1e-8
                assert on_side or on_cap
        assert t >= 0 or not ok
        nhit += ok
    assert nhit > 100


def test_sphere_inside_hit_and_tangent(oracle):
    # from inside: second root, normal still outward (shapes.go:52-93)
    ok, t, n = oracle.shape_first_hit(oracle.SPHERE, (0, 0, 0), (0, 0, 0), 1.0, (0, 0, 0), (0, 0, 2))
    assert ok and t == 0.5 and np.allclose(n, (0, 0, 1))
    # tangent ray misses (discriminant <= 0)
    ok, *_ = oracle.shape_first_hit(oracle.SPHERE, (0, 0, 0), (0, 0, 0), 1.0, (1, 0, -5), (0, 0, 1))
    assert not ok


def test_scene_cast_closest_wins(oracle):
    sc = oracle.Scene()
    m = oracle.MaterialDesc()
    mi = sc.add_material(m)
    sc.add_sphere((0, 0, 5), 1.0, mi)
    sc.add_mesh(oracle.mesh_rect((-1, -1, 1), (1, 1, 2)).astype(np.float32), mi)
    sc.add_rect((-1, -1, 8), (1, 1, 9), mi)
    r = sc.cast(np.array([[0, 0, -3.0], [0, 0, 20.0], [5, 5, 5]]), np.array([[0, 0, 1.0], [0, 0, -1.0], [1, 0, 0]]))
    assert r["obj"].tolist() == [1, 2, -1]
    assert np.allclose(r["t"][:2], [4.0, 11.0])


def test_all_hits_batch_matches_single_ray_walk(oracle):
    """The batched all-hits helper the GPU RayCollisions test compares with == the per-ray
    BVH walk == brute force over every triangle (collisions_test.go:22-76)."""
    rng = np.random.default_rng(5)
    tris = (rng.normal(size=(400, 1, 3)) + rng.normal(size=(400, 3, 3)) * 0.4).astype(np.float32)
    col = oracle.Collider(tris)
    org = rng.normal(size=(300, 3)).astype(np.float32)
    d = rng.normal(size=(300, 3)).astype(np.float32)
    b = col.all_hits_batch(org, d, threads=2)
    assert b["offsets"][-1] > 300
    for i in range(300):
        n, t, prim = col.all_hits(org[i].astype(np.float64), d[i].astype(np.float64), brute=True)
        a0, a1 = b["offsets"][i], b["offsets"][i + 1]
        assert n == a1 - a0
        assert np.array_equal(np.sort(prim), np.sort(b["prim"][a0:a1]))
        assert np.allclose(np.sort(t), b["t"][a0:a1], rtol=0, atol=1e-12)
        p = org[i] + d[i] * b["t"][a0:a1, None]
        q = np.einsum("nk,nkc->nc", b["bary"][a0:a1], tris[b["prim"][a0:a1]].astype(np.float64))
        assert np.abs(p - q).max() < 1e-8 if n else True
