"""GPU parity tests for scenes (Object.Cast batches) and the RayCaster, through the C ABI,
against the float64 oracle.  Contract: RayCaster images within 1/255 per channel; first-hit
object / triangle ids identical except ties; t and normals within 1e-5 relative."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def rays(rng, n, scale=2.0):
    o = (rng.normal(size=(n, 3)) * scale).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def test_scene_cast_matches_oracle(built, oracle):
    spec = scenes.mixed_scene()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    rng = np.random.default_rng(77)
    org, d = rays(rng, 300000)
    got = psc.Cast(org, d)
    ref = osc.cast(org, d, threads=8)
    hit_o, hit_g = ref["obj"] >= 0, got["obj"] >= 0
    flips = (hit_o != hit_g).sum()
    assert flips <= 3, flips  # tangent rays on analytic shapes only
    both = hit_o & hit_g
    same = both & (ref["obj"] == got["obj"]) & (ref["prim"] == got["prim"])
    assert same.sum() >= both.sum() - 5
    rel = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
    assert rel[same].max() < 1e-5
    assert np.abs(got["normal"][same] - ref["normal"][same]).max() < 2e-5
    # every object kind is exercised
    assert set(np.unique(got["obj"][hit_g]).tolist()) == set(range(7))
    mn, mx = psc.Min(), psc.Max()
    assert mn[2] == pytest.approx(-1.5) and mx[0] == pytest.approx(3.0)


def test_unsupported_types_raise(built):
    from model3d_b200 import render3d as R
    from model3d_b200 import UnsupportedError

    class Torus:
        pass

    class Fancy:
        pass

    with pytest.raises(UnsupportedError):
        R.Scene(R.JoinedObject([R.ColliderObject(Collider=Torus(), Material=R.LambertMaterial())]))
    with pytest.raises(UnsupportedError):
        R.Scene(R.JoinedObject([R.ColliderObject(Collider=R.Sphere(), Material=Fancy())]))
    with pytest.raises(UnsupportedError):
        R.Scene(R.JoinedObject([Fancy()]))
    # a rotated Rect is no longer axis-aligned: the C ABI reports UNSUPPORTED
    from model3d_b200 import _native as N
    with pytest.raises(N.M3DError) as ei:
        R.Scene(R.JoinedObject([R.MatrixMultiply(
            R.ColliderObject(Collider=R.Rect(), Material=R.LambertMaterial()), scenes.rotation((0, 0, 1), 0.3))]))
    assert ei.value.code == 2


@pytest.mark.parametrize("size", [(64, 48), (257, 131)])
def test_raycaster_mixed_scene_image(built, oracle, size):
    from model3d_b200 import render3d as R
    W, H = size
    spec = scenes.mixed_scene()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    src, dst, fov = (4.0, -6.0, 3.0), (0.0, 0.0, 0.0), np.pi / 3.6
    ocam = oracle.camera_at(src, dst, fov)
    pcam = R.NewCameraAt(src, dst, fov)
    assert np.allclose(pcam.ScreenX, list(ocam.screen_x)) and np.allclose(pcam.ScreenY, list(ocam.screen_y))
    ol = oracle.PointLight()
    ol.origin[:], ol.color[:], ol.quad_dropoff = (30.0, -40.0, 50.0), (1.0, 0.9, 0.8), 0
    ol2 = oracle.PointLight()
    ol2.origin[:], ol2.color[:], ol2.quad_dropoff = (-3.0, -4.0, 6.0), (40.0, 40.0, 60.0), 1
    ref = osc.render_raycast(ocam, [ol, ol2], W, H, threads=8)
    img = R.Image(W, H)
    img.Data[:] = 0.25  # miss pixels must stay untouched (raycast.go:26-28)
    rc = R.RayCaster(Camera=pcam, Lights=[R.PointLight((30.0, -40.0, 50.0), (1.0, 0.9, 0.8)),
                                          R.PointLight((-3.0, -4.0, 6.0), (40.0, 40.0, 60.0), True)])
    rc.Render(img, psc)
    miss = ref["obj"] < 0
    ref_img = ref["img"].copy()
    ref_img[miss] = 0.25
    diff = np.abs(img.Data.astype(np.float64) - ref_img)
    # pixels whose ray grazes a silhouette may pick another surface; everything else <= 1/255
    bad = (diff.max(axis=2) > 1.0 / 255).sum()
    assert bad <= 2e-3 * W * H, bad
    assert np.median(diff) < 1e-6
    # 8-bit sRGB (image.go:125-145)
    d8 = np.abs(img.RGBA8().astype(int) - oracle.srgb8(np.clip(ref_img, 0, 1)).astype(int))
    assert (d8.max(axis=2) > 1).sum() <= 2e-3 * W * H


def test_raycaster_c1_config(built, oracle):
    """BASELINE config 1: Sphere -> MarchingCubesSearch(0.01, 8) mesh (376,832 triangles, the
    product-side mesh is bit-identical to the oracle's restatement of mc.go), RayCaster 512x512,
    one frame; per-pixel id / t / image parity."""
    from model3d_b200 import render3d as R
    spec = scenes.c1_scene()
    assert spec["objects"][0]["tris"].shape[0] == 376832
    assert np.array_equal(spec["objects"][0]["tris"], oracle.mesh_mc_sphere((0, 0, 0), 1.0, 0.01, 8).astype(np.float32))
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    W = H = 512
    cam = spec["camera"]
    ocam = oracle.camera_at(cam["src"], cam["dst"], cam["fov"])
    pcam = R.NewCameraAt(cam["src"], cam["dst"], cam["fov"])
    lt = spec["lights"][0]
    ol = oracle.PointLight()
    ol.origin[:], ol.color[:], ol.quad_dropoff = lt["origin"], lt["color"], 0
    ref = osc.render_raycast(ocam, [ol], W, H, threads=8)
    img = R.Image(W, H)
    R.RayCaster(Camera=pcam, Lights=[R.PointLight(lt["origin"], lt["color"])]).Render(img, psc)
    hit = ref["obj"] >= 0
    assert 0.2 < hit.mean() < 0.8
    diff = np.abs(img.Data.astype(np.float64) - ref["img"]).max(axis=2)
    assert (diff > 1.0 / 255).sum() <= 1e-3 * W * H, (diff > 1.0 / 255).sum()
    assert diff[hit].mean() < 1e-5
    # first-hit ids and t on the exact camera rays (float32-rounded, as the GPU traces them)
    dirs = oracle.camera_rays(ocam, W, H).astype(np.float32)
    org = np.tile(np.asarray(cam["src"], np.float32), (W * H, 1))
    got = psc.Cast(org, dirs)
    ref2 = osc.cast(org, dirs, threads=8)
    assert np.array_equal(got["obj"] >= 0, ref2["obj"] >= 0)
    h2 = ref2["obj"] >= 0
    same = h2 & (got["prim"] == ref2["prim"])
    assert same.sum() >= h2.sum() - 3
    rel = np.abs(got["t"] - ref2["t"])[same] / np.abs(ref2["t"][same])
    assert rel.max() < 1e-5


def test_raycaster_row_partition(built, oracle):
    """Multi-GPU tiling: row bands rendered separately equal the whole frame."""
    from model3d_b200 import render3d as R
    spec = scenes.mixed_scene()
    psc = scenes.build_product(spec)
    W, H = 96, 80
    pcam = R.NewCameraAt((4.0, -6.0, 3.0), (0, 0, 0), np.pi / 3.6)
    rc = R.RayCaster(Camera=pcam, Lights=[R.PointLight((30.0, -40.0, 50.0), (1.0, 1.0, 1.0))])
    whole = R.Image(W, H)
    rc.Render(whole, psc)
    parts = R.Image(W, H)
    for band in [(0, 13), (13, 40), (40, 80)]:
        rc.Render(parts, psc, partition=band)
    assert np.array_equal(whole.Data, parts.Data)


def test_scene_device_build_same_casts(built, oracle):
    """m3d_scene_build with M3D_MESH_BUILD_DEVICE_COLLAPSE (object / triangle ids travel through
    the device build): the showcase scene casts identically to the host-SAH scene."""
    spec = scenes.showcase(hd=False)
    a = scenes.build_product(spec)
    b = scenes.build_product(spec, device_build=True)
    rng = np.random.default_rng(12)
    n = 200000
    org = (rng.normal(size=(n, 3)) * 2.0 + np.array([0.0, 0.0, 2.0])).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    ra, rb = a.Cast(org, d), b.Cast(org, d)
    assert (ra["obj"] != rb["obj"]).sum() <= 3
    same = (ra["obj"] == rb["obj"]) & (ra["prim"] == rb["prim"])
    assert same.sum() >= n - 6
    assert np.array_equal(ra["t"][same], rb["t"][same])
    assert np.array_equal(ra["normal"][same], rb["normal"][same])


def test_ten_thousand_spheres_object_level_bvh(built, oracle):
    """render3d.BVHToObject over many primitives (object.go:172-185; consumer
    cli/pan_pointcloud/main.go:80-95 builds one over thousands of spheres).  Scenes with more than a
    handful of analytic shapes get an object-level wide BVH; casts must equal the oracle's linear
    JoinedObject scan (object.go:141-153) on the same 10,000 spheres + a floor mesh + a rect + a
    cylinder."""
    rng = np.random.default_rng(42)
    m1 = scenes.lambert(diffuse=scenes.gray(0.5))
    m2 = scenes.phong(20.0, specular=scenes.gray(0.3), diffuse=(0.4, 0.2, 0.1))
    objs = []
    centers = rng.uniform(-10, 10, size=(10000, 3))
    radii = rng.uniform(0.02, 0.25, size=10000)
    for c, r in zip(centers, radii):
        objs.append(dict(kind="sphere", center=tuple(c.tolist()), radius=float(r), material=m1))
    objs.append(dict(kind="mesh", tris=scenes.mesh_rect_tris((-11, -11, -11.5), (11, 11, -11)).astype(np.float32), material=m2))
    objs.append(dict(kind="rect", min=(-1.0, -1.0, 10.5), max=(1.0, 1.0, 11.0), material=m2))
    objs.append(dict(kind="cylinder", p1=(10.5, 0.0, -3.0), p2=(10.8, 0.5, 3.0), radius=0.4, material=m2))
    spec = dict(objects=objs)
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    n = 60000
    org = rng.uniform(-12, 12, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    got = psc.Cast(org, d, counters=True)
    ref = osc.cast(org, d, threads=8)
    hit_o, hit_g = ref["obj"] >= 0, got["obj"] >= 0
    assert (hit_o != hit_g).sum() <= 3  # tangent rays on spheres
    both = hit_o & hit_g
    same = both & (ref["obj"] == got["obj"])
    assert same.sum() >= both.sum() - 5
    rel = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
    assert rel[same].max() < 1e-5
    assert np.abs(got["normal"][same] - ref["normal"][same]).max() < 2e-5
    assert hit_g.mean() > 0.3
    assert len(np.unique(got["obj"][hit_g])) > 3000  # thousands of different spheres are hit
    assert {10000, 10001, 10002} <= set(np.unique(got["obj"][hit_g]).tolist())
    # and the RayCaster / path tracer run on the same hierarchy
    from model3d_b200 import render3d as R
    cam = R.NewCameraAt((0.0, -30.0, 8.0), (0.0, 0.0, 0.0), np.pi / 4)
    lt = R.PointLight(Origin=(20.0, -40.0, 50.0), Color=(1.0, 1.0, 1.0))
    img = R.Image(160, 120)
    R.RayCaster(Camera=cam, Lights=[lt]).Render(img, psc)
    ocam = oracle.camera_at((0.0, -30.0, 8.0), (0.0, 0.0, 0.0), np.pi / 4)
    ol = oracle.PointLight()
    ol.origin[:], ol.color[:], ol.quad_dropoff = (20.0, -40.0, 50.0), (1.0, 1.0, 1.0), 0
    want = osc.render_raycast(ocam, [ol], 160, 120, threads=8)["img"]
    diff = np.abs(np.asarray(img.Data, np.float64) - want)
    assert (diff.max(axis=2) > 1.0 / 255).sum() <= 12  # silhouette pixels of tangent spheres
    tr = R.RecursiveRayTracer(Camera=cam, Lights=[lt], MaxDepth=2, NumSamples=8, Seed=2)
    rgb, _, st = tr.RenderSums(80, 60, psc)
    assert np.isfinite(rgb).all() and rgb.sum() > 0 and st["rays"] > 80 * 60 * 8


def test_golf_balls_instancing_one_copy_of_the_triangles(built, oracle):
    """one collider under many transforms (examples/renderings/golf_balls/main.go:25-39:
    render3d.Translate of a shared ball).  A MeshCollider shared by several objects is instanced: the
    scene stores no triangles of its own for it, rays go to object space at the instance's bounds
    (transform.go:26-31,76-85).  Casts equal the oracle's, which transforms the ray the same way."""
    from model3d_b200 import MeshCollider, meshes, render3d as R
    rng = np.random.default_rng(3)
    ball = meshes.NewMeshIcosphere((0, 0, 0), 0.5, 12).astype(np.float32)  # 2,880 triangles
    col = MeshCollider(ball)
    mat = R.PhongMaterial(Alpha=10.0, SpecularColor=(0.2, 0.2, 0.2), DiffuseColor=(0.6, 0.6, 0.6))
    m_o = scenes.phong(10.0, specular=scenes.gray(0.2), diffuse=scenes.gray(0.6))
    floor = scenes.mesh_rect_tris((-8, -8, -1.2), (8, 8, -1.0)).astype(np.float32)
    objs, ospec = R.JoinedObject(), []
    for k in range(100):
        off = (float(k % 10) * 1.4 - 6.3, float(k // 10) * 1.4 - 6.3, float(rng.uniform(-0.3, 0.3)))
        o = R.ColliderObject(Collider=col, Material=mat)
        if k % 3 == 0:
            rot = scenes.rotation((0.0, 0.6, 0.8), 0.1 * k) * (1.0 + 0.2 * (k % 2))
            objs.append(R.Translate(R.MatrixMultiply(o, rot.reshape(-1).tolist()), off))
            ospec.append(dict(kind="mesh", tris=ball, material=m_o, xf=(rot, off)))
        else:
            objs.append(R.Translate(o, off))
            ospec.append(dict(kind="mesh", tris=ball, material=m_o, xf=(np.eye(3), off)))
    objs.append(R.ColliderObject(Collider=floor, Material=mat))
    ospec.append(dict(kind="mesh", tris=floor, material=m_o))
    psc = R.Scene(objs)
    osc = scenes.build_oracle(dict(objects=ospec))
    info = psc.Info()
    assert info["num_triangles"] == 12  # only the floor lives in the scene's own BVH
    assert col.Info()["num_triangles"] == 2880
    n = 200000
    org = rng.uniform(-8, 8, size=(n, 3)).astype(np.float32)
    org[:, 2] = rng.uniform(1.0, 4.0, size=n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[:, 2] = -np.abs(d[:, 2]) - 0.2
    got = psc.Cast(org, d)
    ref = osc.cast(org, d, threads=8)
    hit_o, hit_g = ref["obj"] >= 0, got["obj"] >= 0
    assert (hit_o != hit_g).sum() <= 4
    both = hit_o & hit_g
    same = both & (ref["obj"] == got["obj"]) & (ref["prim"] == got["prim"])
    # a rotated / scaled instance sees the ray through a float32 matrix: a few rays that pass an edge
    # within rounding may pick the neighbouring triangle
    assert same.sum() >= both.sum() - 40, (both.sum(), same.sum())
    rel = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
    assert rel[same].max() < 2e-5
    assert np.abs(got["normal"][same] - ref["normal"][same]).max() < 1e-4
    assert len(set(np.unique(got["obj"][hit_g]).tolist())) == 101
    # path tracing over instances (bounce and shadow rays start on instanced triangles)
    cam = R.NewCameraAt((0.0, -12.0, 9.0), (0.0, 0.0, 0.0), np.pi / 3.5)
    lt = R.PointLight(Origin=(5.0, -10.0, 20.0), Color=(300.0, 300.0, 300.0), QuadDropoff=True)
    tr = R.RecursiveRayTracer(Camera=cam, Lights=[lt], MaxDepth=3, NumSamples=32, Cutoff=1e-4, Antialias=1.0, Seed=5)
    rgb, _, st = tr.RenderSums(96, 72, psc)
    ocam = oracle.camera_at((0.0, -12.0, 9.0), (0.0, 0.0, 0.0), np.pi / 3.5)
    pp = scenes.oracle_path_params(dict(objects=ospec), osc, 3, 64, cutoff=1e-4, antialias=1.0, seed=9)
    ol = oracle.PointLight()
    ol.origin[:], ol.color[:], ol.quad_dropoff = (5.0, -10.0, 20.0), (300.0, 300.0, 300.0), 1
    want = osc.render_path(ocam, [ol], pp, 96, 72, threads=8)
    mean = rgb.astype(np.float64) / 32
    assert abs(mean.mean() - want["mean"].mean()) < 0.03 * want["mean"].mean()
