"""The reference's example scenes as neutral specs (plain dicts) plus the builder that turns a
spec into product objects (model3d_b200.render3d).  bench.py renders them; tests/scenes.py
builds the same specs a second time for the CPU oracle.

Scenes restate the reference's examples:
  c1_scene       marching-cubes sphere in the SaveRendering setup (render3d/helpers.go:105-123, Objectify :71-99)
  cornell_box    examples/renderings/cornell_box/main.go:15-121
  testing_scene  render3d/bidir_test.go:85-111
"""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

from . import meshes  # noqa: E402
from .render3d import NewColorRGB  # noqa: E402


# ---- neutral spec -------------------------------------------------------------------------
def lambert(diffuse=(0, 0, 0), ambient=(0, 0, 0), emission=(0, 0, 0)):
    return dict(kind="lambert", diffuse=diffuse, ambient=ambient, emission=emission)


def phong(alpha, specular=(0, 0, 0), diffuse=(0, 0, 0), ambient=(0, 0, 0), emission=(0, 0, 0)):
    return dict(kind="phong", alpha=alpha, specular=specular, diffuse=diffuse, ambient=ambient, emission=emission)


def refract(ior, color, specular=(0, 0, 0)):
    return dict(kind="refract", ior=ior, refract=color, specular=specular)


def joined(mats, probs):
    return dict(kind="joined", mats=mats, probs=probs)


def gray(b):
    return (b, b, b)


def mesh_rect_tris(mn, mx):
    """NewMeshRect (model3d/mesh.go:132-165), insertion order."""
    mn, mx = np.asarray(mn, np.float64), np.asarray(mx, np.float64)

    def pt(x, y, z):
        return np.array([mx[0] if x else mn[0], mx[1] if y else mn[1], mx[2] if z else mn[2]])

    quads = [(mn, pt(1, 0, 0), pt(1, 0, 1), pt(0, 0, 1)), (mx, pt(1, 1, 0), pt(0, 1, 0), pt(0, 1, 1)),
             (mn, pt(0, 0, 1), pt(0, 1, 1), pt(0, 1, 0)), (mx, pt(1, 0, 1), pt(1, 0, 0), pt(1, 1, 0)),
             (mn, pt(0, 1, 0), pt(1, 1, 0), pt(1, 0, 0)), (mx, pt(0, 1, 1), pt(0, 0, 1), pt(1, 0, 1))]
    tris = []
    for p1, p2, p3, p4 in quads:
        tris.append([p1, p2, p4])
        tris.append([p2, p3, p4])
    return np.array(tris, np.float64)


def rotation(axis, angle):
    """Right-handed rotation about a unit axis (model3d/matrix.go:31-40), row-major 3x3."""
    a = np.asarray(axis, np.float64)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(angle) * k + (1 - math.cos(angle)) * (k @ k)


# ---- builders -------------------------------------------------------------------------------
def build_product(spec, **scene_kw):
    from . import render3d as R
    cache = {}

    def mat(m):
        if id(m) in cache:
            return cache[id(m)]
        if m["kind"] == "lambert":
            r = R.LambertMaterial(DiffuseColor=m["diffuse"], AmbientColor=m["ambient"], EmissionColor=m["emission"])
        elif m["kind"] == "phong":
            r = R.PhongMaterial(Alpha=m["alpha"], SpecularColor=m["specular"], DiffuseColor=m["diffuse"],
                                AmbientColor=m["ambient"], EmissionColor=m["emission"])
        elif m["kind"] == "refract":
            r = R.RefractMaterial(IndexOfRefraction=m["ior"], RefractColor=m["refract"], SpecularColor=m["specular"])
        elif m["kind"] == "joined":
            r = R.JoinedMaterial(Materials=[mat(s) for s in m["mats"]], Probs=list(m["probs"]))
        elif m["kind"] == "checker":
            r = R.CheckerLambertMaterial(Color1=m["color1"], Color2=m["color2"])
        elif m["kind"] == "zgradient":
            r = R.ZGradientPhongMaterial(Alpha=m["alpha"], SpecularColor=m["specular"], Color1=m["color1"],
                                         Color2=m["color2"], MaxZ=m["max_z"])
        cache[id(m)] = r
        return r

    objs = R.JoinedObject()
    light_of = {l["object"]: l for l in spec.get("area_lights", [])}
    for oi, o in enumerate(spec["objects"]):
        m = mat(o["material"])
        if o["kind"] == "mesh":
            c = np.asarray(o["tris"], np.float32)
            if o.get("vnormals") is not None:  # MeshToInterpNormalCollider (collisions.go:147-162)
                from .model3d import MeshToInterpNormalCollider
                c = MeshToInterpNormalCollider(c, np.asarray(o["vnormals"], np.float32))
        elif o["kind"] == "sphere":
            c = R.Sphere(tuple(o["center"]), o["radius"])
        elif o["kind"] == "rect":
            c = R.Rect(tuple(o["min"]), tuple(o["max"]))
        elif o["kind"] == "cylinder":
            c = R.Cylinder(tuple(o["p1"]), tuple(o["p2"]), o["radius"])
        if oi in light_of:
            # an emitter sampled by the BidirPathTracer: AreaLight object (light.go:131-140,237-252)
            obj = R.AreaLight(c, light_of[oi]["emission"])
        else:
            obj = R.ColliderObject(Collider=c, Material=m, FlipNormals=bool(o.get("flip")))
        xf = o.get("xf")
        if xf is not None:
            obj = R.Translate(R.MatrixMultiply(obj, xf[0]), xf[1])
        objs.append(obj)
    scene = R.Scene(objs, **scene_kw)
    scene.material_of = lambda m: cache[id(m)]
    scene.area_light = R.JoinAreaLights(*[objs[l["object"]] for l in spec.get("area_lights", [])]) \
        if spec.get("area_lights") else None
    return scene


# ---- scenes -----------------------------------------------------------------------------------
def mixed_scene():
    m_l = lambert(diffuse=gray(0.4), ambient=gray(0.05))
    m_p = phong(10.0, specular=gray(0.2), diffuse=(0.5, 0.4, 0.0), ambient=(0.05, 0.04, 0.0))
    m_e = lambert(emission=gray(2.0))
    ico = meshes.NewMeshIcosphere((0.5, 0.5, 0.0), 0.8, 12).astype(np.float32)
    box = mesh_rect_tris((-3, -3, -1.5), (3, 3, -1.2)).astype(np.float32)
    objs = [
        dict(kind="sphere", center=(-1.5, 0.2, 0.1), radius=0.7, material=m_p),
        dict(kind="mesh", tris=ico, material=m_l),
        dict(kind="rect", min=(1.4, -0.5, -0.8), max=(2.2, 0.6, 0.9), material=m_p),
        dict(kind="cylinder", p1=(-0.4, -1.8, -1.0), p2=(0.3, -1.2, 0.8), radius=0.35, material=m_l),
        dict(kind="mesh", tris=box, material=m_l),
        dict(kind="sphere", center=(0.0, 2.0, 1.5), radius=0.3, material=m_e),
        dict(kind="sphere", center=(1.0, 0.0, 0.0), radius=0.5, material=m_p,
             xf=(rotation((0, 0, 1), 0.7) * 1.5, (0.3, -2.5, 0.4))),
    ]
    return dict(objects=objs)


def c1_scene(n=None, delta=0.01, iters=8):
    """BASELINE config 1: model3d.Sphere{radius 1} -> MarchingCubesSearch(0.01, 8) (376,832
    triangles) set up the way SaveRendering does (helpers.go:101-128: camera at `origin`
    looking at the bounding-box centre, one far point light behind the camera, Objectify's
    default yellow Phong material).  `n` selects an icosphere NewMeshIcosphere(0,1,n) instead
    (small test meshes)."""
    yellow = NewColorRGB(224.0 / 255, 209.0 / 255, 0.0)
    mat = phong(10.0, specular=gray(0.2), diffuse=tuple(0.8 * c for c in yellow),
                ambient=tuple(0.1 * c for c in yellow))
    if n is None:
        tris = meshes.MarchingCubesSearch(meshes.SphereSolid((0, 0, 0), 1.0), delta, iters).astype(np.float32)
    else:
        tris = meshes.NewMeshIcosphere((0, 0, 0), 1.0, n).astype(np.float32)
    origin = np.array([2.0, -3.0, 1.5])
    v = tris.reshape(-1, 3).astype(np.float64)
    center = (v.min(0) + v.max(0)) * 0.5  # Mesh.Min().Mid(Mesh.Max()) (helpers.go:107-108)
    return dict(objects=[dict(kind="mesh", tris=tris, material=mat)],
                camera=dict(src=tuple(origin), dst=tuple(center), fov=math.pi / 3.6),
                lights=[dict(origin=tuple(center + (origin - center) * 1000), color=gray(1.0))])


def diamond_tris():
    """LoadDiamond (cornell_box/main.go:127-144): diamond.stl rotated about Y by
    pi/2 + atan(1/1.2), then translated by (0, 4, -(2 + min.z))."""
    tris = np.load(os.path.join(GOLDEN, "diamond_tris.npy")).astype(np.float64)
    rot = rotation((0, 1, 0), 0.5 * math.pi + math.atan(1 / 1.2))
    t = tris.reshape(-1, 3) @ rot.T
    t = t + np.array([0.0, 4.0, -(2 + t[:, 2].min())])
    return t.reshape(-1, 3, 3).astype(np.float32)


def cornell_box(area_light_mesh=True):
    red = NewColorRGB(0.95, 0.2, 0.2)
    m_mirror = phong(400.0, specular=gray(1.0))
    m_red = phong(10.0, specular=gray(0.1), diffuse=tuple(0.5 * c for c in red))
    m_refr = refract(1.3, gray(0.9))
    m_ph50 = phong(50.0, specular=gray(0.1))
    m_glass = joined([m_refr, m_ph50], [0.9, 0.1])
    m_wall = lambert(diffuse=gray(0.4))
    m_light = lambert(emission=gray(25.0))
    walls = mesh_rect_tris((-5, -10, -2), (5, 10, 7)) * np.array([-1.0, 1.0, 1.0])  # MapCoords(XYZ(-1,1,1).Mul)
    light = mesh_rect_tris((-2, 5, 6.8), (2, 7, 7))
    objs = [
        dict(kind="sphere", center=(2, 7, 0), radius=2.0, material=m_mirror),
        dict(kind="sphere", center=(-2, 5.5, -1), radius=1.0, material=m_red),
        dict(kind="mesh", tris=diamond_tris(), material=m_glass),
        dict(kind="mesh", tris=walls.astype(np.float32), material=m_wall),
        dict(kind="mesh", tris=light.astype(np.float32), material=m_light),
    ]
    return dict(objects=objs, camera=dict(src=(0, -7, 2.5), dst=(0, 10, 2.5), fov=math.pi / 3.6),
                focus=[dict(kind="phong", target=(0, 6, 6.9), alpha=40.0, prob=0.3,
                            applies=lambda m: m["kind"] == "lambert" or (m["kind"] == "phong" and sum(m["diffuse"]) > 0))],
                area_lights=[dict(object=4, emission=gray(25.0))])


def testing_scene():
    """render3d/bidir_test.go:85-111: point-mirrored (hence inward-facing) Lambert(0.3) box
    NewMeshRect((-10,-10,-10),(10,20,0)).Scale(-1) + two sphere area lights; camera of
    TestBidirPathTracer (bidir_test.go:14)."""
    m_wall = lambert(diffuse=gray(0.3))
    box = mesh_rect_tris((-10, -10, -10), (10, 20, 0)) * -1.0
    l1 = dict(kind="sphere", center=(0, -19, 5), radius=1.0, material=lambert(emission=gray(100.0)))
    l2 = dict(kind="sphere", center=(3, -19, 5), radius=0.5, material=lambert(emission=gray(130.0)))
    return dict(objects=[dict(kind="mesh", tris=box.astype(np.float32), material=m_wall), l1, l2],
                camera=dict(src=(0, -17, 2), dst=(0, 0, 2), fov=math.pi / 3.6),
                area_lights=[dict(object=1, emission=gray(100.0)), dict(object=2, emission=gray(130.0))],
                focus=[dict(kind="sphere", target=(0, -19, 5), radius=1.0, prob=0.2, applies=lambda m: True),
                       dict(kind="sphere", target=(3, -19, 5), radius=0.5, prob=0.1, applies=lambda m: True)])


# ---- renderer parameter builders (oracle / product) from one spec ----------------------------
def all_materials(spec):
    """Materials of the spec in first-use order, sub-materials of joined ones first (the order
    both builders register them in)."""
    out = []

    def visit(m):
        if any(m is x for x in out):
            return
        if m["kind"] == "joined":
            for s in m["mats"]:
                visit(s)
        out.append(m)

    for o in spec["objects"]:
        visit(o["material"])
    return out


def product_tracer(spec, psc, max_depth, num_samples, cutoff=0.0, antialias=0.0, seed=1, lights=()):
    from . import render3d as R
    cam = spec["camera"]
    fps, probs = [], []
    for f in spec.get("focus", []):
        ok = [psc.material_of(m) for m in all_materials(spec) if f["applies"](m)]
        flt = (lambda mats: (lambda m: any(m is x for x in mats)))(ok)
        if f["kind"] == "phong":
            fps.append(R.PhongFocusPoint(Target=f["target"], Alpha=f["alpha"], MaterialFilter=flt))
        else:
            fps.append(R.SphereFocusPoint(Center=f["target"], Radius=f["radius"], MaterialFilter=flt))
        probs.append(f["prob"])
    return R.RecursiveRayTracer(Camera=R.NewCameraAt(cam["src"], cam["dst"], cam["fov"]), Lights=list(lights),
                                FocusPoints=fps, FocusPointProbs=probs, MaxDepth=max_depth,
                                NumSamples=num_samples, Cutoff=cutoff, Antialias=antialias, Seed=seed)


def product_bidir(spec, psc, max_depth, num_samples, min_depth=0, roulette_delta=0.0, power_heuristic=0.0,
                  cutoff=0.0, antialias=0.0, max_light_depth=0, seed=1):
    from . import render3d as R
    cam = spec["camera"]
    return R.BidirPathTracer(Camera=R.NewCameraAt(cam["src"], cam["dst"], cam["fov"]), Light=psc.area_light,
                             MaxDepth=max_depth, MaxLightDepth=max_light_depth, MinDepth=min_depth,
                             RouletteDelta=roulette_delta, PowerHeuristic=power_heuristic, NumSamples=num_samples,
                             Cutoff=cutoff, Antialias=antialias, Seed=seed)


def glass_scene():
    """Refraction with Fresnel reflection (RefractMaterial with SpecularColor), a glass ball
    and a glass slab inside a lit Lambert room: exercises every Dirac-lobe branch
    (refract / reflect / total internal reflection) of material.go:343-479."""
    m_wall = lambert(diffuse=gray(0.5))
    m_light = lambert(emission=gray(12.0))
    m_glass = refract(1.5, gray(0.95), specular=gray(0.9))
    m_clear = refract(1.2, (0.9, 0.95, 1.0))
    walls = mesh_rect_tris((-4, -4, -2), (4, 6, 4)) * np.array([-1.0, 1.0, 1.0])
    light = mesh_rect_tris((-1.5, 0, 3.8), (1.5, 3, 3.9))
    slab = mesh_rect_tris((-3.0, 3.0, -1.5), (-1.0, 3.4, 1.5))
    objs = [
        dict(kind="mesh", tris=walls.astype(np.float32), material=m_wall),
        dict(kind="mesh", tris=light.astype(np.float32), material=m_light),
        dict(kind="sphere", center=(0.8, 2.0, -0.8), radius=1.2, material=m_glass),
        dict(kind="mesh", tris=slab.astype(np.float32), material=m_clear),
        dict(kind="cylinder", p1=(2.5, 4.0, -2.0), p2=(2.5, 4.0, 0.5), radius=0.6,
             material=phong(30.0, specular=gray(0.3), diffuse=(0.1, 0.3, 0.5))),
    ]
    return dict(objects=objs, camera=dict(src=(0, -3.5, 1.0), dst=(0, 4, 0.2), fov=math.pi / 3.0),
                area_lights=[dict(object=1, emission=gray(12.0))])


def checker(color1, color2):
    """showcase FloorObject (room.go:61-75): Lambert, color2 where the checker test holds."""
    return dict(kind="checker", color1=color1, color2=color2)


def showcase_models():
    """The seven meshes of examples/renderings/showcase/models (tests/golden/showcase_models.npz,
    written by tests/golden/make_fixtures.py from the reference's *.stl.gz), float64 [n,3,3]."""
    z = np.load(os.path.join(GOLDEN, "showcase_models.npz"))
    names = sorted(k[:-2] for k in z.files if k.endswith("_v"))
    return {n: z[n + "_v"][z[n + "_f"]].astype(np.float64) for n in names}


def showcase(hd=False):
    """BASELINE config 4: examples/renderings/showcase (main.go, room.go, models.go,
    constants.go).  The vase is omitted: vase.stl.gz is not part of the reference checkout.
    Go closures become declarative materials / flags: FloorObject -> checker Lambert,
    DomeObject -> flipped normals, the focus point's MaterialFilter -> Lambert/Phong kinds."""
    M = showcase_models()
    light_dir = np.array([2.0, -3.0, 3.0])
    light_dir = light_dir * (1.0 / math.sqrt(float(light_dir @ light_dir)))
    light_center = tuple((light_dir * (1.0 / math.sqrt(float(light_dir @ light_dir))) * 50.0).tolist())
    room_radius, light_radius = 100.0, 5.0

    def bounds(t):
        p = t.reshape(-1, 3)
        return p.min(axis=0), p.max(axis=0)

    def rot(t, axis, angle):
        return (t.reshape(-1, 3) @ rotation(axis, angle).T).reshape(-1, 3, 3)

    def f32(t):
        return np.ascontiguousarray(t, np.float32)

    objs = []
    # NewFloorObject / NewDomeObject / NewLightObject (room.go)
    objs.append(dict(kind="rect", min=(-room_radius, -room_radius, -0.01), max=(room_radius, room_radius, 0.0),
                     material=checker(gray(0.6), gray(0.1))))
    sky = tuple(0.15 * c + 0.15 for c in NewColorRGB(0.5, 0.8, 0.95))
    objs.append(dict(kind="sphere", center=(0.0, 0.0, 0.0), radius=room_radius, material=lambert(diffuse=sky),
                     flip=True))
    objs.append(dict(kind="sphere", center=light_center, radius=light_radius,
                     material=lambert(emission=gray(300.0))))
    # ReadRose (models.go:33-59)
    t = M["rose"]
    mn, mx = bounds(t)
    t = t - (mn + mx) / 2
    t = rot(t, (1, 0, 0), math.pi / 4) + np.array([5.0, 11.0, 7.0])
    objs.append(dict(kind="mesh", tris=f32(t),
                     material=lambert(diffuse=tuple(0.5 * c for c in NewColorRGB(0.95, 0.2, 0.2)))))
    objs.append(dict(kind="cylinder", p1=(5.0, 11.0, 0.0), p2=(5.0, 11.0, 7.0), radius=0.15,
                     material=lambert(diffuse=tuple(0.5 * c for c in NewColorRGB(0.1, 0.55, 0.0)))))
    # ReadWineGlass (models.go:144-176)
    t = M["wine_glass"]
    mn, mx = bounds(t)
    t = t - mn
    t = t + np.array([-5.0 - (mx[0] - mn[0]) / 2, 10.5 - (mx[1] - mn[1]) / 2, 1e-4])
    glass = joined([refract(1.3, gray(0.95)), phong(100.0, specular=gray(0.05))], [0.8, 0.2])
    objs.append(dict(kind="mesh", tris=f32(t), material=glass))
    # ReadPumpkin (models.go:113-142)
    colors = [NewColorRGB(255.0 / 255, 206.0 / 255, 107.0 / 255), NewColorRGB(214.0 / 255, 143.0 / 255, 0),
              NewColorRGB(79.0 / 255, 53.0 / 255, 0)]
    for name, col in zip(["pumpkin_inside", "pumpkin_outside", "pumpkin_stem"], colors):
        t = M[name] + np.array([-2.0, 10.0, 1.1942578125000005])
        objs.append(dict(kind="mesh", tris=f32(t), material=lambert(diffuse=tuple(0.5 * c for c in col))))
    # ReadRocks (models.go:100-111)
    t = M["rocks"]
    mn, mx = bounds(t)
    t = t + np.array([-(mx[0] + mn[0]) / 2, 15.0 - mn[1], 0.0])
    objs.append(dict(kind="mesh", tris=f32(t), material=lambert(diffuse=gray(0.3))))
    # ReadCurvyThing (models.go:14-31)
    t = M["curvy_thing"]
    mn, mx = bounds(t)
    inv_mid = -(mn + mx) / 2
    inv_mid[2] = -mn[2]
    t = rot(t + inv_mid, (0, 0, 1), -math.pi / 4) + np.array([1.8, 9.0, 0.0])
    objs.append(dict(kind="mesh", tris=f32(t), material=phong(20.0, specular=gray(0.05), diffuse=gray(0.3))))
    size = (960, 640) if hd else (480, 320)
    return dict(objects=objs, camera=dict(src=(0.0, -5.0, 4.0), dst=(0.0, room_radius, 4.0), fov=math.pi / 3.6),
                focus=[dict(kind="sphere", target=light_center, radius=light_radius, prob=0.3,
                            applies=lambda m: m["kind"] in ("lambert", "phong", "checker", "zgradient"))],
                size=size, max_depth=10, cutoff=1e-4, antialias=1.0)
