"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so, the float64 CPU restatement of the
reference's ray-tracing path (see oracle/vec.hpp for the parity status).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package
(model3d_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)


class Camera(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("screen_x", C.c_double * 3),
                ("screen_y", C.c_double * 3), ("field_of_view", C.c_double)]


class PointLight(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("color", C.c_double * 3),
                ("quad_dropoff", C.c_int32), ("_pad", C.c_int32)]


class MaterialDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_uint32),
                ("diffuse", C.c_double * 3), ("specular", C.c_double * 3),
                ("emission", C.c_double * 3), ("ambient", C.c_double * 3),
                ("refract", C.c_double * 3), ("alpha", C.c_double),
                ("index_of_refraction", C.c_double), ("diffuse2", C.c_double * 3),
                ("proc_param", C.c_double), ("num_sub", C.c_int32),
                ("sub", C.c_int32 * 4), ("sub_prob", C.c_double * 4)]


class Transform(C.Structure):
    _fields_ = [("matrix", C.c_double * 9), ("offset", C.c_double * 3)]


class FocusPoint(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("target", C.c_double * 3),
                ("alpha", C.c_double), ("radius", C.c_double),
                ("material_mask", C.c_uint64), ("prob", C.c_double)]


class PathParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("num_samples", C.c_int32),
                ("min_samples", C.c_int32), ("num_focus_points", C.c_int32),
                ("max_stddev", C.c_double), ("oversaturated_stddevs", C.c_double),
                ("cutoff", C.c_double), ("antialias", C.c_double), ("epsilon", C.c_double),
                ("focus", FocusPoint * 4), ("seed", C.c_uint64)]


class AreaLight(C.Structure):
    _fields_ = [("object", C.c_int32), ("_pad", C.c_int32), ("emission", C.c_double * 3)]


class BidirParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("max_light_depth", C.c_int32),
                ("min_depth", C.c_int32), ("num_samples", C.c_int32),
                ("roulette_delta", C.c_double), ("power_heuristic", C.c_double),
                ("cutoff", C.c_double), ("antialias", C.c_double), ("epsilon", C.c_double),
                ("seed", C.c_uint64), ("min_samples", C.c_int32), ("_pad", C.c_int32),
                ("max_stddev", C.c_double), ("oversaturated_stddevs", C.c_double)]


def build(force=False):
    """Compile liboracle.so with oracle/Makefile (gcc only; no GPU needed)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "m3d.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_collider_create.restype = C.c_void_p
        L.orc_collider_create_f64.restype = C.c_void_p
        L.orc_scene_create.restype = C.c_void_p
        L.orc_mesh_icosphere_count.restype = C.c_int64
        L.orc_mesh_polar_count.restype = C.c_int64
        L.orc_mesh_mc_sphere_build.restype = C.c_int64
        L.orc_collider_num_nodes.restype = C.c_int64
        L.orc_collider_all_hits.restype = C.c_int64
        L.orc_gamma_expand.restype = C.c_double
        L.orc_gamma_expand.argtypes = [C.c_double]
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def hardware_threads():
    return int(lib().orc_hardware_threads())


def gamma_expand(u):
    return float(lib().orc_gamma_expand(float(u)))


def color_rgb(r, g, b):
    """render3d.NewColorRGB (light.go:32-34)."""
    return (gamma_expand(r), gamma_expand(g), gamma_expand(b))


# ---- meshes --------------------------------------------------------------------
def mesh_icosphere(center, radius, n):
    cnt = lib().orc_mesh_icosphere_count(C.c_int(n))
    out = np.empty((cnt, 3, 3), np.float64)
    lib().orc_mesh_icosphere(C.c_double(center[0]), C.c_double(center[1]), C.c_double(center[2]),
                             C.c_double(radius), C.c_int(n), _p(out, f64p))
    return out


def mesh_rect(mn, mx):
    out = np.empty((12, 3, 3), np.float64)
    lib().orc_mesh_rect(_d3(mn), _d3(mx), _p(out, f64p))
    return out


def mesh_mc_sphere(center, radius, delta, iters):
    """MarchingCubesSearch(&Sphere{center, radius}, delta, iters) (mc.go:45-51), scan order."""
    cnt = lib().orc_mesh_mc_sphere_build(C.c_double(center[0]), C.c_double(center[1]), C.c_double(center[2]),
                                         C.c_double(radius), C.c_double(delta), C.c_int(iters))
    if cnt < 0:
        raise RuntimeError("oracle marching cubes failed")
    out = np.empty((cnt, 3, 3), np.float64)
    lib().orc_mesh_mc_sphere_fetch(_p(out, f64p))
    return out


def mc_table():
    """mcLookupTable (mc.go:431-454): (counts[256], corners[256,5,6])."""
    counts = np.zeros(256, np.int32)
    corners = np.zeros((256, 5, 6), np.uint8)
    lib().orc_mc_table(_p(counts, C.POINTER(C.c_int32)), _p(corners, C.POINTER(C.c_uint8)))
    return counts, corners


def triangle_closest(tri, p):
    """Triangle.Closest (primitives.go:153-175)."""
    tri = np.ascontiguousarray(tri, np.float64).reshape(9)
    out = np.empty(3, np.float64)
    lib().orc_triangle_closest(_p(tri, f64p), _d3(p), _p(out, f64p))
    return out


def triangle_sphere_collision(tri, c, r):
    """Triangle.SphereCollision (primitives.go:253-279)."""
    tri = np.ascontiguousarray(tri, np.float64).reshape(9)
    return bool(lib().orc_triangle_sphere_collision(_p(tri, f64p), _d3(c), C.c_double(r)))


def mesh_polar(ra, rb, stops):
    cnt = lib().orc_mesh_polar_count(C.c_int(stops))
    out = np.empty((cnt, 3, 3), np.float64)
    lib().orc_mesh_polar(C.c_double(ra), C.c_double(rb), C.c_int(stops), _p(out, f64p))
    return out


# ---- collider ----------------------------------------------------------------------
class Collider:
    """MeshToCollider restatement (collisions.go:138-142)."""

    def __init__(self, tris, vnormals=None):
        tris = np.ascontiguousarray(tris)
        self.n = int(tris.shape[0])
        if tris.dtype == np.float32:
            vn = None if vnormals is None else np.ascontiguousarray(vnormals, np.float32)
            self.h = lib().orc_collider_create(_p(tris, f32p), C.c_int64(self.n), _p(vn, f32p))
        else:
            assert vnormals is None
            tris = np.ascontiguousarray(tris, np.float64)
            self.h = lib().orc_collider_create_f64(_p(tris, f64p), C.c_int64(self.n))
        self.h = C.c_void_p(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_collider_destroy(self.h)
            self.h = None

    def bounds(self):
        mn, mx = np.zeros(3), np.zeros(3)
        lib().orc_collider_bounds(self.h, _p(mn, f64p), _p(mx, f64p))
        return mn, mx

    def num_nodes(self):
        return int(lib().orc_collider_num_nodes(self.h))

    def order(self):
        out = np.empty(self.n, np.int32)
        lib().orc_collider_order(self.h, _p(out, i32p))
        return out

    def first_hits(self, org, dir, threads=1, counters=False):
        """Returns dict(t, prim, normal, bary[, nodes, tris])."""
        n = int(org.shape[0])
        t = np.zeros(n, np.float64)
        prim = np.full(n, -1, np.int32)
        normal = np.zeros((n, 3), np.float64)
        bary = np.zeros((n, 3), np.float64)
        cnt = np.zeros(2, np.int64) if counters else None
        if org.dtype == np.float32 and dir.dtype == np.float32:
            org = np.ascontiguousarray(org)
            dir = np.ascontiguousarray(dir)
            lib().orc_collider_first_hits_f32(self.h, _p(org, f32p), _p(dir, f32p), C.c_int64(n),
                                              _p(t, f64p), _p(prim, i32p), _p(normal, f64p),
                                              _p(bary, f64p), _p(cnt, i64p), C.c_int(threads))
        else:
            org = np.ascontiguousarray(org, np.float64)
            dir = np.ascontiguousarray(dir, np.float64)
            lib().orc_collider_first_hits(self.h, _p(org, f64p), _p(dir, f64p), C.c_int64(n),
                                          _p(t, f64p), _p(prim, i32p), _p(normal, f64p),
                                          _p(bary, f64p), _p(cnt, i64p), C.c_int(threads))
        res = dict(t=t, prim=prim, normal=normal, bary=bary)
        if counters:
            res["nodes"], res["tris"] = int(cnt[0]), int(cnt[1])
        return res

    def hit_counts(self, org, dir, threads=1):
        org = np.ascontiguousarray(org, np.float32)
        dir = np.ascontiguousarray(dir, np.float32)
        out = np.zeros(org.shape[0], np.int32)
        lib().orc_collider_hit_counts(self.h, _p(org, f32p), _p(dir, f32p), C.c_int64(org.shape[0]),
                                      _p(out, i32p), C.c_int(threads))
        return out

    def all_hits_batch(self, org, dir, threads=1):
        """Collider.RayCollisions(r, f) per ray (collisions.go:263-273): dict(offsets, t, prim,
        normal, bary), the hits of ray i in rows offsets[i]:offsets[i+1] ordered by t."""
        org = np.ascontiguousarray(org, np.float32)
        dir = np.ascontiguousarray(dir, np.float32)
        n = org.shape[0]
        counts = self.hit_counts(org, dir, threads)
        offsets = np.zeros(n + 1, np.int64)
        np.cumsum(counts, out=offsets[1:])
        total = int(offsets[n])
        t = np.zeros(total, np.float64)
        prim = np.zeros(total, np.int32)
        normal = np.zeros((total, 3), np.float64)
        bary = np.zeros((total, 3), np.float64)
        lib().orc_collider_all_hits_batch(self.h, _p(org, f32p), _p(dir, f32p), C.c_int64(n),
                                          _p(offsets, i64p), _p(t, f64p), _p(prim, i32p),
                                          _p(normal, f64p), _p(bary, f64p), C.c_int(threads))
        return dict(offsets=offsets, t=t, prim=prim, normal=normal, bary=bary)

    def sdf(self, pts, threads=1):
        """MeshToSDF(mesh).FaceSDF per point (sdf.go:229-240): (sdf, nearest point, face id)."""
        pts = np.ascontiguousarray(pts, np.float32)
        n = pts.shape[0]
        sdf = np.empty(n, np.float64)
        point = np.empty((n, 3), np.float64)
        face = np.empty(n, np.int32)
        lib().orc_collider_sdf(self.h, _p(pts, f32p), C.c_int64(n), _p(sdf, f64p), _p(point, f64p),
                               _p(face, i32p), C.c_int(threads))
        return sdf, point, face

    def sphere_collisions(self, centers, radii, threads=1):
        """Collider.SphereCollision per (center, radius) (collisions.go:292-303)."""
        centers = np.ascontiguousarray(centers, np.float32)
        radii = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float64), centers.shape[:1]))
        out = np.empty(centers.shape[0], np.uint8)
        lib().orc_collider_sphere_collisions(self.h, _p(centers, f32p), _p(radii, f64p),
                                             C.c_int64(centers.shape[0]), _p(out, u8p), C.c_int(threads))
        return out.astype(bool)

    def contains_margin(self, pts, margin, solid=0, threads=1):
        """ColliderContains(c, p, margin) (collisions.go:119-134); solid=1: ColliderSolid.Contains of
        NewColliderSolid, solid=2: of NewColliderSolidInset(c, margin) (solid.go:256-300)."""
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.empty(pts.shape[0], np.uint8)
        lib().orc_collider_contains_margin(self.h, _p(pts, f32p), C.c_int64(pts.shape[0]), C.c_double(margin),
                                           C.c_int(solid), _p(out, u8p), C.c_int(threads))
        return out.astype(bool)

    def contains(self, pts, threads=1):
        pts = np.ascontiguousarray(pts, np.float32)
        out = np.zeros(pts.shape[0], np.uint8)
        lib().orc_collider_contains(self.h, _p(pts, f32p), C.c_int64(pts.shape[0]), _p(out, u8p), C.c_int(threads))
        return out.astype(bool)

    def all_hits(self, org, dir, brute=False, cap=4096):
        t = np.zeros(cap, np.float64)
        prim = np.zeros(cap, np.int32)
        n = lib().orc_collider_all_hits(self.h, _d3(org), _d3(dir), C.c_int(1 if brute else 0),
                                        _p(t, f64p), _p(prim, i32p), C.c_int64(cap))
        n = int(n)
        return n, t[:min(n, cap)], prim[:min(n, cap)]


def triangle_first_hit(tri, org, dir):
    tri = np.ascontiguousarray(tri, np.float64).reshape(9)
    t = C.c_double(0)
    normal, bary = np.zeros(3), np.zeros(3)
    ok = lib().orc_triangle_first_hit(_p(tri, f64p), _d3(org), _d3(dir), C.byref(t),
                                      _p(normal, f64p), _p(bary, f64p))
    return (bool(ok), t.value, normal, bary)


SPHERE, RECT, CYLINDER = 1, 2, 3


def shape_first_hit(kind, p0, p1, radius, org, dir):
    t = C.c_double(0)
    normal = np.zeros(3)
    ok = lib().orc_shape_first_hit(C.c_int(kind), _d3(p0), _d3(p1), C.c_double(radius), _d3(org),
                                   _d3(dir), C.byref(t), _p(normal, f64p))
    return (bool(ok), t.value, normal)


# ---- scene -----------------------------------------------------------------------------
def _xf(xf):
    if xf is None:
        return None
    m, off = xf
    t = Transform()
    t.matrix[:] = [float(x) for x in np.asarray(m, np.float64).reshape(9)]
    t.offset[:] = [float(x) for x in off]
    return C.byref(t)


class Scene:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_scene_create())
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_destroy(self.h)
            self.h = None

    def add_material(self, desc):
        return int(lib().orc_scene_add_material(self.h, C.byref(desc)))

    def add_mesh(self, tris, material, vnormals=None, flags=0, xf=None):
        tris = np.ascontiguousarray(tris, np.float32)
        vn = None if vnormals is None else np.ascontiguousarray(vnormals, np.float32)
        return int(lib().orc_scene_add_mesh(self.h, _p(tris, f32p), C.c_int64(tris.shape[0]),
                                            _p(vn, f32p), C.c_int32(material), C.c_uint32(flags), _xf(xf)))

    def add_sphere(self, center, radius, material, flags=0, xf=None):
        return int(lib().orc_scene_add_sphere(self.h, _d3(center), C.c_double(radius),
                                              C.c_int32(material), C.c_uint32(flags), _xf(xf)))

    def add_rect(self, mn, mx, material, flags=0, xf=None):
        return int(lib().orc_scene_add_rect(self.h, _d3(mn), _d3(mx), C.c_int32(material),
                                            C.c_uint32(flags), _xf(xf)))

    def add_cylinder(self, p1, p2, radius, material, flags=0, xf=None):
        return int(lib().orc_scene_add_cylinder(self.h, _d3(p1), _d3(p2), C.c_double(radius),
                                                C.c_int32(material), C.c_uint32(flags), _xf(xf)))

    def cast(self, org, dir, threads=1):
        org = np.ascontiguousarray(org, np.float64)
        dir = np.ascontiguousarray(dir, np.float64)
        n = org.shape[0]
        t = np.zeros(n)
        obj = np.full(n, -1, np.int32)
        prim = np.full(n, -1, np.int32)
        normal = np.zeros((n, 3))
        lib().orc_scene_cast(self.h, _p(org, f64p), _p(dir, f64p), C.c_int64(n), _p(t, f64p),
                             _p(obj, i32p), _p(prim, i32p), _p(normal, f64p), C.c_int(threads))
        return dict(t=t, obj=obj, prim=prim, normal=normal)

    def render_raycast(self, cam, lights, W, H, threads=1, img=None):
        if img is None:
            img = np.zeros((H, W, 3), np.float64)
        t = np.zeros((H, W))
        obj = np.zeros((H, W), np.int32)
        prim = np.zeros((H, W), np.int32)
        arr = (PointLight * max(1, len(lights)))(*lights)
        lib().orc_render_raycast(self.h, C.byref(cam), arr, C.c_int(len(lights)), C.c_int(W), C.c_int(H),
                                 _p(img, f64p), _p(t, f64p), _p(obj, i32p), _p(prim, i32p),
                                 C.c_int(threads))
        return dict(img=img, t=t, obj=obj, prim=prim)

    def render_path(self, cam, lights, params, W, H, threads=1):
        mean = np.zeros((H, W, 3))
        var = np.zeros((H, W, 3))
        rays = C.c_int64(0)
        arr = (PointLight * max(1, len(lights)))(*lights)
        lib().orc_render_path(self.h, C.byref(cam), arr, C.c_int(len(lights)), C.byref(params),
                              C.c_int(W), C.c_int(H), _p(mean, f64p), _p(var, f64p), C.byref(rays),
                              C.c_int(threads))
        return dict(mean=mean, var_of_mean=var, rays=rays.value)

    def render_bidir(self, cam, area_lights, params, W, H, threads=1):
        mean = np.zeros((H, W, 3))
        var = np.zeros((H, W, 3))
        rays = C.c_int64(0)
        arr = (AreaLight * max(1, len(area_lights)))(*area_lights)
        lib().orc_render_bidir(self.h, C.byref(cam), arr, C.c_int(len(area_lights)), C.byref(params),
                               C.c_int(W), C.c_int(H), _p(mean, f64p), _p(var, f64p), C.byref(rays),
                               C.c_int(threads))
        return dict(mean=mean, var_of_mean=var, rays=rays.value)

    def material_eval(self, mat, normal, source, dest):
        bsdf = np.zeros(3)
        sd, dd = C.c_double(0), C.c_double(0)
        lib().orc_material_eval(self.h, C.c_int32(mat), _d3(normal), _d3(source), _d3(dest),
                                _p(bsdf, f64p), C.byref(sd), C.byref(dd))
        return bsdf, sd.value, dd.value

    def material_eval_batch(self, mat, normal, sources, dests):
        sources = np.ascontiguousarray(sources, np.float64)
        dests = np.ascontiguousarray(np.broadcast_to(dests, sources.shape), np.float64)
        n = sources.shape[0]
        bsdf, sd, dd = np.zeros((n, 3)), np.zeros(n), np.zeros(n)
        lib().orc_material_eval_batch(self.h, C.c_int32(mat), _d3(normal), _p(sources, f64p), _p(dests, f64p),
                                      C.c_int64(n), _p(bsdf, f64p), _p(sd, f64p), _p(dd, f64p))
        return bsdf, sd, dd

    def material_sample_source(self, mat, seed, normal, dest, n):
        out = np.zeros((n, 3))
        lib().orc_material_sample_source(self.h, C.c_int32(mat), C.c_uint64(seed), _d3(normal),
                                         _d3(dest), C.c_int64(n), _p(out, f64p))
        return out

    def material_sample_dest(self, mat, seed, normal, source, n):
        out = np.zeros((n, 3))
        lib().orc_material_sample_dest(self.h, C.c_int32(mat), C.c_uint64(seed), _d3(normal),
                                       _d3(source), C.c_int64(n), _p(out, f64p))
        return out


def sample_around_uniform(seed, min_cos, direction, n):
    """sampleAroundUniform + densityAroundUniform (focus_point.go:155-177): (directions [n,3], densities [n])."""
    out = np.zeros((n, 3))
    dens = np.zeros(n)
    lib().orc_sample_around_uniform(C.c_uint64(seed), C.c_double(min_cos), _d3(direction), C.c_int64(n),
                                    _p(out, f64p), _p(dens, f64p))
    return out, dens


def camera_at(src, dst, fov):
    cam = Camera()
    lib().orc_camera_at(_d3(src), _d3(dst), C.c_double(fov), C.byref(cam))
    return cam


def camera_rays(cam, W, H):
    out = np.zeros((H * W, 3))
    lib().orc_camera_rays(C.byref(cam), C.c_int(W), C.c_int(H), _p(out, f64p))
    return out


def srgb8(rgb):
    rgb = np.ascontiguousarray(rgb, np.float64)
    out = np.zeros(rgb.shape, np.uint8)
    lib().orc_srgb8(_p(rgb, f64p), C.c_int64(rgb.size), _p(out, u8p))
    return out
