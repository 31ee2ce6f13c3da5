"""Host-side mirror of the reference's ``render3d`` interface for the GPU path.

Same names and field meanings as render3d (Go): ``Camera`` / ``NewCameraAt``
(render3d/camera.go:19-66), ``PointLight`` (light.go:57-66), ``LambertMaterial`` /
``PhongMaterial`` / ``RefractMaterial`` / ``JoinedMaterial`` (material.go:119-631),
``ColliderObject`` / ``JoinedObject`` / ``Translate`` / ``MatrixMultiply`` (object.go:26-153,
transform.go:6-85), ``Image`` (image.go:17-47), ``RayCaster`` (raycast.go:9-39),
``RecursiveRayTracer`` (raytrace.go:14-119), ``BidirPathTracer`` (bidir.go:14-84).
Objects are compiled by a type switch into a device scene through libm3dgpu's C ABI;
unsupported object / collider / material types raise ``UnsupportedError`` -- there is no
CPU fallback.
"""
import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .model3d import MeshCollider, UnsupportedError

Vec = Tuple[float, float, float]
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


# ---- colours (light.go:19-55) -------------------------------------------------------
def NewColor(b):
    return (float(b), float(b), float(b))


def _gamma_expand(u):
    return u / 12.92 if u <= 0.04045 else math.pow((u + 0.055) / 1.055, 2.4)


def NewColorRGB(r, g, b):
    return (_gamma_expand(r), _gamma_expand(g), _gamma_expand(b))


def _scale(c, s):
    return (c[0] * s, c[1] * s, c[2] * s)


# ---- analytic colliders (model3d/shapes.go) ------------------------------------------
@dataclass
class Sphere:
    Center: Vec = (0.0, 0.0, 0.0)
    Radius: float = 1.0


@dataclass
class Rect:
    MinVal: Vec = (0.0, 0.0, 0.0)
    MaxVal: Vec = (1.0, 1.0, 1.0)


@dataclass
class Cylinder:
    P1: Vec = (0.0, 0.0, 0.0)
    P2: Vec = (0.0, 0.0, 1.0)
    Radius: float = 1.0


# ---- materials (material.go) ------------------------------------------------------------
ZERO = (0.0, 0.0, 0.0)


@dataclass(eq=False)
class LambertMaterial:
    DiffuseColor: Vec = ZERO
    AmbientColor: Vec = ZERO
    EmissionColor: Vec = ZERO


@dataclass(eq=False)
class PhongMaterial:
    Alpha: float = 0.0
    SpecularColor: Vec = ZERO
    DiffuseColor: Vec = ZERO
    EmissionColor: Vec = ZERO
    AmbientColor: Vec = ZERO
    NoFluxCorrection: bool = False


@dataclass(eq=False)
class RefractMaterial:
    IndexOfRefraction: float = 1.0
    RefractColor: Vec = ZERO
    SpecularColor: Vec = ZERO


@dataclass(eq=False)
class JoinedMaterial:
    Materials: List[object] = field(default_factory=list)
    Probs: List[float] = field(default_factory=list)


@dataclass(eq=False)
class CheckerLambertMaterial:
    """Declarative form of showcase's FloorObject (examples/renderings/showcase/room.go:61-75):
    Lambert whose diffuse colour is Color2 where int(mod(x+300,2)) == int(mod(y+301,2)),
    else Color1."""
    Color1: Vec = ZERO
    Color2: Vec = ZERO


@dataclass(eq=False)
class ZGradientPhongMaterial:
    """Declarative form of showcase's VaseObject (models.go:79-97): Phong whose diffuse colour
    is Color1*frac + Color2*(1-frac), frac = z / MaxZ."""
    Alpha: float = 0.0
    SpecularColor: Vec = ZERO
    Color1: Vec = ZERO
    Color2: Vec = ZERO
    MaxZ: float = 1.0


def material_desc(m, index_of):
    """render3d.Material -> m3d_material_desc (raises UnsupportedError)."""
    d = N.MaterialDesc()

    def put(dst, c):
        dst[:] = [float(x) for x in c]

    if isinstance(m, LambertMaterial):
        d.kind = N.MAT_LAMBERT
        put(d.diffuse, m.DiffuseColor)
        put(d.ambient, m.AmbientColor)
        put(d.emission, m.EmissionColor)
    elif isinstance(m, PhongMaterial):
        d.kind = N.MAT_PHONG
        d.alpha = m.Alpha
        put(d.specular, m.SpecularColor)
        put(d.diffuse, m.DiffuseColor)
        put(d.emission, m.EmissionColor)
        put(d.ambient, m.AmbientColor)
        if m.NoFluxCorrection:
            d.flags |= N.MAT_NO_FLUX_CORRECTION
    elif isinstance(m, RefractMaterial):
        d.kind = N.MAT_REFRACT
        d.index_of_refraction = m.IndexOfRefraction
        put(d.refract, m.RefractColor)
        put(d.specular, m.SpecularColor)
    elif isinstance(m, JoinedMaterial):
        if len(m.Probs) != len(m.Materials):
            raise ValueError("mismatched probabilities and materials")  # material.go:573-575
        if len(m.Materials) > 4:
            raise UnsupportedError("JoinedMaterial with more than 4 parts")
        d.kind = N.MAT_JOINED
        d.num_sub = len(m.Materials)
        for i, sub in enumerate(m.Materials):
            d.sub[i] = index_of(sub)
            d.sub_prob[i] = float(m.Probs[i])
    elif isinstance(m, CheckerLambertMaterial):
        d.kind = N.MAT_LAMBERT
        d.flags |= N.MAT_CHECKER
        put(d.diffuse, m.Color1)
        put(d.diffuse2, m.Color2)
    elif isinstance(m, ZGradientPhongMaterial):
        d.kind = N.MAT_PHONG
        d.flags |= N.MAT_Z_GRADIENT
        d.alpha = m.Alpha
        put(d.specular, m.SpecularColor)
        put(d.diffuse, m.Color1)
        put(d.diffuse2, m.Color2)
        d.proc_param = m.MaxZ
    else:
        raise UnsupportedError("material type %s is not supported on the GPU path" % type(m).__name__)
    return d


# ---- objects (object.go, transform.go) ----------------------------------------------------
@dataclass(eq=False)
class ColliderObject:
    Collider: object = None
    Material: object = None
    FlipNormals: bool = False  # declarative form of showcase's DomeObject (room.go:40-44)


class JoinedObject(list):
    """render3d.JoinedObject ([]Object)."""


@dataclass(eq=False)
class _Transformed:
    Object: object
    Matrix: Optional[Sequence[float]]  # row-major 3x3 (model3d/matrix.go:11-12)
    Offset: Vec


def Translate(obj, offset):
    """render3d.Translate (transform.go:6-31)."""
    return _Transformed(obj, None, tuple(float(x) for x in offset))


def MatrixMultiply(obj, m):
    """render3d.MatrixMultiply (transform.go:48-85); m is row-major 3x3."""
    return _Transformed(obj, [float(x) for x in np.asarray(m, np.float64).reshape(9)], (0.0, 0.0, 0.0))


def _compose(outer, inner):
    """x -> outer(inner(x)) for (matrix|None, offset) pairs."""
    if outer is None:
        return inner
    if inner is None:
        return outer
    mo = np.eye(3) if outer[0] is None else np.asarray(outer[0], np.float64).reshape(3, 3)
    mi = np.eye(3) if inner[0] is None else np.asarray(inner[0], np.float64).reshape(3, 3)
    return ((mo @ mi).reshape(9).tolist(), tuple((mo @ np.asarray(inner[1]) + np.asarray(outer[1])).tolist()))


class Scene:
    """A render3d.Object compiled to a device scene (m3d_scene)."""

    def __init__(self, obj, ctx=None, device_lbvh=False, device_build=False, record=None):
        """record: optional list that receives every builder call as (name, args...) tuples with the
        ctypes structs as bytes -- tests/test_c_abi.py replays them from a C program."""
        self.ctx = ctx or N.default_context()
        self._record = record
        L = N.lib()
        b = C.c_void_p()
        N.check(L.m3d_scene_builder_create(self.ctx.h, C.byref(b)))
        self.materials = []      # python material objects, index == device index
        self._mat_index = {}
        self.objects = []        # leaf objects in device order
        # A device-resident MeshCollider that several objects of the tree share (the reference's
        # golf_balls example translates one collider many times, golf_balls/main.go:25-39) is
        # INSTANCED: the scene keeps one copy of its triangles and hierarchy (m3d_scene_add_instance).
        # A collider used once is merged into the scene's world-space BVH as before.
        self._uses = {}
        self._instanced = []     # keeps the instanced colliders alive as long as the scene
        self._count_uses(obj)
        try:
            self._add(b, obj, None)
            self.h = C.c_void_p()
            N.check(L.m3d_scene_build(b, C.c_uint32(N.MESH_BUILD_DEVICE_COLLAPSE if device_build else (N.MESH_BUILD_DEVICE_LBVH if device_lbvh else 0)), C.byref(self.h)))
        finally:
            L.m3d_scene_builder_destroy(b)

    def _count_uses(self, obj):
        if isinstance(obj, (list, tuple)):
            for o in obj:
                self._count_uses(o)
        elif isinstance(obj, _Transformed):
            self._count_uses(obj.Object)
        elif isinstance(obj, ColliderObject) and isinstance(obj.Collider, MeshCollider):
            self._uses[id(obj.Collider)] = self._uses.get(id(obj.Collider), 0) + 1

    def _material(self, b, m):
        if id(m) in self._mat_index:
            return self._mat_index[id(m)]
        d = material_desc(m, lambda sub: self._material(b, sub))
        idx = C.c_int32(-1)
        N.check(N.lib().m3d_scene_add_material(b, C.byref(d), C.byref(idx)))
        if self._record is not None:
            self._record.append(("material", bytes(d)))
        self._mat_index[id(m)] = idx.value
        self.materials.append(m)
        assert idx.value == len(self.materials) - 1
        return idx.value

    def _add(self, b, obj, xf):
        L = N.lib()
        if isinstance(obj, (list, tuple)):
            for o in obj:
                self._add(b, o, xf)
            return
        if isinstance(obj, _Transformed):
            self._add(b, obj.Object, _compose(xf, (obj.Matrix, obj.Offset)))
            return
        if not isinstance(obj, ColliderObject):
            raise UnsupportedError("object type %s is not supported on the GPU path" % type(obj).__name__)
        mat = self._material(b, obj.Material)
        flags = N.OBJ_FLIP_NORMAL if obj.FlipNormals else 0
        t = None
        if xf is not None:
            t = N.Transform()
            t.matrix[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1] if xf[0] is None else xf[0]
            t.offset[:] = list(xf[1])
        tp = C.byref(t) if t is not None else None
        c = obj.Collider
        idx = C.c_int32(-1)
        if isinstance(c, Sphere):
            N.check(L.m3d_scene_add_sphere(b, _d3(c.Center), C.c_double(c.Radius), C.c_int32(mat),
                                           C.c_uint32(flags), tp, C.byref(idx)))
        elif isinstance(c, Rect):
            N.check(L.m3d_scene_add_rect(b, _d3(c.MinVal), _d3(c.MaxVal), C.c_int32(mat),
                                         C.c_uint32(flags), tp, C.byref(idx)))
        elif isinstance(c, Cylinder):
            N.check(L.m3d_scene_add_cylinder(b, _d3(c.P1), _d3(c.P2), C.c_double(c.Radius), C.c_int32(mat),
                                             C.c_uint32(flags), tp, C.byref(idx)))
        elif (isinstance(c, MeshCollider) and self._uses.get(id(c), 0) > 1 and c.ctx is self.ctx
              and c.num_triangles > 0):
            N.check(L.m3d_scene_add_instance(b, c.h, C.c_int32(mat), C.c_uint32(flags), tp, C.byref(idx)))
            self._instanced.append(c)
            if self._record is not None:
                raise UnsupportedError("recording does not cover instanced colliders")
            self.objects.append(obj)
            return
        elif isinstance(c, MeshCollider) or isinstance(c, np.ndarray):
            tris = c.triangles if isinstance(c, MeshCollider) else c
            vn = c.vertex_normals if isinstance(c, MeshCollider) else None
            tris = np.ascontiguousarray(np.asarray(tris, np.float32).reshape(-1, 9))
            vn = None if vn is None else np.ascontiguousarray(np.asarray(vn, np.float32).reshape(-1, 9))
            N.check(L.m3d_scene_add_mesh(b, _p(tris, f32p), C.c_int64(tris.shape[0]), _p(vn, f32p),
                                         C.c_int32(mat), C.c_uint32(flags), tp, C.byref(idx)))
        else:
            raise UnsupportedError("collider type %s is not supported on the GPU path" % type(c).__name__)
        if self._record is not None:
            xb = bytes(t) if t is not None else None
            if isinstance(c, Sphere):
                self._record.append(("sphere", mat, flags, xb, tuple(c.Center), float(c.Radius)))
            elif isinstance(c, Rect):
                self._record.append(("rect", mat, flags, xb, tuple(c.MinVal), tuple(c.MaxVal)))
            elif isinstance(c, Cylinder):
                self._record.append(("cylinder", mat, flags, xb, tuple(c.P1), tuple(c.P2), float(c.Radius)))
            else:
                self._record.append(("mesh", mat, flags, xb, tris.copy(), None if vn is None else vn.copy()))
        self.objects.append(obj)

    def close(self):
        if getattr(self, "h", None):
            N.lib().m3d_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Min(self):
        return self._bounds()[0]

    def Max(self):
        return self._bounds()[1]

    def Info(self):
        """m3d_scene_get_info: the scene's own merged triangle hierarchy (instanced colliders keep
        their triangles in their MeshCollider and are not counted)."""
        info = N.MeshInfo()
        N.check(N.lib().m3d_scene_get_info(self.h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in info._fields_}

    def _bounds(self):
        mn, mx = (C.c_double * 3)(), (C.c_double * 3)()
        N.check(N.lib().m3d_scene_bounds(self.h, mn, mx))
        return tuple(mn), tuple(mx)

    def Cast(self, origins, directions, counters=False):
        """Batched Object.Cast (object.go:141-153): dict(t, obj, prim, normal)."""
        org = np.ascontiguousarray(np.asarray(origins, np.float32).reshape(-1, 3))
        dr = np.ascontiguousarray(np.asarray(directions, np.float32).reshape(-1, 3))
        n = org.shape[0]
        t = np.zeros(n, np.float32)
        obj = np.full(n, -1, np.int32)
        prim = np.full(n, -1, np.int32)
        normal = np.zeros((n, 3), np.float32)
        stats = N.Stats()
        N.check(N.lib().m3d_scene_cast(self.h, _p(org, f32p), _p(dr, f32p), C.c_int64(n), _p(t, f32p),
                                       _p(obj, i32p), _p(prim, i32p), _p(normal, f32p),
                                       C.c_uint32(N.TRACE_COUNTERS if counters else 0), C.byref(stats)))
        return dict(t=t, obj=obj, prim=prim, normal=normal,
                    stats={k: getattr(stats, k) for k, _ in stats._fields_})


def _as_scene(obj, ctx=None):
    """A Scene is used as is; any other object tree is compiled for this call only, like the
    reference re-reads the object on every Render (callers that render the same objects many
    times build a Scene once and pass it: no hidden cache that could go stale when the object
    tree, a material or a mesh is edited between renders)."""
    if isinstance(obj, Scene):
        return obj
    return Scene(obj, ctx)


# ---- camera (camera.go) --------------------------------------------------------------------
DefaultFieldOfView = math.pi / 2


@dataclass
class Camera:
    Origin: Vec
    ScreenX: Vec
    ScreenY: Vec
    FieldOfView: float

    def _c(self):
        c = N.Camera()
        c.origin[:] = list(self.Origin)
        c.screen_x[:] = list(self.ScreenX)
        c.screen_y[:] = list(self.ScreenY)
        c.field_of_view = self.FieldOfView
        return c


def _camera_axes(cam, w, h):
    """Camera.axes (camera.go:100-113)."""
    plane = 1.0 / math.tan(cam.FieldOfView / 2)
    x, y = np.asarray(cam.ScreenX, np.float64), np.asarray(cam.ScreenY, np.float64)
    z = np.cross(x, y)
    z = z * (1 / math.sqrt(float(z @ z)))
    if w > h:
        y = y * (h / w)
    else:
        x = x * (w / h)
    return x, y, z * plane


def CasterRays(cam, imageWidth, imageHeight):
    """Camera.Caster (camera.go:68-82) for every pixel: float64 directions [H*W, 3], row-major
    (idx = x + y*W), built from imageWidth-1 / imageHeight-1 like the renderers' callers do
    (raycast.go:16-18)."""
    w, h = float(imageWidth - 1), float(imageHeight - 1)
    x, y, z = _camera_axes(cam, w, h)
    cx, cy = w / 2, h / 2
    fx = (np.arange(imageWidth, dtype=np.float64) - cx) / cx
    fy = (np.arange(imageHeight, dtype=np.float64) - cy) / cy
    d = fx[None, :, None] * x[None, None, :] + fy[:, None, None] * y[None, None, :] + z[None, None, :]
    return d.reshape(-1, 3)


def Uncaster(cam, imageWidth, imageHeight):
    """Camera.Uncaster (camera.go:84-98): spatial -> screen coordinates."""
    x, y, z = _camera_axes(cam, imageWidth, imageHeight)
    inv = np.linalg.inv(np.stack([x, y, z], axis=1))
    cx, cy = imageWidth / 2, imageHeight / 2
    org = np.asarray(cam.Origin, np.float64)

    def f(coord):
        v = inv @ (np.asarray(coord, np.float64) - org)
        return v[0] / v[2] * cx + cx, v[1] / v[2] * cy + cy
    return f


def NewCameraAt(source, dest, fov=0.0):
    """render3d.NewCameraAt (camera.go:48-66), float64 host arithmetic."""
    if fov == 0:
        fov = DefaultFieldOfView
    s, d = np.asarray(source, np.float64), np.asarray(dest, np.float64)
    z = d - s
    z = z * (1 / math.sqrt(float(z @ z)))
    x = np.array([z[1], -z[0], 0.0])
    if math.sqrt(float(x @ x)) < 1e-5:
        ex = np.array([1.0, 0.0, 0.0])
        x = ex - z * float(z @ ex)  # X(1).ProjectOut(zAxis); z is unit
    x = x * (1 / math.sqrt(float(x @ x)))
    y = np.cross(z, x)
    return Camera(tuple(s.tolist()), tuple(x.tolist()), tuple(y.tolist()), float(fov))


@dataclass
class PointLight:
    Origin: Vec
    Color: Vec
    QuadDropoff: bool = False

    def _c(self):
        l = N.PointLight()
        l.origin[:] = list(self.Origin)
        l.color[:] = list(self.Color)
        l.quad_dropoff = 1 if self.QuadDropoff else 0
        return l


def _image_zeros(shape):
    """Zero-filled float32 pixels in page-locked memory (m3d_host_alloc) when the library and a
    device are there: Render copies the image to the device and back every frame, which is 5x
    faster from pinned memory.  Plain numpy memory otherwise (an Image is also a host-side container
    for the PNG / GIF helpers, which need no device)."""
    try:
        a = N.host_empty(shape, np.float32)
    except (N.M3DError, ImportError, OSError):
        return np.zeros(shape, np.float32)
    a[...] = 0
    return a


class Image:
    """render3d.Image (image.go:17-47): Data is [Height, Width, 3] linear RGB."""

    def __init__(self, width, height):
        self.Width, self.Height = int(width), int(height)
        self.Data = _image_zeros((self.Height, self.Width, 3))

    def RGBA8(self):
        """8-bit sRGB like Image.RGBA (image.go:125-145, light.go:41-47)."""
        c = np.clip(self.Data.astype(np.float64), 0.0, 1.0)
        s = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055)
        return (s * (256.0 - 0.001)).astype(np.uint8)

    def SetAll(self, c):
        """Image.SetAll (image.go:48-53)."""
        self.Data[...] = np.asarray(c, np.float32)

    def CopyFrom(self, i1, x, y):
        """Image.CopyFrom (image.go:55-69): copy i1 into this image at (x, y), clipped."""
        w = min(i1.Width, self.Width - x)
        h = min(i1.Height, self.Height - y)
        if w > 0 and h > 0:
            self.Data[y:y + h, x:x + w] = i1.Data[:h, :w]

    def FillRange(self):
        """Image.FillRange (image.go:71-86)."""
        m = float(self.Data.max()) if self.Data.size else 0.0
        if m > 0:
            self.Data *= np.float32(1.0 / m)

    def Scale(self, s):
        self.Data *= np.float32(s)

    def Downsample(self, factor):
        """Image.Downsample (image.go:95-120): box filter in linear RGB."""
        if self.Width % factor or self.Height % factor:
            raise ValueError("image size %d x %d cannot be divided evenly by factor %d"
                             % (self.Width, self.Height, factor))
        out = Image(self.Width // factor, self.Height // factor)
        d = self.Data.astype(np.float64).reshape(out.Height, factor, out.Width, factor, 3)
        out.Data = (d.sum(axis=(1, 3)) * (1.0 / (factor * factor))).astype(np.float32)
        return out

    def Gray8(self):
        """Image.Gray (image.go:147-172): 8-bit sRGB, then Go's color.GrayModel luma
        (19595 R + 38470 G + 7471 B + 2^15) >> 24 on 16-bit channels."""
        rgb = self.RGBA8().astype(np.uint32) * 0x101
        y = (19595 * rgb[..., 0] + 38470 * rgb[..., 1] + 7471 * rgb[..., 2] + (1 << 15)) >> 24
        return y.astype(np.uint8)

    def Save(self, path):
        """Image.Save (image.go:174-199): PNG only here (the extension decides; JPEG is not
        implemented on this side)."""
        ext = path.lower().rsplit(".", 1)[-1] if "." in path else ""
        if ext != "png":
            raise ValueError("save image: unknown extension '.%s' (only .png is supported here)" % ext)
        from .helpers import write_png
        write_png(path, self.RGBA8())


def _lights(lights):
    arr = (N.PointLight * max(1, len(lights)))(*[l._c() for l in lights])
    return arr


@dataclass
class RayCaster:
    """render3d.RayCaster (raycast.go:9-39)."""
    Camera: Camera = None
    Lights: List[PointLight] = field(default_factory=list)

    def Render(self, img: Image, obj, partition=None):
        sc = _as_scene(obj)
        cam = self.Camera._c()
        data = np.ascontiguousarray(img.Data, np.float32)
        stats = N.Stats()
        part = None
        if partition is not None:
            part = N.Partition(int(partition[0]), int(partition[1]), 0)
        N.check(N.lib().m3d_render_raycast(sc.h, C.byref(cam), _lights(self.Lights), C.c_int32(len(self.Lights)),
                                           C.c_int32(img.Width), C.c_int32(img.Height),
                                           C.byref(part) if part is not None else None,
                                           _p(data, f32p), C.byref(stats)))
        img.Data = data
        return {k: getattr(stats, k) for k, _ in stats._fields_}

    @staticmethod
    def RenderViews(casters, width, height, obj, downsample=1):
        """Many RayCaster frames of one object in one library call (m3d_render_raycast_views): view
        v is rendered like casters[v].Render into a fresh (black) width x height image and, with
        downsample > 1, box-filtered like Image.Downsample.  Returns (float32 [V, H/f, W/f, 3], stats).
        What SaveRandomGrid / SaveRotatingGIF render view by view in the reference
        (helpers.go:133-236); on a multi-device context the views are spread over the GPUs."""
        sc = _as_scene(obj)
        nv = len(casters)
        cams = (N.Camera * max(1, nv))(*[c.Camera._c() for c in casters])
        all_lights = [l for c in casters for l in c.Lights]
        begin = np.zeros(nv + 1, np.int32)
        for v, c in enumerate(casters):
            begin[v + 1] = begin[v] + len(c.Lights)
        lights = _lights(all_lights)
        out = N.host_empty((nv, height // downsample, width // downsample, 3), np.float32) if nv else \
            np.zeros((0, height // downsample, width // downsample, 3), np.float32)
        stats = N.Stats()
        N.check(N.lib().m3d_render_raycast_views(sc.h, cams, C.c_int32(nv), lights, _p(begin, i32p), C.c_int32(width),
                                                 C.c_int32(height), C.c_int32(downsample), _p(out, f32p),
                                                 C.byref(stats)))
        return out, {k: getattr(stats, k) for k, _ in stats._fields_}

    def RenderDevice(self, width, height, obj, d_rgb, partition=None, stream=0):
        """Render into a device buffer (pointer to W*H*3 float32; pixels whose ray misses keep
        their value, raycast.go:26-28): what a multi-GPU driver gathers by row band."""
        sc = _as_scene(obj)
        cam = self.Camera._c()
        stats = N.Stats()
        part = None
        if partition is not None:
            part = N.Partition(int(partition[0]), int(partition[1]), 0)
        N.check(N.lib().m3d_render_raycast_device(
            sc.h, C.byref(cam), _lights(self.Lights), C.c_int32(len(self.Lights)), C.c_int32(width),
            C.c_int32(height), C.byref(part) if part is not None else None, C.c_void_p(d_rgb),
            C.c_void_p(stream or None), C.byref(stats)))
        return {k: getattr(stats, k) for k, _ in stats._fields_}


# ---- focus points (focus_point.go) ----------------------------------------------------------
@dataclass(eq=False)
class PhongFocusPoint:
    """render3d.PhongFocusPoint (focus_point.go:30-71)."""
    Target: Vec = (0.0, 0.0, 0.0)
    Alpha: float = 0.0
    MaterialFilter: Optional[Callable[[object], bool]] = None


@dataclass(eq=False)
class SphereFocusPoint:
    """render3d.SphereFocusPoint (focus_point.go:73-153)."""
    Center: Vec = (0.0, 0.0, 0.0)
    Radius: float = 0.0
    MaterialFilter: Optional[Callable[[object], bool]] = None


def _focus_desc(fp, prob, materials):
    """FocusPoint -> m3d_focus_point.  MaterialFilter closures cannot cross the C ABI: the
    filter is evaluated once per scene material into a bit mask (include/m3d.h)."""
    d = N.FocusPoint()
    if isinstance(fp, PhongFocusPoint):
        d.kind = N.FOCUS_PHONG
        d.target[:] = [float(x) for x in fp.Target]
        d.alpha = float(fp.Alpha)
    elif isinstance(fp, SphereFocusPoint):
        d.kind = N.FOCUS_SPHERE
        d.target[:] = [float(x) for x in fp.Center]
        d.radius = float(fp.Radius)
    else:
        raise UnsupportedError("focus point type %s is not supported on the GPU path" % type(fp).__name__)
    if len(materials) > 64:
        raise UnsupportedError("focus-point material filters support at most 64 scene materials")
    mask = 0
    for i, m in enumerate(materials):
        if fp.MaterialFilter is None or fp.MaterialFilter(m):
            mask |= 1 << i
    d.material_mask = mask
    d.prob = float(prob)
    return d


def _samples_partition(partition):
    """(row_begin, row_end, sample_begin) or None -> m3d_partition pointer (or None)."""
    if partition is None:
        return None
    flags = int(partition[3]) if len(partition) > 3 else 0
    return N.Partition(int(partition[0]), int(partition[1]), int(partition[2]), flags, 0)


@dataclass
class RecursiveRayTracer:
    """render3d.RecursiveRayTracer (raytrace.go:14-119): same exported fields.

    Render() runs the wavefront path tracer of libm3dgpu (m3d_render_path).  Early stopping
    with MinSamples/MaxStddev/OversaturatedStddevs follows the reference's per-sample rule
    (ray_renderer.go:128-148); a Convergence callback cannot cross the ABI and raises
    UnsupportedError.  Adaptive renders cannot be sharded by sample index (shard by rows)."""
    Camera: Camera = None
    Lights: List[PointLight] = field(default_factory=list)
    FocusPoints: List[object] = field(default_factory=list)
    FocusPointProbs: List[float] = field(default_factory=list)
    MaxDepth: int = 0
    NumSamples: int = 0
    MinSamples: int = 0
    MaxStddev: float = 0.0
    OversaturatedStddevs: float = 0.0
    Convergence: Optional[Callable] = None
    Cutoff: float = 0.0
    Antialias: float = 0.0
    Epsilon: float = 0.0
    LogFunc: Optional[Callable[[float, float], None]] = None
    Seed: int = 0  # Philox key (the reference seeds math/rand from the global source)

    def _params(self, sc, num_samples):
        if self.NumSamples == 0 and num_samples == 0:
            raise ValueError("must set NumSamples to non-zero for rayRenderer")  # ray_renderer.go:26-28
        if len(self.FocusPoints) != len(self.FocusPointProbs):
            raise ValueError("FocusPoints and FocusPointProbs must match in length")  # raytrace.go:186-188
        if self.Convergence is not None:
            raise UnsupportedError("Convergence callbacks cannot run on the GPU path (MinSamples/MaxStddev/"
                                   "OversaturatedStddevs are supported)")
        if len(self.FocusPoints) > 4:
            raise UnsupportedError("at most 4 focus points are supported on the GPU path")
        p = N.PathParams()
        p.max_depth = int(self.MaxDepth)
        p.num_samples = int(num_samples or self.NumSamples)
        p.min_samples = int(self.MinSamples)
        p.max_stddev = float(self.MaxStddev)
        p.oversaturated_stddevs = float(self.OversaturatedStddevs)
        p.cutoff = float(self.Cutoff)
        p.antialias = float(self.Antialias)
        p.epsilon = float(self.Epsilon)
        p.num_focus_points = len(self.FocusPoints)
        for i, (fp, prob) in enumerate(zip(self.FocusPoints, self.FocusPointProbs)):
            p.focus[i] = _focus_desc(fp, prob, sc.materials)
        p.seed = int(self.Seed)
        return p

    def RenderSums(self, width, height, obj, partition=None, sample_count=None, variance=False,
                   antialias=None):
        """Per-pixel SUMS (and sums of squares) over `sample_count` samples of this shard:
        the quantity that adds up across GPUs.  partition = (row_begin, row_end,
        sample_begin)."""
        sc = _as_scene(obj)
        n = int(self.NumSamples if sample_count is None else sample_count)
        p = self._params(sc, n)
        if antialias is not None:
            p.antialias = float(antialias)
        if variance or sample_count is not None or partition is not None and partition[2] != 0:
            p.min_samples = 0  # fixed-count shard / estimateVariance: no early stop
        cam = self.Camera._c()
        rgb = np.zeros((height, width, 3), np.float32)
        sq = np.zeros((height, width, 3), np.float32) if variance else None
        stats = N.Stats()
        part = _samples_partition(partition)
        N.check(N.lib().m3d_render_path(sc.h, C.byref(cam), _lights(self.Lights), C.c_int32(len(self.Lights)),
                                        C.byref(p), C.c_int32(width), C.c_int32(height),
                                        C.byref(part) if part is not None else None, C.c_int32(n),
                                        _p(rgb, f32p), _p(sq, f32p), C.byref(stats)))
        return rgb, sq, {k: getattr(stats, k) for k, _ in stats._fields_}

    def RenderSumsDevice(self, width, height, obj, d_rgb_sum, d_rgb_sumsq=0, partition=None,
                         sample_count=None, stream=0):
        """Like RenderSums on device buffers (pointers to W*H*3 float32 accumulators that the
        call ADDS into); what a multi-GPU driver reduces with NCCL."""
        sc = _as_scene(obj)
        n = int(self.NumSamples if sample_count is None else sample_count)
        p = self._params(sc, n)
        cam = self.Camera._c()
        stats = N.Stats()
        part = _samples_partition(partition)
        N.check(N.lib().m3d_render_path_device(
            sc.h, C.byref(cam), _lights(self.Lights), C.c_int32(len(self.Lights)), C.byref(p),
            C.c_int32(width), C.c_int32(height), C.byref(part) if part is not None else None, C.c_int32(n),
            C.c_void_p(d_rgb_sum), C.c_void_p(d_rgb_sumsq or None), C.c_void_p(stream or None), C.byref(stats)))
        return {k: getattr(stats, k) for k, _ in stats._fields_}

    def Render(self, img: Image, obj):
        """(*RecursiveRayTracer).Render (raytrace.go:98-100)."""
        rgb, _, stats = self.RenderSums(img.Width, img.Height, obj)
        img.Data = rgb / np.float32(self.NumSamples)  # colorSum.Scale(1/numSamples) ray_renderer.go:150
        if self.LogFunc is not None:
            self.LogFunc(1.0, float(self.NumSamples))
        return stats

    def RenderVariance(self, img: Image, obj, numSamples, antialias=None):
        """rayRenderer.RenderVariance (ray_renderer.go:59-67,90-110): per-pixel sample variance
        with Bessel's correction, clamped at zero."""
        if numSamples < 2:
            raise ValueError("need to take at least two samples")
        rgb, sq, stats = self.RenderSums(img.Width, img.Height, obj, sample_count=numSamples, variance=True,
                                         antialias=antialias)
        n = float(numSamples)
        mean = rgb.astype(np.float64) / n
        var = (sq.astype(np.float64) / n - mean * mean) * (n / (n - 1))
        img.Data = np.maximum(var, 0.0).astype(np.float32)
        return stats

    def RayVariance(self, obj, width, height, samples):
        """rayRenderer.RayVariance (ray_renderer.go:69-88): mean variance, no antialiasing."""
        if samples < 2:
            raise ValueError("need to take at least two samples")
        img = Image(width, height)
        self.RenderVariance(img, obj, samples, antialias=0.0)
        return float(img.Data.astype(np.float64).sum() / (3 * width * height))


# ---- area lights (light.go:104-314) -----------------------------------------------------------
class AreaLight(ColliderObject):
    """render3d.AreaLight: an Object that can also be sampled as an emitter.  Created by
    NewSphereAreaLight / NewMeshAreaLight; its material is Lambert with only an emission
    colour, like the reference (light.go:131-140, 237-252)."""

    def __init__(self, collider, emission):
        super().__init__(Collider=collider, Material=LambertMaterial(EmissionColor=tuple(emission)))
        self.Emission = tuple(float(x) for x in emission)


def NewSphereAreaLight(sphere: Sphere, emission):
    """render3d.NewSphereAreaLight (light.go:131-140)."""
    return AreaLight(sphere, emission)


def NewMeshAreaLight(mesh, emission):
    """render3d.NewMeshAreaLight (light.go:237-252); mesh: triangle array [n,3,3] or MeshCollider."""
    return AreaLight(mesh, emission)


class JoinedAreaLight(JoinedObject):
    """render3d.JoinAreaLights (light.go:283-301): the lights are also scene objects."""


def JoinAreaLights(*lights):
    for l in lights:
        if not isinstance(l, AreaLight):
            raise UnsupportedError("area light type %s is not supported on the GPU path" % type(l).__name__)
    return JoinedAreaLight(lights)


def _area_lights(sc, light):
    """AreaLight | JoinedAreaLight -> m3d_area_light[] (scene object indices by identity)."""
    leaves = list(light) if isinstance(light, (list, tuple)) else [light]
    arr = (N.AreaLight * max(1, len(leaves)))()
    for i, l in enumerate(leaves):
        if not isinstance(l, AreaLight):
            raise UnsupportedError("area light type %s is not supported on the GPU path" % type(l).__name__)
        idx = [k for k, o in enumerate(sc.objects) if o is l]
        if not idx:
            raise ValueError("the BidirPathTracer's Light must be part of the rendered scene")
        arr[i].object = idx[0]
        arr[i].emission[:] = list(l.Emission)
    return arr, len(leaves)


@dataclass
class BidirPathTracer:
    """render3d.BidirPathTracer (bidir.go:14-84): same exported fields; Render() runs
    libm3dgpu's wavefront bidirectional path tracer (m3d_render_bidir)."""
    Camera: Camera = None
    Light: object = None
    MaxDepth: int = 0
    MaxLightDepth: int = 0
    MinDepth: int = 0
    RouletteDelta: float = 0.0
    PowerHeuristic: float = 0.0
    NumSamples: int = 0
    MinSamples: int = 0
    MaxStddev: float = 0.0
    OversaturatedStddevs: float = 0.0
    Convergence: Optional[Callable] = None
    Cutoff: float = 0.0
    Antialias: float = 0.0
    Epsilon: float = 0.0
    LogFunc: Optional[Callable[[float, float], None]] = None
    Seed: int = 0

    def _params(self, num_samples):
        if self.NumSamples == 0 and num_samples == 0:
            raise ValueError("must set NumSamples to non-zero for rayRenderer")
        if self.Convergence is not None:
            raise UnsupportedError("Convergence callbacks cannot run on the GPU path (MinSamples/MaxStddev/"
                                   "OversaturatedStddevs are supported)")
        p = N.BidirParams()
        p.max_depth, p.max_light_depth, p.min_depth = int(self.MaxDepth), int(self.MaxLightDepth), int(self.MinDepth)
        p.num_samples = int(num_samples or self.NumSamples)
        p.roulette_delta, p.power_heuristic = float(self.RouletteDelta), float(self.PowerHeuristic)
        p.cutoff, p.antialias, p.epsilon = float(self.Cutoff), float(self.Antialias), float(self.Epsilon)
        p.min_samples, p.max_stddev = int(self.MinSamples), float(self.MaxStddev)
        p.oversaturated_stddevs = float(self.OversaturatedStddevs)
        p.seed = int(self.Seed)
        return p

    def RenderSums(self, width, height, obj, partition=None, sample_count=None, variance=False, antialias=None):
        sc = _as_scene(obj)
        n = int(self.NumSamples if sample_count is None else sample_count)
        p = self._params(n)
        if antialias is not None:
            p.antialias = float(antialias)
        if variance or sample_count is not None or partition is not None and partition[2] != 0:
            p.min_samples = 0  # fixed-count shard / estimateVariance: no early stop
        cam = self.Camera._c()
        lights, nl = _area_lights(sc, self.Light)
        rgb = np.zeros((height, width, 3), np.float32)
        sq = np.zeros((height, width, 3), np.float32) if variance else None
        stats = N.Stats()
        part = _samples_partition(partition)
        N.check(N.lib().m3d_render_bidir(sc.h, C.byref(cam), lights, C.c_int32(nl), C.byref(p), C.c_int32(width),
                                         C.c_int32(height), C.byref(part) if part is not None else None,
                                         C.c_int32(n), _p(rgb, f32p), _p(sq, f32p), C.byref(stats)))
        return rgb, sq, {k: getattr(stats, k) for k, _ in stats._fields_}

    def RenderSumsDevice(self, width, height, obj, d_rgb_sum, d_rgb_sumsq=0, partition=None, sample_count=None,
                         stream=0):
        sc = _as_scene(obj)
        n = int(self.NumSamples if sample_count is None else sample_count)
        p = self._params(n)
        cam = self.Camera._c()
        lights, nl = _area_lights(sc, self.Light)
        stats = N.Stats()
        part = _samples_partition(partition)
        N.check(N.lib().m3d_render_bidir_device(
            sc.h, C.byref(cam), lights, C.c_int32(nl), C.byref(p), C.c_int32(width), C.c_int32(height),
            C.byref(part) if part is not None else None, C.c_int32(n), C.c_void_p(d_rgb_sum),
            C.c_void_p(d_rgb_sumsq or None), C.c_void_p(stream or None), C.byref(stats)))
        return {k: getattr(stats, k) for k, _ in stats._fields_}

    def Render(self, img: Image, obj):
        """(*BidirPathTracer).Render (bidir.go:66-68)."""
        rgb, _, stats = self.RenderSums(img.Width, img.Height, obj)
        img.Data = rgb / np.float32(self.NumSamples)
        if self.LogFunc is not None:
            self.LogFunc(1.0, float(self.NumSamples))
        return stats

    def RenderVariance(self, img: Image, obj, numSamples, antialias=None):
        """rayRenderer.RenderVariance (ray_renderer.go:59-67,90-110)."""
        if numSamples < 2:
            raise ValueError("need to take at least two samples")
        rgb, sq, stats = self.RenderSums(img.Width, img.Height, obj, sample_count=numSamples, variance=True,
                                         antialias=antialias)
        n = float(numSamples)
        mean = rgb.astype(np.float64) / n
        img.Data = np.maximum((sq.astype(np.float64) / n - mean * mean) * (n / (n - 1)), 0.0).astype(np.float32)
        return stats

    def RayVariance(self, obj, width, height, samples):
        """rayRenderer.RayVariance (ray_renderer.go:69-88)."""
        img = Image(width, height)
        self.RenderVariance(img, obj, samples, antialias=0.0)
        return float(img.Data.astype(np.float64).sum() / (3 * width * height))
