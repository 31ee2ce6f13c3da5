"""Host-side mirror of the reference's collision interface for the GPU path.

Names follow model3d (Go): ``Ray`` / ``RayCollision`` / ``TriangleCollision``
(model3d/collisions.go:12-46) and the ``Collider`` interface
(collisions.go:52-72).  ``MeshCollider`` is what ``MeshToCollider(mesh)``
(collisions.go:138-142) returns here: a device-resident wide BVH queried through
libm3dgpu's C ABI.  Unsupported queries raise; nothing falls back to the CPU.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import _native as N

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


@dataclass
class Ray:
    """model3d.Ray (collisions.go:12-15). Direction is not normalised."""
    Origin: Tuple[float, float, float]
    Direction: Tuple[float, float, float]


@dataclass
class TriangleCollision:
    """model3d.TriangleCollision (collisions.go:39-46); Triangle is the id (index in
    the caller's triangle array) instead of a pointer."""
    Triangle: int
    Barycentric: Tuple[float, float, float]


@dataclass
class RayCollision:
    """model3d.RayCollision (collisions.go:19-35)."""
    Scale: float
    Normal: Tuple[float, float, float]
    Extra: Optional[TriangleCollision] = None


@dataclass
class BatchCollisions:
    """Result of MeshCollider.FirstRayCollisions: structure-of-arrays RayCollision."""
    Collides: np.ndarray      # bool  [n]
    Scale: np.ndarray         # f32   [n]
    Normal: np.ndarray        # f32   [n,3]
    Triangle: np.ndarray      # i32   [n]  (-1 == no collision)
    Barycentric: np.ndarray   # f32   [n,3]
    Stats: dict = field(default_factory=dict)


class UnsupportedError(NotImplementedError):
    pass


class MeshCollider:
    """GPU-backed model3d.Collider for a triangle mesh.

    ``triangles``: array [n,3,3] (or [n,9]) of float32-representable vertices; the
    index of a triangle in this array is its id.  ``vertex_normals`` ([n,3,3]) selects
    MeshToInterpNormalCollider semantics (collisions.go:147-162).
    """

    def __init__(self, triangles, vertex_normals=None, ctx=None, device_lbvh=False, device_build=False):
        """device_lbvh: build the binary hierarchy on the GPU (Morton codes + radix sort + Karras,
        M3D_MESH_BUILD_DEVICE_LBVH) instead of the host binned-SAH build: much faster to build,
        somewhat more nodes visited per ray; query results are identical.  device_build: the whole
        build on the GPU (M3D_MESH_BUILD_DEVICE_COLLAPSE: LBVH + cost-optimal 8-wide collapse + node
        emission)."""
        self.ctx = ctx or N.default_context()
        tris = np.ascontiguousarray(np.asarray(triangles, dtype=np.float32).reshape(-1, 9))
        vn = None
        if vertex_normals is not None:
            vn = np.ascontiguousarray(np.asarray(vertex_normals, dtype=np.float32).reshape(-1, 9))
            if vn.shape != tris.shape:
                raise ValueError("vertex_normals must match triangles")
        self.num_triangles = int(tris.shape[0])
        # host copy: scenes merge the triangles of all mesh objects into one BVH
        self.triangles = tris.reshape(-1, 3, 3)
        self.vertex_normals = None if vn is None else vn.reshape(-1, 3, 3)
        self.h = C.c_void_p()
        N.check(N.lib().m3d_mesh_create(self.ctx.h, _p(tris, f32p), C.c_int64(self.num_triangles),
                                        _p(vn, f32p), C.c_uint32(N.MESH_BUILD_DEVICE_COLLAPSE if device_build else (N.MESH_BUILD_DEVICE_LBVH if device_lbvh else 0)),
                                        C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            N.lib().m3d_mesh_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Bounder ---------------------------------------------------------------
    def _bounds(self):
        mn = (C.c_double * 3)()
        mx = (C.c_double * 3)()
        N.check(N.lib().m3d_mesh_bounds(self.h, mn, mx))
        return tuple(mn), tuple(mx)

    def Min(self):
        return self._bounds()[0]

    def Max(self):
        return self._bounds()[1]

    def Info(self):
        info = N.MeshInfo()
        N.check(N.lib().m3d_mesh_get_info(self.h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in info._fields_}

    # -- Collider --------------------------------------------------------------
    def FirstRayCollisions(self, origins, directions, counters=False, refine=True,
                           want_stats=False, normals=True, barycentric=True) -> BatchCollisions:
        """Batched FirstRayCollision (collisions.go:275-290) over host arrays [n,3].  `origins` may
        be a single point [3] shared by all rays (camera batches: only the directions cross the
        link); normals / barycentric = False drop those outputs (None in the result)."""
        org = np.ascontiguousarray(np.asarray(origins, dtype=np.float32).reshape(-1, 3))
        dr = np.ascontiguousarray(np.asarray(directions, dtype=np.float32).reshape(-1, 3))
        shared = org.shape[0] == 1 and dr.shape[0] != 1
        if org.shape != dr.shape and not shared:
            raise ValueError("origins and directions must have the same shape")
        n = int(dr.shape[0])
        t = np.zeros(n, np.float32)
        prim = np.full(n, -1, np.int32)
        normal = np.zeros((n, 3), np.float32) if normals else None
        bary = np.zeros((n, 3), np.float32) if barycentric else None
        flags = ((N.TRACE_COUNTERS if counters else 0) | (0 if refine else N.TRACE_NO_REFINE) |
                 (N.TRACE_SHARED_ORIGIN if shared else 0))
        stats = N.Stats()
        N.check(N.lib().m3d_mesh_first_ray_collisions(
            self.h, _p(org, f32p), _p(dr, f32p), C.c_int64(n), _p(t, f32p), _p(prim, i32p),
            _p(normal, f32p), _p(bary, f32p), C.c_uint32(flags),
            C.byref(stats) if (counters or want_stats) else None))
        st = {k: getattr(stats, k) for k, _ in stats._fields_} if (counters or want_stats) else {}
        return BatchCollisions(Collides=prim >= 0, Scale=t, Normal=normal, Triangle=prim,
                               Barycentric=bary, Stats=st)

    def FirstRayCollision(self, r: Ray):
        """Collider.FirstRayCollision: a batch of one (correct, but use the batch call)."""
        b = self.FirstRayCollisions([r.Origin], [r.Direction])
        if not b.Collides[0]:
            return RayCollision(0.0, (0.0, 0.0, 0.0)), False
        return RayCollision(
            Scale=float(b.Scale[0]), Normal=tuple(float(x) for x in b.Normal[0]),
            Extra=TriangleCollision(int(b.Triangle[0]), tuple(float(x) for x in b.Barycentric[0]))), True

    def FirstRayCollisionsDevice(self, d_org_tmin, d_dir_tmax, n, d_hit0, d_hit1, stream=0,
                                 counters=False, refine=True, want_stats=False):
        """Device-resident batch: arguments are raw device pointers (ints) to float4 SoA
        buffers as documented in include/m3d.h."""
        flags = (N.TRACE_COUNTERS if counters else 0) | (0 if refine else N.TRACE_NO_REFINE)
        stats = N.Stats()
        N.check(N.lib().m3d_mesh_first_ray_collisions_device(
            self.h, C.c_void_p(d_org_tmin), C.c_void_p(d_dir_tmax), C.c_int64(n),
            C.c_void_p(d_hit0), C.c_void_p(d_hit1), C.c_uint32(flags), C.c_void_p(stream),
            C.byref(stats) if (counters or want_stats) else None))
        return {k: getattr(stats, k) for k, _ in stats._fields_} if (counters or want_stats) else {}

    def RayCollisionCounts(self, origins, directions):
        """Batched Collider.RayCollisions(r, nil) (collisions.go:263-273): the number of
        triangles each ray's forward half-line crosses, int32 [n]."""
        org = np.ascontiguousarray(np.asarray(origins, dtype=np.float32).reshape(-1, 3))
        dr = np.ascontiguousarray(np.asarray(directions, dtype=np.float32).reshape(-1, 3))
        if org.shape != dr.shape:
            raise ValueError("origins and directions must have the same shape")
        counts = np.zeros(org.shape[0], np.int32)
        N.check(N.lib().m3d_mesh_ray_collision_counts(self.h, _p(org, f32p), _p(dr, f32p),
                                                      C.c_int64(org.shape[0]), _p(counts, i32p), None))
        return counts

    def RayCollisionsBatch(self, origins, directions, normals=True, barycentric=True):
        """Batched Collider.RayCollisions(r, f) with the collisions delivered (collisions.go:263-273,
        primitives.go:189-196).  Returns a dict: offsets int64 [n+1] (the collisions of ray i are
        rows offsets[i]:offsets[i+1], ordered by Scale), Scale float32 [total], Triangle int32
        [total], Normal float32 [total,3], Barycentric float32 [total,3] (None when not asked for)."""
        org = np.ascontiguousarray(np.asarray(origins, dtype=np.float32).reshape(-1, 3))
        dr = np.ascontiguousarray(np.asarray(directions, dtype=np.float32).reshape(-1, 3))
        if org.shape != dr.shape:
            raise ValueError("origins and directions must have the same shape")
        n = org.shape[0]
        offsets = np.zeros(n + 1, np.int64)
        i64p = C.POINTER(C.c_int64)
        fn = N.lib().m3d_mesh_ray_collisions
        N.check(fn(self.h, _p(org, f32p), _p(dr, f32p), C.c_int64(n), C.c_int64(0), _p(offsets, i64p),
                   None, None, None, None, None))
        total = int(offsets[n])
        t = np.zeros(total, np.float32)
        prim = np.zeros(total, np.int32)
        nrm = np.zeros((total, 3), np.float32) if normals else None
        bary = np.zeros((total, 3), np.float32) if barycentric else None
        if total > 0:
            N.check(fn(self.h, _p(org, f32p), _p(dr, f32p), C.c_int64(n), C.c_int64(total),
                       _p(offsets, i64p), _p(t, f32p), _p(prim, i32p),
                       _p(nrm, f32p) if normals else None, _p(bary, f32p) if barycentric else None, None))
        return dict(offsets=offsets, Scale=t, Triangle=prim, Normal=nrm, Barycentric=bary)

    def RayCollisions(self, r, f=None):
        """Collider.RayCollisions (collisions.go:263-273): calls f(RayCollision) for every triangle
        the ray crosses (in order of Scale) and returns their number."""
        if f is None:
            return int(self.RayCollisionCounts([r.Origin], [r.Direction])[0])
        b = self.RayCollisionsBatch([r.Origin], [r.Direction])
        for k in range(int(b["offsets"][1])):
            f(RayCollision(
                Scale=float(b["Scale"][k]), Normal=tuple(float(x) for x in b["Normal"][k]),
                Extra=TriangleCollision(int(b["Triangle"][k]), tuple(float(x) for x in b["Barycentric"][k]))))
        return int(b["offsets"][1])

    def Contains(self, coords, margin=0.0):
        """Batched model3d.ColliderContains(self, p, margin) (collisions.go:113-134), bool [n]."""
        pts = np.ascontiguousarray(np.asarray(coords, dtype=np.float32).reshape(-1, 3))
        inside = np.zeros(pts.shape[0], np.uint8)
        N.check(N.lib().m3d_mesh_contains(self.h, _p(pts, f32p), C.c_int64(pts.shape[0]), C.c_double(margin),
                                          inside.ctypes.data_as(C.POINTER(C.c_uint8)), None))
        return inside.astype(bool)

    def SphereCollisions(self, centers, radii):
        """Batched Collider.SphereCollision (collisions.go:292-303, primitives.go:253-279):
        bool [n], some triangle is closer than radii[i] to centers[i]."""
        pts = np.ascontiguousarray(np.asarray(centers, dtype=np.float32).reshape(-1, 3))
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, dtype=np.float32), pts.shape[:1]))
        out = np.zeros(pts.shape[0], np.uint8)
        N.check(N.lib().m3d_mesh_sphere_collisions(self.h, _p(pts, f32p), _p(rad, f32p), C.c_int64(pts.shape[0]),
                                                   out.ctypes.data_as(C.POINTER(C.c_uint8)), None))
        return out.astype(bool)

    def SphereCollision(self, c, r):
        """Collider.SphereCollision: a batch of one."""
        return bool(self.SphereCollisions([c], [r])[0])

    def FaceSDF(self, coords, want_stats=False):
        """Batched meshSDF.FaceSDF (sdf.go:229-240) on this collider's hierarchy: (face ids [n],
        nearest points [n,3], signed distances [n], flat normals [n,3])."""
        pts = np.ascontiguousarray(np.asarray(coords, dtype=np.float32).reshape(-1, 3))
        n = pts.shape[0]
        sdf = np.zeros(n, np.float32)
        cp = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32)
        face = np.full(n, -1, np.int32)
        stats = N.Stats()
        N.check(N.lib().m3d_mesh_sdf(self.h, _p(pts, f32p), C.c_int64(n), _p(sdf, f32p), _p(cp, f32p),
                                     _p(face, i32p), _p(nrm, f32p), C.byref(stats) if want_stats else None))
        if want_stats:
            return face, cp, sdf, nrm, {k: getattr(stats, k) for k, _ in stats._fields_}
        return face, cp, sdf, nrm


def MeshToCollider(triangles, ctx=None) -> MeshCollider:
    """model3d.MeshToCollider (collisions.go:138-142)."""
    return MeshCollider(triangles, ctx=ctx)


def MeshToInterpNormalCollider(triangles, vertex_normals=None, ctx=None) -> MeshCollider:
    """model3d.MeshToInterpNormalCollider (collisions.go:147-162): like the reference, computes
    the vertex normals itself (Mesh.VertexNormals, mesh_ops.go:146-169) unless they are given."""
    if vertex_normals is None:
        from . import meshes
        vertex_normals = meshes.VertexNormals(triangles)
    return MeshCollider(triangles, vertex_normals=vertex_normals, ctx=ctx)


def ColliderContains(c: MeshCollider, coords, margin=0.0):
    """model3d.ColliderContains (collisions.go:113-134), batched over coords [n,3]."""
    return c.Contains(coords, margin)


class ColliderSolid:
    """model3d.ColliderSolid (model3d/solid.go:236-300): a Solid whose Contains is the collider's
    parity test, batched; NewColliderSolidInset / NewColliderSolidHollow add the margin / shell
    variants through the nearest-triangle query."""

    def __init__(self, collider: MeshCollider, inset=0.0, radius=0.0):
        self.collider = collider
        cmin, cmax = np.asarray(collider.Min(), np.float64), np.asarray(collider.Max(), np.float64)
        self.inset, self.radius = float(inset), float(radius)
        if radius != 0:  # solid.go:274-279
            self.min, self.max = cmin - radius, cmax + radius
        elif inset != 0:  # solid.go:265-270
            self.min = cmin + inset
            self.max = np.maximum(self.min, cmax - inset)
        else:
            self.min, self.max = cmin, cmax

    def Min(self):
        return tuple(self.min)

    def Max(self):
        return tuple(self.max)

    def Contains(self, coords):
        pts = np.asarray(coords, dtype=np.float32).reshape(-1, 3)
        p64 = pts.astype(np.float64)
        inb = np.all((p64 >= np.asarray(self.min)) & (p64 <= np.asarray(self.max)), axis=1)  # InBounds (solid.go:293-295)
        out = np.zeros(pts.shape[0], bool)
        if inb.any():
            if self.radius != 0:
                out[inb] = self.collider.SphereCollisions(pts[inb], self.radius)
            else:
                out[inb] = self.collider.Contains(pts[inb], self.inset)
        return out


def NewColliderSolid(c: MeshCollider) -> ColliderSolid:
    return ColliderSolid(c)


def NewColliderSolidInset(c: MeshCollider, inset) -> ColliderSolid:
    """model3d.NewColliderSolidInset (solid.go:260-270)."""
    return ColliderSolid(c, inset=inset)


def NewColliderSolidHollow(c: MeshCollider, r) -> ColliderSolid:
    """model3d.NewColliderSolidHollow (solid.go:272-279)."""
    return ColliderSolid(c, radius=r)


class MeshSDF:
    """model3d.MeshToSDF (sdf.go:186-240): a FaceSDF over a triangle mesh, batched.  Every method
    takes coords [n,3]; distances are positive inside the mesh, negative outside."""

    def __init__(self, collider: MeshCollider):
        if collider.Info()["num_triangles"] == 0:
            raise ValueError("cannot create empty SDF")  # sdf.go:198-200
        self.collider = collider

    def Min(self):
        return self.collider.Min()

    def Max(self):
        return self.collider.Max()

    def SDF(self, coords):
        return self.collider.FaceSDF(coords)[2]

    def PointSDF(self, coords):
        _, cp, sdf, _ = self.collider.FaceSDF(coords)
        return cp, sdf

    def NormalSDF(self, coords):
        _, _, sdf, nrm = self.collider.FaceSDF(coords)
        return nrm, sdf

    def FaceSDF(self, coords):
        face, cp, sdf, _ = self.collider.FaceSDF(coords)
        return face, cp, sdf


def MeshToSDF(triangles, ctx=None) -> MeshSDF:
    """model3d.MeshToSDF (sdf.go:186-191); `triangles` may be an existing MeshCollider (the SDF
    and the collider share one device hierarchy)."""
    if isinstance(triangles, MeshCollider):
        return MeshSDF(triangles)
    return MeshSDF(MeshCollider(triangles, ctx=ctx))
