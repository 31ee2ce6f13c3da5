"""Host-side mesh synthesis mirroring the reference's generators, used to build benchmark
and example inputs without touching the oracle:

  NewMeshIcosphere   model3d/mesh.go:124-128 (NewMeshIcosahedron :304-330 +
                     SubdivideEdges subdivision.go:88-139 + Normalize/Scale/Translate)
  NewMeshRect        model3d/mesh.go:132-165

Triangles are returned as float64 arrays [n, 3, 3] in insertion order (the reference's
Mesh is an unordered set; insertion order is this package's deterministic triangle id).
"""
import math

import numpy as np


def _geo(lat, lon):
    """GeoCoord.Coord3D (model3d/coords.go:32-38)."""
    return np.array([math.sin(lon) * math.cos(lat), math.sin(lat), math.cos(lon) * math.cos(lat)])


def NewMeshIcosahedron():
    mid_lat = math.atan(0.5)

    def top_c(i):
        return _geo(-mid_lat, math.pi * 2 * float(i % 5) / 5.0)

    def bot_c(i):
        return _geo(mid_lat, math.pi * 2 * (1.0 / 10.0 + float(i % 5) / 5.0))

    top, bottom = _geo(-math.pi / 2, 0), _geo(math.pi / 2, 0)
    tris = []
    for i in range(5):
        tris.append([top, top_c(i + 1), top_c(i)])
        tris.append([bottom, bot_c(i), bot_c(i + 1)])
        tris.append([top_c(i), top_c(i + 1), bot_c(i)])
        tris.append([bot_c(i + 1), bot_c(i), top_c(i + 1)])
    return np.array(tris, np.float64)


def _first_is(p1, p2):
    """NewSegment's canonical order (model3d/primitives.go:547-554)."""
    return (p1[0] < p2[0] or (p1[0] == p2[0] and p1[1] < p2[1]) or
            (p1[0] == p2[0] and p1[1] == p2[1] and p1[2] < p2[2]))


def _divide_segment(c1, c2, length):
    """divideSegment (subdivision.go:117-139): `length` points from c1 to c2, interpolated
    from the canonically-first endpoint so that shared edges get identical points."""
    if length == 1:
        return c1[None, :].copy()
    if not _first_is(c1, c2) and not np.array_equal(c1, c2):
        return _divide_segment(c2, c1, length)[::-1].copy()
    t = (np.arange(length, dtype=np.float64) / float(length - 1))[:, None]
    res = c1[None, :] * (1 - t) + c2[None, :] * t
    res[0] = c1
    res[-1] = c2
    return res


def SubdivideEdges(tris, n):
    """subdivision.go:88-115: every triangle becomes n*n triangles."""
    out = np.empty((tris.shape[0] * n * n, 3, 3), np.float64)
    w = 0
    for t in tris:
        side1 = _divide_segment(t[0], t[1], n + 1)
        side2 = _divide_segment(t[0], t[2], n + 1)
        for i in range(n):
            nl = i + 1
            narrow = _divide_segment(side1[i], side2[i], nl)
            wide = _divide_segment(side1[i + 1], side2[i + 1], nl + 1)
            # k = 0: (narrow0, wide0, wide1); k > 0: (narrow_k, wide_k, wide_k+1) then
            # (narrow_k, narrow_k-1, wide_k)
            cnt = 2 * nl - 1
            blk = out[w:w + cnt]
            up = np.stack([narrow, wide[:-1], wide[1:]], axis=1)          # nl triangles
            blk[0] = up[0]
            if nl > 1:
                down = np.stack([narrow[1:], narrow[:-1], wide[1:-1]], axis=1)  # nl-1 triangles
                blk[1::2] = up[1:]
                blk[2::2] = down
            w += cnt
    assert w == out.shape[0]
    return out


def NewMeshIcosphere(center, radius, n):
    """model3d.NewMeshIcosphere (mesh.go:124-128): 20*n*n triangles on a sphere."""
    m = SubdivideEdges(NewMeshIcosahedron(), int(n))
    x, y, z = m[..., 0], m[..., 1], m[..., 2]
    inv = 1.0 / np.sqrt(x * x + y * y + z * z)  # Coord3D.Normalize = Scale(1/Norm) (coords.go:379-381)
    m = m * inv[..., None]
    m = m * float(radius)
    return m + np.asarray(center, np.float64)[None, None, :]


def NewMeshRect(mn, mx):
    """model3d.NewMeshRect (mesh.go:132-165): 12 triangles, insertion order."""
    mn, mx = np.asarray(mn, np.float64), np.asarray(mx, np.float64)

    def pt(x, y, z):
        return np.array([mx[0] if x else mn[0], mx[1] if y else mn[1], mx[2] if z else mn[2]])

    quads = [(mn, pt(1, 0, 0), pt(1, 0, 1), pt(0, 0, 1)), (mx, pt(1, 1, 0), pt(0, 1, 0), pt(0, 1, 1)),
             (mn, pt(0, 0, 1), pt(0, 1, 1), pt(0, 1, 0)), (mx, pt(1, 0, 1), pt(1, 0, 0), pt(1, 1, 0)),
             (mn, pt(0, 1, 0), pt(1, 1, 0), pt(1, 0, 0)), (mx, pt(0, 1, 1), pt(0, 0, 1), pt(1, 0, 1))]
    tris = []
    for p1, p2, p3, p4 in quads:  # AddQuad (mesh.go:389-397)
        tris.append([p1, p2, p4])
        tris.append([p2, p3, p4])
    return np.array(tris, np.float64)


# ---------------------------------------------------------------------------------------
# Marching cubes (model3d/mc.go), vectorised over grid layers.  BASELINE config C1 is
# "model3d.Sphere -> MarchingCubesSearch(0.01, 8)".

_MC_BASE = [  # baseTriangleTable (mc.go:460-586): (corners inside, triangles as 3 cube edges)
    ((), ()),
    ((0,), ((0, 1, 0, 2, 0, 4),)),
    ((0, 1), ((0, 4, 1, 5, 0, 2), (1, 5, 1, 3, 0, 2))),
    ((0, 5), ((0, 1, 0, 2, 0, 4), (5, 7, 1, 5, 4, 5))),
    ((0, 7), ((0, 1, 0, 2, 0, 4), (6, 7, 3, 7, 5, 7))),
    ((1, 2, 3), ((0, 1, 1, 5, 0, 2), (0, 2, 1, 5, 2, 6), (2, 6, 1, 5, 3, 7))),
    ((0, 1, 7), ((0, 4, 1, 5, 0, 2), (1, 5, 1, 3, 0, 2), (6, 7, 3, 7, 5, 7))),
    ((1, 4, 7), ((4, 6, 4, 5, 0, 4), (1, 5, 1, 3, 0, 1), (6, 7, 3, 7, 5, 7))),
    ((0, 1, 2, 3), ((0, 4, 1, 5, 3, 7), (0, 4, 3, 7, 2, 6))),
    ((0, 2, 3, 6), ((0, 1, 4, 6, 0, 4), (0, 1, 6, 7, 4, 6), (0, 1, 1, 3, 6, 7), (1, 3, 3, 7, 6, 7))),
    ((1, 2, 5, 6), ((0, 2, 2, 3, 6, 7), (0, 2, 6, 7, 4, 6), (0, 1, 4, 5, 5, 7), (5, 7, 1, 3, 0, 1))),
    ((0, 2, 3, 7), ((0, 4, 0, 1, 2, 6), (0, 1, 5, 7, 2, 6), (2, 6, 5, 7, 6, 7), (0, 1, 1, 3, 5, 7))),
    ((1, 2, 3, 4), ((0, 1, 1, 5, 0, 2), (0, 2, 1, 5, 2, 6), (2, 6, 1, 5, 3, 7), (4, 5, 0, 4, 4, 6))),
    ((1, 2, 4, 7), ((0, 1, 1, 5, 1, 3), (0, 2, 2, 3, 2, 6), (4, 5, 0, 4, 4, 6), (5, 7, 6, 7, 3, 7))),
    ((1, 2, 3, 6), ((0, 2, 0, 1, 4, 6), (0, 1, 3, 7, 4, 6), (0, 1, 1, 5, 3, 7), (4, 6, 3, 7, 6, 7))),
    ((0, 2, 3, 5, 6), ((0, 1, 4, 6, 0, 4), (0, 1, 6, 7, 4, 6), (0, 1, 1, 3, 6, 7), (1, 3, 3, 7, 6, 7),
                       (5, 7, 1, 5, 4, 5))),
    ((2, 3, 4, 5, 6), ((5, 7, 1, 5, 0, 4), (0, 4, 6, 7, 5, 7), (0, 2, 6, 7, 0, 4), (0, 2, 3, 7, 6, 7),
                       (0, 2, 1, 3, 3, 7))),
    ((0, 4, 5, 6, 7), ((1, 5, 0, 1, 0, 2), (0, 2, 2, 6, 1, 5), (1, 5, 2, 6, 3, 7))),
    ((1, 2, 3, 4, 5, 6), ((0, 2, 0, 1, 0, 4), (3, 7, 6, 7, 5, 7))),
    ((1, 2, 3, 4, 6, 7), ((0, 2, 4, 5, 0, 4), (0, 2, 5, 7, 4, 5), (0, 2, 1, 5, 5, 7), (0, 1, 1, 5, 0, 2))),
    ((2, 3, 4, 5, 6, 7), ((1, 5, 0, 4, 0, 2), (1, 3, 1, 5, 0, 2))),
    ((1, 2, 3, 4, 5, 6, 7), ((0, 2, 0, 1, 0, 4),)),
    ((0, 1, 2, 3, 4, 5, 6, 7), ()),
]
_MC_TABLE = None


def _mc_lookup_table():
    """mcLookupTable (mc.go:431-454): the 24 cube rotations (closure of a z and an x quarter
    turn, sorted lexicographically, mc.go:312-352) applied to the base cases; the first
    rotation that produces a corner mask defines its triangles.  Returns (counts[256],
    corners[256, 5, 6])."""
    global _MC_TABLE
    if _MC_TABLE is not None:
        return _MC_TABLE
    zr, xr = (2, 0, 3, 1, 6, 4, 7, 5), (2, 3, 6, 7, 0, 1, 4, 5)
    queue, seen = [tuple(range(8))], {tuple(range(8))}
    while queue:
        nxt = queue.pop(0)
        for op in (zr, xr):
            rot = tuple(op[nxt[i]] for i in range(8))  # Compose (mc.go:355-361)
            if rot not in seen:
                seen.add(rot)
                queue.append(rot)
    rots = sorted(seen)
    assert len(rots) == 24
    counts = np.zeros(256, np.int32)
    corners = np.zeros((256, 5, 6), np.uint8)
    done = np.zeros(256, bool)
    for inside, tris in _MC_BASE:
        for rot in rots:
            bits = 0
            for c in inside:
                bits |= 1 << rot[c]
            if done[bits]:
                continue
            done[bits] = True
            counts[bits] = len(tris)
            for k, t in enumerate(tris):
                corners[bits, k] = [rot[c] for c in t]
    assert done.all()
    _MC_TABLE = (counts, corners)
    return _MC_TABLE


class SphereSolid:
    """model3d.Sphere as a Solid (shapes.go:17-31), Contains vectorised over [n, 3] points."""

    def __init__(self, center, radius):
        self.Center = np.asarray(center, np.float64)
        self.Radius = float(radius)

    def Min(self):
        return self.Center - self.Radius

    def Max(self):
        return self.Center + self.Radius

    def Contains(self, pts):
        d = pts - self.Center
        return np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]) <= self.Radius


def _mc_spacer(solid, delta):
    """newSquareSpacer (mc.go:594-612): the same running sums as the reference's loops."""
    out = []
    mn, mx = solid.Min(), solid.Max()
    for a in range(3):
        vals, v = [], float(mn[a]) - delta
        while v <= float(mx[a]) + delta:
            vals.append(v)
            v += delta
        out.append(np.array(vals, np.float64))
    return out


def MarchingCubes(solid, delta):
    """model3d.MarchingCubes (mc.go:14-37).  Triangles in scan order (z, y, x, table order);
    the reference's Mesh is an unordered set."""
    counts, tcorn = _mc_lookup_table()
    xs, ys, zs = _mc_spacer(solid, float(delta))
    nx, ny, nz = len(xs), len(ys), len(zs)
    gx, gy = np.meshgrid(xs, ys)  # [ny, nx]

    def layer(z):
        pts = np.stack([gx.ravel(), gy.ravel(), np.full(nx * ny, zs[z])], axis=1)
        v = np.asarray(solid.Contains(pts), bool).reshape(ny, nx)
        if v.any() and (z == 0 or z == nz - 1 or v[0].any() or v[-1].any() or v[:, 0].any() or v[:, -1].any()):
            raise ValueError("solid is true outside of bounds")  # mc.go:686-688
        return v

    def square(v):  # GetSquare (mc.go:697-710): bit = x + 2*y
        v = v.astype(np.uint8)
        return v[:-1, :-1] | (v[:-1, 1:] << 1) | (v[1:, :-1] << 2) | (v[1:, 1:] << 3)

    # corner c of a cell = (x + (c&1), y + ((c>>1)&1), z-1 + (c>>2)) (mc.go:289-301)
    out = []
    bottom = layer(0)
    for z in range(1, nz):
        top = layer(z)
        bits = square(bottom) | (square(top) << 4)
        cy, cx = np.nonzero(counts[bits])
        if cy.size:
            b = bits[cy, cx]
            cnt = counts[b]
            cell = np.repeat(np.arange(cy.size), cnt)
            first = np.cumsum(cnt) - cnt
            k = np.arange(cell.size) - np.repeat(first, cnt)
            tc = tcorn[b[cell], k]  # [m, 6] corner ids
            px = xs[cx[cell][:, None] + (tc & 1)]
            py = ys[cy[cell][:, None] + ((tc >> 1) & 1)]
            pz = zs[(z - 1) + (tc >> 2)]
            p = np.stack([px, py, pz], axis=-1)  # [m, 6, 3]
            out.append((p[:, 0::2] + p[:, 1::2]) * 0.5)  # Coord3D.Mid (coords.go:196-198)
        bottom = top
    if not out:
        return np.zeros((0, 3, 3), np.float64)
    return np.concatenate(out, axis=0)


def MarchingCubesSearch(solid, delta, iters):
    """model3d.MarchingCubesSearch (mc.go:45-51): every vertex is moved along its cube edge by
    `iters` bisection steps of solid.Contains (mcSearch / mcSearchPoint mc.go:182-260,
    LookupEdgePoint :647-660).  The result depends on the vertex alone, so all triangle
    corners are processed independently."""
    mesh = MarchingCubes(solid, delta)
    if iters == 0 or mesh.shape[0] == 0:
        return mesh
    sp = _mc_spacer(solid, float(delta))
    d = sp[0][1] - sp[0][0]
    pts = mesh.reshape(-1, 3).copy()
    n = pts.shape[0]
    axis = np.full(n, -1)
    fp = np.zeros(n)
    tp = np.zeros(n)
    for a in range(3):
        rel = pts[:, a] - sp[a][0]
        modulus = np.abs(np.fmod(rel, d))
        sel = (axis < 0) & (modulus > d / 4) & (modulus < 3 * d / 4)
        idx = (rel[sel] / d).astype(np.int64)
        fp[sel] = sp[a][idx]
        tp[sel] = sp[a][idx + 1]
        axis[sel] = a
    if (axis < 0).any():
        raise ValueError("vertex not on edge")
    rows = np.arange(n)
    probe = pts.copy()
    probe[rows, axis] = tp
    swap = ~np.asarray(solid.Contains(probe), bool)
    fp[swap], tp[swap] = tp[swap], fp[swap].copy()
    for _ in range(int(iters)):
        mid = (fp + tp) / 2
        probe[rows, axis] = mid
        inside = np.asarray(solid.Contains(probe), bool)
        tp = np.where(inside, mid, tp)
        fp = np.where(inside, fp, mid)
    pts[rows, axis] = (fp + tp) / 2
    return pts.reshape(-1, 3, 3)


def VertexNormals(tris):
    """Mesh.VertexNormals (model3d/mesh_ops.go:146-169) as per-corner normals [n,3,3]: for every
    vertex the sum of the flat normals of its triangles, each weighted by the triangle's interior
    angle at that vertex, normalised.  Vertices are identified by exact coordinates like the
    reference's CoordMap.  This is what MeshToInterpNormalCollider (collisions.go:147-162) feeds to
    its InterpNormalTriangles."""
    t = np.asarray(tris, np.float64).reshape(-1, 3, 3)
    n = t.shape[0]

    def unit(v):
        return v * (1.0 / np.sqrt((v * v).sum(axis=-1, keepdims=True)))

    edges = np.stack([unit(t[:, 0] - t[:, 1]), unit(t[:, 1] - t[:, 2]), unit(t[:, 2] - t[:, 0])], axis=1)
    normal = unit(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]))
    weighted = np.zeros((n, 3, 3))
    for i in range(3):
        e1, e2 = edges[:, (i + 2) % 3], edges[:, i]
        theta = np.arccos(np.clip(-(e1 * e2).sum(axis=1), -1.0, 1.0))
        weighted[:, i] = normal * theta[:, None]
    keys = np.ascontiguousarray(t.reshape(-1, 3) + 0.0).view(np.dtype((np.void, 24))).ravel()  # -0.0 == 0.0
    _, inv = np.unique(keys, return_inverse=True)
    sums = np.zeros((inv.max() + 1, 3))
    np.add.at(sums, inv, weighted.reshape(-1, 3))
    return unit(sums)[inv].reshape(n, 3, 3)
