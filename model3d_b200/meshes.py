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
