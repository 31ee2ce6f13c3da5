"""On-disk mesh input for the GPU path: STL (binary or ASCII, optionally gzip-compressed)
straight to the float32 triangle arrays the device BVH is built from.

Mirrors fileformats.STLReader / model3d.ReadSTL (fileformats/stl.go:60-133,243-261,
model3d/import.go:13-43): binary STL is an 80-byte header, a little-endian uint32 triangle
count and 50-byte records (facet normal 3 x f32 -- ignored like the reference --, three
vertices 3 x f32, 2 attribute bytes).  Files that start with "solid" and contain only ASCII in
their first 512 bytes are parsed as text (stl.go:101-114).  Vertices are float32 on disk, i.e.
already the device format: no precision is lost on the way to HBM.
"""
import gzip
import io
import struct

import numpy as np

_REC = np.dtype([("normal", "<f4", (3,)), ("verts", "<f4", (3, 3)), ("attr", "<u2")])
assert _REC.itemsize == 50


def _is_ascii_chunk(chunk: bytes) -> bool:
    if len(chunk) < 5 or chunk[:5] != b"solid":
        return False
    return all(0 < b <= 127 for b in chunk)


def _read_all(src) -> bytes:
    if isinstance(src, (bytes, bytearray)):
        data = bytes(src)
    elif hasattr(src, "read"):
        data = src.read()
    else:
        with open(src, "rb") as f:
            data = f.read()
    if data[:2] == b"\x1f\x8b":  # gzip magic (the examples ship *.stl.gz)
        data = gzip.decompress(data)
    return data


def ReadSTL(src) -> np.ndarray:
    """Decode an STL file (path, bytes or file object; gzip detected by magic) into a float32
    array [n, 3, 3] in file order (file order is the triangle id order of the GPU path)."""
    data = _read_all(src)
    if len(data) == 0:
        raise ValueError("read STL: read STL header: unexpected EOF")
    if _is_ascii_chunk(data[:512]):
        return _read_ascii(data)
    if len(data) < 84:
        raise ValueError("read STL: read STL header: unexpected EOF")
    (n,) = struct.unpack_from("<I", data, 80)
    if len(data) < 84 + 50 * n:
        raise ValueError("read STL: unexpected EOF after %d of %d triangles" % ((len(data) - 84) // 50, n))
    rec = np.frombuffer(data, dtype=_REC, count=n, offset=84)
    return np.ascontiguousarray(rec["verts"], dtype=np.float32)


def _read_ascii(data: bytes) -> np.ndarray:
    verts = []
    for line in io.BytesIO(data):
        parts = line.split()
        if len(parts) == 4 and parts[0] == b"vertex":
            verts.append([float(parts[1]), float(parts[2]), float(parts[3])])
    if len(verts) % 3 != 0:
        raise ValueError("read STL: incomplete facet in ASCII STL")
    return np.asarray(verts, np.float32).reshape(-1, 3, 3)


def WriteSTL(dst, triangles, compress=False) -> None:
    """Binary STL with facet normals by the right-hand rule (fileformats/stl.go:19-58)."""
    tris = np.asarray(triangles, np.float32).reshape(-1, 3, 3)
    rec = np.zeros(tris.shape[0], dtype=_REC)
    n = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]).astype(np.float64)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    rec["normal"] = (n / np.where(ln > 0, ln, 1.0)).astype(np.float32)
    rec["verts"] = tris
    data = b"\x00" * 80 + struct.pack("<I", tris.shape[0]) + rec.tobytes()
    if compress:
        data = gzip.compress(data)
    if hasattr(dst, "write"):
        dst.write(data)
    else:
        with open(dst, "wb") as f:
            f.write(data)
