"""STL input of the GPU path (model3d_b200/fileformats.py) against the format rules of the
reference reader (fileformats/stl.go:60-133,243-261; model3d/import.go:13-43).  CPU only."""
import gzip
import io
import struct

import numpy as np
import pytest

from model3d_b200 import fileformats as F


def tris(n, seed=0):
    return np.random.default_rng(seed).normal(size=(n, 3, 3)).astype(np.float32)


def test_binary_roundtrip_and_gzip(tmp_path):
    t = tris(1000)
    p = tmp_path / "a.stl"
    F.WriteSTL(str(p), t)
    data = p.read_bytes()
    assert len(data) == 84 + 50 * 1000 and struct.unpack_from("<I", data, 80)[0] == 1000
    assert np.array_equal(F.ReadSTL(str(p)).view(np.uint32), t.view(np.uint32))  # bit exact
    pz = tmp_path / "a.stl.gz"
    F.WriteSTL(str(pz), t, compress=True)
    assert pz.read_bytes()[:2] == b"\x1f\x8b"
    assert np.array_equal(F.ReadSTL(str(pz)), t)
    assert np.array_equal(F.ReadSTL(io.BytesIO(gzip.compress(data))), t)
    assert np.array_equal(F.ReadSTL(data), t)


def test_binary_that_starts_with_solid():
    """stl.go:101-114: 'solid' prefix alone does not make a file ASCII."""
    t = tris(3, 1)
    buf = io.BytesIO()
    F.WriteSTL(buf, t)
    data = bytearray(buf.getvalue())
    data[:5] = b"solid"
    assert np.array_equal(F.ReadSTL(bytes(data)), t)


def test_ascii():
    txt = b"solid x\nfacet normal 0 0 1\nouter loop\nvertex 0 0 0\nvertex 1 0 0\nvertex 0 1.5 0\nendloop\nendfacet\nendsolid x\n"
    got = F.ReadSTL(txt)
    assert got.shape == (1, 3, 3) and got[0, 2, 1] == 1.5


def test_errors_and_empty():
    with pytest.raises(ValueError):
        F.ReadSTL(b"")
    with pytest.raises(ValueError):
        F.ReadSTL(b"\x00" * 50)
    buf = io.BytesIO()
    F.WriteSTL(buf, tris(4))
    with pytest.raises(ValueError):
        F.ReadSTL(buf.getvalue()[:-10])  # truncated record
    buf = io.BytesIO()
    F.WriteSTL(buf, np.zeros((0, 3, 3), np.float32))
    assert F.ReadSTL(buf.getvalue()).shape == (0, 3, 3)


def test_showcase_fixture_shapes():
    from model3d_b200 import examples
    m = examples.showcase_models()
    counts = {k: v.shape[0] for k, v in m.items()}
    assert counts == {"curvy_thing": 48090, "pumpkin_inside": 9387, "pumpkin_outside": 101965,
                      "pumpkin_stem": 3272, "rocks": 11956, "rose": 54758, "wine_glass": 96708}
    spec = examples.showcase()
    assert len(spec["objects"]) == 11 and sum(o["tris"].shape[0] for o in spec["objects"] if o["kind"] == "mesh") == 326136
