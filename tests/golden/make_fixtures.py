"""Generates the committed fixtures under tests/golden/ from the reference checkout.

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py
The GPU box has no /root/reference; tests only read the .npy files written here.

  diamond_tris.npy   the 46 triangles of examples/renderings/cornell_box/diamond.stl
                     (binary STL: 80-byte header, uint32 count, 50 B/triangle, float32 LE --
                     fileformats/stl.go:121-133,243-261), float32 [46,3,3].
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def read_binary_stl(path):
    data = open(path, "rb").read()
    (n,) = struct.unpack_from("<I", data, 80)
    tris = np.zeros((n, 3, 3), np.float32)
    for i in range(n):
        vals = struct.unpack_from("<12f", data, 84 + 50 * i)
        tris[i] = np.array(vals[3:12], np.float32).reshape(3, 3)  # skip the facet normal
    return tris


if __name__ == "__main__":
    tris = read_binary_stl(os.path.join(REF, "examples/renderings/cornell_box/diamond.stl"))
    assert tris.shape == (46, 3, 3), tris.shape
    np.save(os.path.join(HERE, "diamond_tris.npy"), tris)
    print("diamond:", tris.shape, tris.min(axis=(0, 1)), tris.max(axis=(0, 1)))
