"""Generates the committed fixtures under tests/golden/ from the reference checkout.

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py
The GPU box has no /root/reference; tests only read the .npy files written here.

  diamond_tris.npy   the 46 triangles of examples/renderings/cornell_box/diamond.stl
                     (binary STL: 80-byte header, uint32 count, 50 B/triangle, float32 LE --
                     fileformats/stl.go:121-133,243-261), float32 [46,3,3].
  showcase_models.npz  the seven meshes of examples/renderings/showcase/models/*.stl.gz
                     (BASELINE config 4; vase.stl.gz is missing from the reference checkout),
                     each stored indexed: <name>_v float32 [V,3] unique vertices in first-use
                     order, <name>_f int32 [n,3]; triangles = v[f] in file order.
  ref_cornell_box_output.png / ref_cornell_box_output_hd.png
                     byte copies of examples/renderings/cornell_box/output.png (200x200; the
                     configuration of cornell_box/main.go: MaxDepth 5, NumSamples 400, Antialias 1,
                     Cutoff 1e-4, PhongFocusPoint 0.3) and output_hd.png (500x500; README.md:
                     "MaxDepth to maybe 15, NumSamples to 20000"): renderings produced by the Go
                     reference itself and committed upstream -- the only reference OUTPUTS for
                     this path, used to pin the oracle and the GPU path statistically
                     (tests/test_reference_golden.py).
  ref_showcase_output.png
                     byte copy of examples/renderings/showcase/output.png (480x320; the configuration of
                     showcase/main.go: MaxDepth 10, NumSamples 50, Antialias 1, Cutoff 1e-4,
                     SphereFocusPoint 0.3), produced by the Go reference and committed upstream as a
                     244-colour palette PNG.  It shows the vase whose mesh is missing from the checkout;
                     tests mask that region (tests/test_reference_golden.py).  ref_showcase_output_hd.png
                     is output_hd.png of the same directory (960x640, main.go's HighRes mode: adaptive
                     1000...100000 spp until MaxStddev 0.02).
  ref_smooth_shading_rendering.png
                     byte copy of examples/renderings/smooth_shading/rendering.png (768x432): the
                     deterministic RayCaster image of smooth_shading/main.go -- NewMeshIcosphere(0,1,4)
                     twice, flat (MeshToCollider) and smooth (MeshToInterpNormalCollider), Phong
                     material, one point light, rendered at 4x and box-filtered.  The text labels under
                     the spheres need model2d and are masked.
  ref_rose_rendering.png
                     byte copy of examples/decoration/rose/rendering.png: render3d.SaveRendering of
                     the rose mesh from (0,-2,4) at 500x500 (rose/main.go:31), a deterministic
                     RayCaster image produced by the Go reference.  showcase/models/rose.stl.gz is a
                     later export of the same model (same silhouette, slightly different facets).
"""
import os
import shutil
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


SHOWCASE = ["curvy_thing", "pumpkin_inside", "pumpkin_outside", "pumpkin_stem", "rocks", "rose", "wine_glass"]


def read_stl_gz_indexed(path):
    import gzip
    data = gzip.decompress(open(path, "rb").read())
    (n,) = struct.unpack_from("<I", data, 80)
    rec = np.dtype([("normal", "<f4", (3,)), ("verts", "<f4", (3, 3)), ("attr", "<u2")])
    tris = np.frombuffer(data, dtype=rec, count=n, offset=84)["verts"].reshape(-1, 3)
    # unique vertices in first-use order (bit patterns, so -0.0 and 0.0 stay distinct)
    keys = np.ascontiguousarray(tris).view(np.dtype((np.void, 12))).ravel()
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    verts = tris[first[order]]
    faces = rank[inv].astype(np.int32).reshape(n, 3)
    assert np.array_equal(verts[faces].reshape(-1, 3).view(np.uint32), tris.view(np.uint32))
    return verts.astype(np.float32), faces


if __name__ == "__main__":
    out = {}
    for name in SHOWCASE:
        v, f = read_stl_gz_indexed(os.path.join(REF, "examples/renderings/showcase/models", name + ".stl.gz"))
        out[name + "_v"], out[name + "_f"] = v, f
        print(name, f.shape[0], "triangles", v.shape[0], "vertices")
    np.savez_compressed(os.path.join(HERE, "showcase_models.npz"), **out)

    tris = read_binary_stl(os.path.join(REF, "examples/renderings/cornell_box/diamond.stl"))
    assert tris.shape == (46, 3, 3), tris.shape
    np.save(os.path.join(HERE, "diamond_tris.npy"), tris)
    print("diamond:", tris.shape, tris.min(axis=(0, 1)), tris.max(axis=(0, 1)))

    shutil.copyfile(os.path.join(REF, "examples/decoration/rose/rendering.png"), os.path.join(HERE, "ref_rose_rendering.png"))
    shutil.copyfile(os.path.join(REF, "examples/renderings/smooth_shading/rendering.png"),
                    os.path.join(HERE, "ref_smooth_shading_rendering.png"))
    shutil.copyfile(os.path.join(REF, "examples/renderings/showcase/output.png"), os.path.join(HERE, "ref_showcase_output.png"))
    shutil.copyfile(os.path.join(REF, "examples/renderings/showcase/output_hd.png"), os.path.join(HERE, "ref_showcase_output_hd.png"))
    for src, dst in (("output.png", "ref_cornell_box_output.png"), ("output_hd.png", "ref_cornell_box_output_hd.png")):
        shutil.copyfile(os.path.join(REF, "examples/renderings/cornell_box", src), os.path.join(HERE, dst))
        print("copied", src, "->", dst)
