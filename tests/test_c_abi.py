"""The cgo binding's call sequence, made from C (tests/c_abi/cgo_sequence.c).

go/gpu3d cannot be compiled in the build container (no Go toolchain), so the same sequence of
C-ABI calls -- context, scene builder with materials / flags / transforms / vertex normals,
m3d_scene_cast, the three renderers with chunked sample ranges and pinned host buffers, the error
path, teardown -- is compiled with gcc against include/m3d.h, linked with libm3dgpu.so and run on
the BASELINE scenes C3 (cornell box) and C4 (showcase).  Its output must equal the Python
binding's render of the same scene and seed (both are thin layers over the same calls) and agree
with the oracle statistically (the Python binding's own parity tests cover the rest)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "cgo_sequence.c")


def build_program(tmp_path):
    exe = str(tmp_path / "cgo_sequence")
    libdir = os.path.join(ROOT, "model3d_b200")
    subprocess.check_call(["gcc", "-Wall", "-Wextra", "-Werror", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"),
                           SRC, "-L", libdir, "-lm3dgpu", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_c_program_compiles_against_the_header(built, tmp_path):
    """CPU part: include/m3d.h is valid C99 and every symbol the sequence uses links."""
    build_program(tmp_path)


def d3(v):
    return struct.pack("<3d", *[float(x) for x in v])


def write_scene(path, record, cam, W, H, mode, tail):
    with open(path, "wb") as f:
        for r in record:
            if r[0] == "material":
                f.write(struct.pack("<i", 1) + r[1])
                continue
            tag = {"sphere": 2, "rect": 3, "cylinder": 4, "mesh": 5}[r[0]]
            f.write(struct.pack("<iiIi", tag, r[1], r[2], 1 if r[3] is not None else 0))
            if r[3] is not None:
                f.write(r[3])
            if r[0] == "sphere":
                f.write(d3(r[4]) + struct.pack("<d", r[5]))
            elif r[0] == "rect":
                f.write(d3(r[4]) + d3(r[5]))
            elif r[0] == "cylinder":
                f.write(d3(r[4]) + d3(r[5]) + struct.pack("<d", r[6]))
            else:
                tris, vn = r[4], r[5]
                f.write(struct.pack("<qi", tris.shape[0], 1 if vn is not None else 0))
                f.write(np.ascontiguousarray(tris, np.float32).tobytes())
                if vn is not None:
                    f.write(np.ascontiguousarray(vn, np.float32).tobytes())
        f.write(struct.pack("<i", 0))
        f.write(struct.pack("<i", mode) + bytes(cam) + struct.pack("<ii", W, H))
        f.write(tail)


def read_out(path):
    raw = open(path, "rb").read()
    W, H, var = struct.unpack_from("<iii", raw, 0)
    a = np.frombuffer(raw, np.float32, offset=12)
    n = W * H * 3
    return a[:n].reshape(H, W, 3), (a[n:2 * n].reshape(H, W, 3) if var else None)


def run_program(exe, scene_file, out_file, devices=1, chunks=1):
    p = subprocess.run([exe, scene_file, out_file, str(devices), str(chunks)], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.strip().endswith("ok"), p.stdout
    return p.stdout, p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c3", "c4"])
def test_path_tracer_from_c_equals_python_binding(built, tmp_path, case):
    from model3d_b200 import render3d as R
    exe = build_program(tmp_path)
    if case == "c3":
        spec, W, H, depth, spp = scenes.cornell_box(), 96, 96, 5, 48
    else:
        spec, W, H, depth, spp = scenes.showcase(hd=False), 120, 80, 10, 24
    rec = []
    psc = scenes.build_product(spec, record=rec)
    tr = scenes.product_tracer(spec, psc, depth, spp, cutoff=1e-4, antialias=1.0, seed=77)
    want, want_sq, _ = tr.RenderSums(W, H, psc, sample_count=spp, variance=True)
    p = tr._params(psc, spp)
    p.min_samples = 0
    tail = bytes(p) + struct.pack("<i", 0) + struct.pack("<ii", spp, 1)
    sf, of = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(sf, rec, tr.Camera._c(), W, H, 1, tail)
    out, err = run_program(exe, sf, of, chunks=3)
    assert err.count("LogFunc(") == 3  # progress between the chunks, like the Go LogFunc
    got, got_sq = read_out(of)
    # same kernels, same Philox streams; three sample chunks only change the float32 summation order
    assert np.allclose(got, want, rtol=2e-5, atol=1e-4)
    assert np.allclose(got_sq, want_sq, rtol=1e-4, atol=1e-3)
    assert "objects %d materials" % len(psc.objects) in out
    assert want.mean() / spp > 0.02  # the scene is lit
    kinds = {r[0] for r in rec}
    if case == "c4":  # flags, transforms, all collider kinds and procedural materials crossed the ABI from C
        assert {"mesh", "sphere", "rect", "cylinder", "material"} <= kinds
        assert any(r[0] != "material" and r[2] != 0 for r in rec)


@pytest.mark.gpu
def test_raycaster_and_bidir_from_c_equal_python_binding(built, tmp_path):
    from model3d_b200 import render3d as R, _native as N
    exe = build_program(tmp_path)
    # RayCaster on a smooth-shaded (vertex normals) + flat mesh scene with a transformed sphere
    spec = scenes.mixed_scene()
    spec["camera"] = dict(src=(4.0, -6.0, 3.0), dst=(0.0, 0.0, 0.0), fov=np.pi / 3.6)
    rec = []
    psc = scenes.build_product(spec, record=rec)
    cam = spec["camera"]
    lt = R.PointLight(Origin=(30.0, -40.0, 50.0), Color=(1.0, 0.9, 0.8))
    rc = R.RayCaster(Camera=R.NewCameraAt(cam["src"], cam["dst"], cam["fov"]), Lights=[lt])
    img = R.Image(80, 60)
    rc.Render(img, psc)
    sf, of = str(tmp_path / "rc.bin"), str(tmp_path / "rc_out.bin")
    write_scene(sf, rec, rc.Camera._c(), 80, 60, 0, struct.pack("<i", 1) + bytes(lt._c()))
    run_program(exe, sf, of)
    got, _ = read_out(of)
    assert np.array_equal(got, np.asarray(img.Data, np.float32))
    # BidirPathTracer with the C5 parameters, adaptive fields zero, two sample chunks
    import bench
    spec = scenes.cornell_box()
    rec = []
    psc = scenes.build_product(spec, record=rec)
    bd = scenes.product_bidir(spec, psc, num_samples=16, seed=5, **bench.C5_KW)
    W = H = 48
    want, _, _ = bd.RenderSums(W, H, psc, sample_count=16)
    lights, nl = R._area_lights(psc, bd.Light)
    tail = bytes(bd._params(16)) + struct.pack("<i", nl) + bytes(lights)[:nl * C.sizeof(N.AreaLight)] + struct.pack("<ii", 16, 0)
    sf, of = str(tmp_path / "bd.bin"), str(tmp_path / "bd_out.bin")
    write_scene(sf, rec, bd.Camera._c(), W, H, 2, tail)
    run_program(exe, sf, of, chunks=2)
    got, _ = read_out(of)
    assert np.allclose(got, want, rtol=2e-5, atol=1e-4)


@pytest.mark.gpu
def test_c_program_on_a_multi_device_context(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    exe = build_program(tmp_path)
    spec = scenes.cornell_box()
    rec = []
    psc = scenes.build_product(spec, record=rec)
    tr = scenes.product_tracer(spec, psc, 5, 40, cutoff=1e-4, antialias=1.0, seed=3)
    W = H = 64
    want, _, _ = tr.RenderSums(W, H, psc, sample_count=40)
    p = tr._params(psc, 40)
    tail = bytes(p) + struct.pack("<i", 0) + struct.pack("<ii", 40, 0)
    sf, of = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(sf, rec, tr.Camera._c(), W, H, 1, tail)
    out, _ = run_program(exe, sf, of, devices=torch.cuda.device_count())
    assert "devices %d" % torch.cuda.device_count() in out
    got, _ = read_out(of)
    assert np.allclose(got, want, rtol=2e-5, atol=1e-4)
