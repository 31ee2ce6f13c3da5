"""The C-ABI library loads and exports every symbol include/m3d.h declares; without a GPU
every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "m3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m3d_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(built, s), "libm3dgpu.so does not export %s" % s


def test_python_symbol_list_matches_header():
    from model3d_b200 import _native
    assert sorted(_native.SYMBOLS) == declared_symbols()


def test_abi_version(built):
    assert built.m3d_abi_version() == 5


def test_struct_sizes_match_header(built):
    """ctypes mirrors must have the layout of the C structs (checked against a tiny C probe)."""
    import subprocess
    import tempfile
    from model3d_b200 import _native as N
    names = {"m3d_stats": N.Stats, "m3d_mesh_info": N.MeshInfo, "m3d_camera": N.Camera,
             "m3d_point_light": N.PointLight, "m3d_material_desc": N.MaterialDesc,
             "m3d_transform": N.Transform, "m3d_partition": N.Partition,
             "m3d_focus_point": N.FocusPoint, "m3d_path_params": N.PathParams,
             "m3d_area_light": N.AreaLight, "m3d_bidir_params": N.BidirParams}
    prog = '#include <stdio.h>\n#include "m3d.h"\nint main(){' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "p.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "p"),
                               os.path.join(td, "p.c")])
        out = subprocess.check_output([os.path.join(td, "p")]).decode()
    sizes = dict(l.split() for l in out.strip().splitlines())
    for n, cls in names.items():
        assert int(sizes[n]) == C.sizeof(cls), n
    # the oracle's ctypes mirrors share the same PODs
    from oracle import pyoracle as O
    for n, cls in {"m3d_camera": O.Camera, "m3d_point_light": O.PointLight,
                   "m3d_material_desc": O.MaterialDesc, "m3d_transform": O.Transform,
                   "m3d_focus_point": O.FocusPoint, "m3d_path_params": O.PathParams,
                   "m3d_area_light": O.AreaLight, "m3d_bidir_params": O.BidirParams}.items():
        assert int(sizes[n]) == C.sizeof(cls), n


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from model3d_b200 import _native as N
    with pytest.raises(N.M3DError) as ei:
        N.Context()
    assert ei.value.code == 3  # M3D_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package must never load, link or call oracle/ (tests, smoke, bench only)."""
    pkg = os.path.join(ROOT, "model3d_b200")
    banned = ("import oracle", "from oracle", "liboracle", "oracle/", "pyoracle", "orc_")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for b in banned:
                    assert b not in txt, "%s references the oracle (%s)" % (f, b)


def test_multi_context_and_shared_memory_fail_loudly_without_gpu(built):
    """m3d_ctx_create_multi / m3d_host_alloc have no CPU stand-in either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from model3d_b200 import _native as N
    with pytest.raises(N.M3DError) as ei:
        N.MultiContext()
    assert ei.value.code == 3
    p = C.c_void_p()
    assert built.m3d_host_alloc(C.c_int64(4096), C.byref(p)) != 0
    assert built.m3d_ctx_num_devices(None) == 0
