"""CPU check of the product's traversal logic: the __host__ __device__ core of the CUDA
kernel (model3d_b200/csrc/trace_core.cuh) and the host BVH builder are compiled with g++
(tests/emul) and compared with the oracle.  This does not replace the -m gpu parity tests;
it lets the kernel logic be verified where there is no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "emul")], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(HERE, "emul", "libemul.so"))
    L.emul_build.restype = C.c_void_p
    return L


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def emul_trace(L, tris32, org, d, refine=1):
    h = C.c_void_p(L.emul_build(P(tris32, C.c_float), C.c_int64(tris32.shape[0])))
    n = org.shape[0]
    t = np.zeros(n, np.float32)
    prim = np.zeros(n, np.int32)
    nrm = np.zeros((n, 3), np.float32)
    bary = np.zeros((n, 3), np.float32)
    cnt = np.zeros(2, np.int64)
    L.emul_trace(h, P(org, C.c_float), P(d, C.c_float), C.c_int64(n), P(t, C.c_float), P(prim, C.c_int32),
                 P(nrm, C.c_float), P(bary, C.c_float), P(cnt, C.c_int64), C.c_int(refine))
    L.emul_destroy(h)
    return dict(t=t, prim=prim, normal=nrm, bary=bary, nodes=int(cnt[0]), tris=int(cnt[1]))


def rays(rng, n, scale=1.0):
    o = (rng.normal(size=(n, 3)) * scale).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def compare(oracle, tris32, org, d, got, tie_rel=1e-6, t_rel=1e-5):
    ref = oracle.Collider(tris32).first_hits(org, d, threads=4)
    hit_o, hit_g = ref["prim"] >= 0, got["prim"] >= 0
    # hit/miss may differ only for grazing hits; none expected on these inputs
    assert (hit_o != hit_g).sum() <= 1e-4 * len(hit_o)
    both = hit_o & hit_g
    same = both & (ref["prim"] == got["prim"])
    rel = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
    assert rel[same].max(initial=0) < t_rel
    diff = both & ~same
    assert rel[diff].max(initial=0) < tie_rel * 4, "different triangle that is not a tie"
    ndot = (got["normal"][same] * ref["normal"][same]).sum(1)
    assert ndot.min(initial=1) > 1 - 1e-5
    return ref


@pytest.mark.parametrize("mesh", ["polar10", "polar100", "rect", "ico32"])
def test_emulated_kernel_matches_oracle(emul, oracle, mesh):
    rng = np.random.default_rng(42)
    tris = {"polar10": lambda: oracle.mesh_polar(0.5, 0.1, 10),
            "polar100": lambda: oracle.mesh_polar(0.5, 0.1, 100),
            "rect": lambda: oracle.mesh_rect((-1, -2, -3), (1, 2, 3)),
            "ico32": lambda: oracle.mesh_icosphere((0, 0, 0), 1, 32)}[mesh]().astype(np.float32)
    org, d = rays(rng, 20000)
    got = emul_trace(emul, tris, org, d)
    compare(oracle, tris, org, d, got)


def test_emulated_axis_aligned_rays(emul, oracle):
    """Zero direction components (bvh.go:328-333 special-cases rate == 0)."""
    rng = np.random.default_rng(3)
    tris = oracle.mesh_icosphere((0, 0, 0), 1, 16).astype(np.float32)
    n = 6000
    org = (rng.uniform(-1.5, 1.5, size=(n, 3))).astype(np.float32)
    d = np.zeros((n, 3), np.float32)
    d[np.arange(n), rng.integers(0, 3, n)] = rng.choice([-1.0, 1.0, 2.5], n)
    got = emul_trace(emul, tris, org, d)
    compare(oracle, tris, org, d, got)
    # negative zeros: the reciprocal's sign and the near / far plane choice must agree
    d[::2] = np.where(d[::2] == 0, np.float32(-0.0), d[::2])
    got = emul_trace(emul, tris, org, d)
    ref = compare(oracle, tris, org, d, got)
    assert (ref["prim"] >= 0).sum() > n // 10


def test_emulated_no_refine_within_tolerance(emul, oracle):
    rng = np.random.default_rng(9)
    tris = oracle.mesh_icosphere((0, 0, 0), 1, 32).astype(np.float32)
    org, d = rays(rng, 20000)
    got = emul_trace(emul, tris, org, d, refine=0)
    ref = oracle.Collider(tris).first_hits(org, d, threads=4)
    same = (ref["prim"] >= 0) & (ref["prim"] == got["prim"])
    rel = np.abs(got["t"] - ref["t"])[same] / np.abs(ref["t"][same])
    assert np.quantile(rel, 0.999) < 1e-5


def test_emulated_collect_hits_matches_oracle(emul, oracle):
    """collect_bvh_hits (Collider.RayCollisions with the hits delivered, collisions.go:263-273) on
    the product's wide BVH == the oracle's all-hits walk: same triangles per ray, t within 1e-5."""
    rng = np.random.default_rng(9)
    tris = (rng.normal(size=(2000, 1, 3)) + rng.normal(size=(2000, 3, 3)) * 0.3).astype(np.float32).reshape(-1, 9)
    org, d = rays(rng, 4000)
    h = C.c_void_p(emul.emul_build(P(tris, C.c_float), C.c_int64(tris.shape[0])))
    n, cap = org.shape[0], 64
    counts, counts2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
    t = np.zeros((n, cap), np.float32)
    prim = np.zeros((n, cap), np.int32)
    emul.emul_collect(h, P(org, C.c_float), P(d, C.c_float), C.c_int64(n), C.c_int(cap), P(counts, C.c_int32),
                      P(counts2, C.c_int32), P(t, C.c_float), P(prim, C.c_int32))
    emul.emul_destroy(h)
    assert np.array_equal(counts, counts2) and counts.max() < cap and counts.max() > 4
    ref = oracle.Collider(tris.reshape(-1, 3, 3)).all_hits_batch(org, d, threads=4)
    rc = np.diff(ref["offsets"])
    assert (rc != counts).sum() <= 1
    for i in np.nonzero(rc == counts)[0]:
        a0 = ref["offsets"][i]
        go, ro = np.argsort(prim[i, :counts[i]]), np.argsort(ref["prim"][a0:a0 + rc[i]])
        assert np.array_equal(prim[i, :counts[i]][go], ref["prim"][a0:a0 + rc[i]][ro])
        rt = ref["t"][a0:a0 + rc[i]][ro]
        # raw float32 t of the traversal (the kernel re-evaluates every hit in float64 afterwards)
        assert np.all(np.abs(t[i, :counts[i]][go] - rt) <= 1e-4 * np.maximum(np.abs(rt), 1.0))  # glancing hits are ill-conditioned in float32


def test_bounds_cull_predicate_is_conservative(emul, oracle):
    """ray_misses_bounds (trace_core.cuh), the test of the optional bounds cull in front of large mesh
    batches: whenever it retires a ray, the float64 oracle reports a miss -- for origins near and far,
    axis-parallel and signed-zero direction components, rays grazing the bounds, origins inside."""
    tris = oracle.mesh_icosphere((0.3, -0.2, 0.1), 1.0, 12).astype(np.float32)
    v = tris.reshape(-1, 3)
    bmin, bmax = v.min(0).astype(np.float32), v.max(0).astype(np.float32)
    rng = np.random.default_rng(11)
    sets = []
    for scale in (0.5, 2.0, 50.0, 1e4):
        o = (rng.normal(size=(20000, 3)) * scale).astype(np.float32)
        d = rng.normal(size=(20000, 3)).astype(np.float32)
        sets.append((o, d))
    # axis-parallel rays along the faces of the bounds (grazing) and signed zeros
    o = (rng.uniform(-3, 3, size=(20000, 3))).astype(np.float32)
    d = np.zeros((20000, 3), np.float32)
    ax = rng.integers(0, 3, size=20000)
    d[np.arange(20000), ax] = rng.choice([-1.0, 1.0], size=20000)
    d[::3] = np.where(d[::3] == 0, -0.0, d[::3])
    k = rng.integers(0, 3, size=20000)
    face = np.where(rng.random(20000) < 0.5, bmin[k], bmax[k])
    sel = k != ax
    o[sel, k[sel]] = face[sel]
    sets.append((o, d))
    # rays aimed at points on the bounds' surface from outside
    o = (rng.normal(size=(20000, 3)) * 4).astype(np.float32)
    tgt = rng.uniform(bmin, bmax, size=(20000, 3)).astype(np.float32)
    kk = rng.integers(0, 3, size=20000)
    tgt[np.arange(20000), kk] = np.where(rng.random(20000) < 0.5, bmin[kk], bmax[kk])
    sets.append((o, (tgt - o).astype(np.float32)))
    col = oracle.Collider(tris)
    culled_total = 0
    for o, d in sets:
        o, d = np.ascontiguousarray(o), np.ascontiguousarray(d)
        out = np.zeros(o.shape[0], np.uint8)
        emul.emul_ray_misses_bounds(P(o, C.c_float), P(d, C.c_float), C.c_int64(o.shape[0]), C.c_float(0.0),
                                    C.c_float(np.inf), P(bmin, C.c_float), P(bmax, C.c_float), P(out, C.c_uint8))
        ref = col.first_hits(o, d, threads=8)
        bad = (out == 1) & (ref["prim"] >= 0)
        assert not bad.any(), "the cull retired %d rays the oracle hits" % int(bad.sum())
        culled_total += int(out.sum())
    assert culled_total > 20000  # and it does retire rays
