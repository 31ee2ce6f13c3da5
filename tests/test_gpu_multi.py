"""Multi-GPU paths of the library: the shared-accumulator flush
(M3D_PART_ATOMIC: system-scope red.add into one frame accumulator, also through a CUDA IPC
mapping from another process) and the multi-device context (m3d_ctx_create_multi: replicated
meshes / scenes, one host thread per device, sample / row / ray-slice partition inside).

The single-GPU tests exercise the same kernels and the IPC mapping on one device; the tests that
need two devices skip on a one-GPU box (run them with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torch():
    import torch
    return torch


def _num_gpus():
    return _torch().cuda.device_count()


def _c3(ctx=None, spp=32, seed=11):
    spec = scenes.cornell_box()
    psc = scenes.build_product(spec, ctx=ctx) if ctx is not None else scenes.build_product(spec)
    tr = scenes.product_tracer(spec, psc, 5, spp, cutoff=1e-4, antialias=1.0, seed=seed)
    return spec, psc, tr


def _sums_device(tr, psc, W, H, spp, flags=0, sample_begin=0):
    torch = _torch()
    acc = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    tr.RenderSumsDevice(W, H, psc, acc.data_ptr(), partition=(0, 0, sample_begin, flags), sample_count=spp)
    torch.cuda.synchronize()
    return acc.cpu().numpy()


def test_atomic_flush_equals_plain_flush_single_batch():
    from model3d_b200 import _native as N
    _, psc, tr = _c3()
    W = H = 96
    plain = _sums_device(tr, psc, W, H, 32)
    red = _sums_device(tr, psc, W, H, 32, flags=N.PART_ATOMIC)
    assert np.array_equal(plain, red)


def test_atomic_flush_unaligned_rows_and_sumsq():
    """Row bands whose first pixel is not a multiple of four take the scalar red.add kernel."""
    from model3d_b200 import _native as N
    torch = _torch()
    _, psc, tr = _c3()
    W, H = 97, 40
    out = []
    for flags in (0, N.PART_ATOMIC):
        acc = torch.zeros((2, H, W, 3), dtype=torch.float32, device="cuda:0")
        for band in ((0, 13), (13, 40)):
            tr.RenderSumsDevice(W, H, psc, acc[0].data_ptr(), d_rgb_sumsq=acc[1].data_ptr(),
                                partition=(band[0], band[1], 0, flags), sample_count=16)
        torch.cuda.synchronize()
        out.append(acc.cpu().numpy())
    assert np.array_equal(out[0], out[1])
    assert out[0][1].sum() > 0


def test_atomic_flush_multi_batch_carry():
    """More path slots than one batch holds (2^26): the partial sums of the pixel range are
    carried in local memory and only the last batch adds to the shared accumulator."""
    from model3d_b200 import _native as N
    _, psc, tr = _c3()
    W = H = 512
    spp = 320  # 512 * 512 * 320 = 1.25 * 2^26 slots -> two sample batches per pixel range
    plain = _sums_device(tr, psc, W, H, spp)
    red = _sums_device(tr, psc, W, H, spp, flags=N.PART_ATOMIC)
    # same per-batch sums; only the association of the final adds differs
    assert np.allclose(plain, red, rtol=2e-6, atol=1e-5)
    assert abs(plain.sum() - red.sum()) <= 1e-6 * plain.sum()


def test_two_shards_into_one_accumulator_equal_the_whole():
    """Sample shards flushed with red.add into one buffer == the unsharded render (the Philox
    stream is keyed by the absolute sample index)."""
    from model3d_b200 import _native as N
    torch = _torch()
    _, psc, tr = _c3()
    W = H = 128
    whole = _sums_device(tr, psc, W, H, 48)
    acc = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
    tr.RenderSumsDevice(W, H, psc, acc.data_ptr(), partition=(0, 0, 0, N.PART_ATOMIC), sample_count=20)
    tr.RenderSumsDevice(W, H, psc, acc.data_ptr(), partition=(0, 0, 20, N.PART_ATOMIC), sample_count=28)
    torch.cuda.synchronize()
    assert np.allclose(whole, acc.cpu().numpy(), rtol=2e-6, atol=1e-5)


def test_ipc_accumulator_shared_between_processes():
    """One process per GPU (torchrun): rank 0 exports its accumulator, another process maps it
    with m3d_ipc_open and flushes its sample shard into it.  Here both processes use GPU 0."""
    from model3d_b200 import _native as N
    ctx = N.default_context(0)
    _, psc, tr = _c3()
    W = H = 128
    whole = _sums_device(tr, psc, W, H, 48)
    nbytes = W * H * 3 * 4
    acc = N.device_alloc(ctx, nbytes)
    try:
        handle = N.ipc_export(ctx, acc)
        tr.RenderSumsDevice(W, H, psc, acc, partition=(0, 0, 0, N.PART_ATOMIC), sample_count=20)
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ipc_worker.py"), handle.hex(),
                              str(W), str(H), "20", "28"], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        from model3d_b200 import distributed as D
        torch = _torch()
        torch.cuda.synchronize()
        got = torch.as_tensor(D.DevicePointer(acc, (H, W, 3)), device="cuda:0").clone()
        assert np.allclose(whole, got.cpu().numpy(), rtol=2e-6, atol=1e-5)
    finally:
        N.device_free(ctx, acc)


def test_pinned_host_arrays():
    from model3d_b200 import _native as N, MeshCollider
    from model3d_b200 import meshes
    tris = meshes.NewMeshIcosphere((0, 0, 0), 1.0, 8).astype(np.float32).reshape(-1, 9)
    col = MeshCollider(tris)
    rng = np.random.default_rng(5)
    n = 50000
    org = N.host_empty((n, 3), np.float32)
    d = N.host_empty((n, 3), np.float32)
    org[:] = rng.normal(size=(n, 3)) * 2
    d[:] = rng.normal(size=(n, 3))
    a = col.FirstRayCollisions(org, d)
    b = col.FirstRayCollisions(np.array(org), np.array(d))  # pageable copies
    assert np.array_equal(a.Triangle, b.Triangle) and np.array_equal(a.Scale, b.Scale)
    # registering memory the caller already owns
    import ctypes as C
    own = np.zeros((n, 3), np.float32)
    N.check(N.lib().m3d_host_register(C.c_void_p(own.ctypes.data), C.c_int64(own.nbytes)))
    N.check(N.lib().m3d_host_unregister(C.c_void_p(own.ctypes.data)))
    del org, d


# ---- two or more devices -----------------------------------------------------------------------

needs2 = pytest.mark.skipif("_num_gpus() < 2", reason="needs two GPUs (gpurun --gpus 2)")


@needs2
def test_multi_context_first_hits_equal_single_device():
    from model3d_b200 import _native as N, MeshCollider, meshes
    tris = meshes.NewMeshIcosphere((0, 0, 0), 1.0, 24).astype(np.float32).reshape(-1, 9)
    rng = np.random.default_rng(9)
    n = 300001
    org = rng.normal(size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    one = MeshCollider(tris, ctx=N.Context(0)).FirstRayCollisions(org, d)
    mctx = N.MultiContext(list(range(_num_gpus())))
    assert mctx.num_devices == _num_gpus()
    many = MeshCollider(tris, ctx=mctx).FirstRayCollisions(org, d)
    assert np.array_equal(one.Triangle, many.Triangle)
    assert np.array_equal(one.Scale, many.Scale)
    assert np.array_equal(one.Normal, many.Normal)


@needs2
def test_multi_context_path_render_equals_single_device():
    from model3d_b200 import _native as N
    W = H = 160
    _, psc1, tr1 = _c3(ctx=N.Context(0), spp=50)
    one, _, st1 = tr1.RenderSums(W, H, psc1, sample_count=50)
    mctx = N.MultiContext(list(range(_num_gpus())))
    _, pscm, trm = _c3(ctx=mctx, spp=50)
    many, _, stm = trm.RenderSums(W, H, pscm, sample_count=50)
    assert stm["samples"] == st1["samples"] == W * H * 50
    assert stm["rays"] == st1["rays"]  # same paths, whatever the partition
    assert np.allclose(one, many, rtol=3e-6, atol=2e-5)


@needs2
def test_multi_context_bidir_and_adaptive():
    from model3d_b200 import _native as N
    W = H = 64
    spec = scenes.cornell_box()
    res = []
    for ctx in (N.Context(0), N.MultiContext(list(range(_num_gpus())))):
        psc = scenes.build_product(spec, ctx=ctx)
        bd = scenes.product_bidir(spec, psc, 4, 24, min_depth=2, roulette_delta=0.2, power_heuristic=2.0,
                                  cutoff=1e-4, antialias=1.0, seed=5)
        rgb, _, st = bd.RenderSums(W, H, psc, sample_count=24)
        tr = scenes.product_tracer(spec, psc, 5, 64, cutoff=1e-4, antialias=1.0, seed=3)
        tr.MinSamples, tr.MaxStddev = 8, 0.05
        ad, _, sta = tr.RenderSums(W, H, psc)
        res.append((rgb, ad, st, sta))
    assert np.allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-4)
    assert np.allclose(res[0][1], res[1][1], rtol=1e-5, atol=1e-4)  # adaptive: row bands, same per-pixel stops
    assert res[0][3]["samples"] == res[1][3]["samples"]


@needs2
def test_multi_context_raycast_large_frame():
    from model3d_b200 import _native as N, render3d as R
    spec = scenes.c1_scene(n=16)
    cam = spec["camera"]
    lt = spec["lights"][0]
    imgs = []
    for ctx in (N.Context(0), N.MultiContext(list(range(_num_gpus())))):
        psc = scenes.build_product(spec, ctx=ctx)
        rc = R.RayCaster(Camera=R.NewCameraAt(cam["src"], cam["dst"], cam["fov"]),
                         Lights=[R.PointLight(lt["origin"], lt["color"])])
        img = R.Image(2048, 1536)  # >= 2^21 pixels: row bands over the devices
        rc.Render(img, psc)
        imgs.append(np.array(img.Data))
    assert np.array_equal(imgs[0], imgs[1])
    assert imgs[0].sum() > 0


@needs2
def test_multi_context_render_views_spread_over_devices():
    from model3d_b200 import _native as N, helpers as H, render3d as R
    spec = scenes.c1_scene(n=10)
    outs = []
    for ctx in (N.Context(0), N.MultiContext(list(range(_num_gpus())))):
        psc = scenes.build_product(spec, ctx=ctx)
        g = H.SaveRandomGrid(None, psc, 3, 3, 48, seed=4)
        outs.append(np.array(g.Data))
    assert np.array_equal(outs[0], outs[1]) and outs[0].sum() > 0
