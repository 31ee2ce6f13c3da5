"""GPU parity tests for the wavefront bidirectional path tracer (m3d_render_bidir behind
render3d.BidirPathTracer) against the float64 oracle (oracle/render.hpp Bidir) and against
the reference's own cross-check, BDPT == RecursiveRayTracer (render3d/bidir_test.go:12-65).
Statistical parity: per-pixel means within 3 sigma of the combined Monte-Carlo noise."""
import numpy as np
import pytest

import scenes
from test_gpu_path import check_statistical_parity

pytestmark = pytest.mark.gpu


def gpu_bidir(spec, psc, W, H, n, **kw):
    tr = scenes.product_bidir(spec, psc, num_samples=n, seed=21, **kw)
    rgb, sq, stats = tr.RenderSums(W, H, psc, sample_count=n, variance=True)
    mean = rgb.astype(np.float64) / n
    var = np.maximum(sq.astype(np.float64) / n - mean * mean, 0.0) * n / (n - 1)
    assert np.isfinite(mean).all()
    return mean, var / n, stats


def oracle_bidir(oracle, spec, osc, W, H, n, **kw):
    cam = spec["camera"]
    ocam = oracle.camera_at(cam["src"], cam["dst"], cam["fov"])
    bp, lights = scenes.oracle_bidir_params(spec, num_samples=n, seed=4, **kw)
    return osc.render_bidir(ocam, lights, bp, W, H, threads=8)


def test_bidir_testing_scene_vs_oracle_and_path_tracer(built, oracle):
    """TestBidirPathTracer (bidir_test.go:12-65): testingScene, MaxDepth 10, two sphere area
    lights; plus the RoulettePath variant (MinDepth 1)."""
    spec = scenes.testing_scene()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    W = H = 8
    ref = oracle_bidir(oracle, spec, osc, W, H, 20000, max_depth=10)
    mean, var, stats = gpu_bidir(spec, psc, W, H, 40000, max_depth=10)
    check_statistical_parity(mean, var, ref["mean"], ref["var_of_mean"], block=2)
    # ground truth from the (oracle) path tracer with focus points, as the reference test does
    cam = spec["camera"]
    ocam = oracle.camera_at(cam["src"], cam["dst"], cam["fov"])
    pp = scenes.oracle_path_params(spec, osc, 10, 60000, seed=8)
    truth = osc.render_path(ocam, [], pp, W, H, threads=8)["mean"]
    assert np.linalg.norm(mean - truth, axis=2).max() < 0.02
    mean2, var2, _ = gpu_bidir(spec, psc, W, H, 100000, max_depth=10, min_depth=1)
    assert np.linalg.norm(mean2 - truth, axis=2).max() < 0.02


def test_bidir_cornell_box_mesh_light(built, oracle):
    """BASELINE config 5 parameters (MinDepth 3, RouletteDelta 0.2, PowerHeuristic 2, Antialias 1,
    Cutoff 1e-4; mesh area light = the ceiling light) at a test-sized resolution and depth."""
    spec = scenes.cornell_box()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    kw = dict(max_depth=6, min_depth=3, roulette_delta=0.2, power_heuristic=2.0, antialias=1.0, cutoff=1e-4)
    W = H = 32
    ref = oracle_bidir(oracle, spec, osc, W, H, 384, **kw)
    mean, var, stats = gpu_bidir(spec, psc, W, H, 1024, **kw)
    assert ref["mean"].mean() > 0.05
    check_statistical_parity(mean, var, ref["mean"], ref["var_of_mean"])
    assert stats["rays"] > W * H * 1024 * 2


def test_bidir_cornell_box_c5_literal_parameters(built, oracle):
    """BASELINE config 5 at its OWN parameters -- MaxDepth 10, MinDepth 3, RouletteDelta 0.2,
    PowerHeuristic 2, Antialias 1, Cutoff 1e-4 (bench.py C5_KW), the ceiling light as
    MeshAreaLight -- on a small frame: per-pixel statistical parity with the float64 oracle's
    restatement of bidir.go:101-576, and the same scene through the path tracer's estimator
    (the reference's own BDPT test compares the two, bidir_test.go:12-65)."""
    import bench
    assert bench.C5_KW["max_depth"] == 10 and bench.C5_KW["min_depth"] == 3 and bench.C5_KW["roulette_delta"] == 0.2
    spec = scenes.cornell_box()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    W = H = 40
    ref = oracle_bidir(oracle, spec, osc, W, H, 1024, **bench.C5_KW)
    mean, var, stats = gpu_bidir(spec, psc, W, H, 4096, **bench.C5_KW)
    assert ref["mean"].mean() > 0.05
    # Measured on B200 (scripts/c5_bias_check.py, 32x32, 32768 vs 4096 spp): the GPU BDPT image mean is
    # 0.9124 +- 0.0002, the oracle's BDPT 0.9108 +- 0.0006, the path tracers of both 0.9118 / 0.9120 --
    # i.e. GPU BDPT +0.07 % and oracle BDPT -0.13 % around the path-traced value, all far inside the
    # 0.02 the reference's own BDPT test allows (bidir_test.go:57-63).  At 1024 oracle samples that
    # 0.18 % shows up as a mean z of 0.17, hence the wider bound on the mean z here.
    check_statistical_parity(mean, var, ref["mean"], ref["var_of_mean"], max_mean_z=0.3)
    # eye sub-paths up to depth 10 and light sub-paths up to depth 10 really ran
    assert stats["rays"] > W * H * 4096 * 4
    # image means agree to Monte-Carlo accuracy (a biased deep-path weight would show here)
    assert abs(mean.mean() - ref["mean"].mean()) < 0.005 * ref["mean"].mean() + 4 * np.sqrt(var.sum() + ref["var_of_mean"].sum()) / mean.size


def test_bidir_glass_scene_dirac_lobes(built, oracle):
    """Specular chains (refraction + Fresnel reflection) through the float64 MIS weights:
    the Dirac magnitudes (2e8 per specular vertex) must cancel exactly as in the reference."""
    spec = scenes.glass_scene()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    kw = dict(max_depth=7, min_depth=3, power_heuristic=2.0, cutoff=1e-3)
    W, H = 32, 24
    ref = oracle_bidir(oracle, spec, osc, W, H, 512, **kw)
    mean, var, _ = gpu_bidir(spec, psc, W, H, 2048, **kw)
    check_statistical_parity(mean, var, ref["mean"], ref["var_of_mean"])


def test_bidir_balance_heuristic_and_light_depth(built, oracle):
    """PowerHeuristic 0 (sum of densities) and MaxLightDepth != MaxDepth."""
    spec = scenes.cornell_box()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    kw = dict(max_depth=5, max_light_depth=3, min_depth=2)
    W = H = 24
    ref = oracle_bidir(oracle, spec, osc, W, H, 384, **kw)
    mean, var, _ = gpu_bidir(spec, psc, W, H, 1024, **kw)
    check_statistical_parity(mean, var, ref["mean"], ref["var_of_mean"])


def test_bidir_general_power_heuristic_with_binding_depth_limits(built, oracle):
    """PowerHeuristic 3 (neither the balance sum nor the squared special case) with MaxDepth 4 /
    MaxLightDepth 2: most joined paths are cut by one of the two limits, which is where the per-sample
    MIS tables of the connection stage (H[j][t_lo], K[i][m_lo]) take their lower limits from."""
    spec = scenes.cornell_box()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    kw = dict(max_depth=4, max_light_depth=2, min_depth=2, power_heuristic=3.0)
    W = H = 24
    ref = oracle_bidir(oracle, spec, osc, W, H, 384, **kw)
    mean, var, _ = gpu_bidir(spec, psc, W, H, 1024, **kw)
    # two light vertices at most leave fewer strategies per path and heavier-tailed pixels: the mean z of
    # the skewed per-pixel estimates sits at 0.15-0.23 for 384 as for 8192 oracle samples while the image
    # means agree to 0.5 sigma (scripts/bidir_param_check.py), so the bound on it is wider here
    check_statistical_parity(mean, var, ref["mean"], ref["var_of_mean"], max_mean_z=0.3)


def test_bidir_partitions_add_up(built):
    spec = scenes.cornell_box()
    psc = scenes.build_product(spec)
    tr = scenes.product_bidir(spec, psc, 4, 32, min_depth=2, power_heuristic=2.0, antialias=1.0, seed=5)
    W, H = 24, 16
    full, _, _ = tr.RenderSums(W, H, psc)
    a, _, _ = tr.RenderSums(W, H, psc, partition=(0, 0, 0), sample_count=10)
    b, _, _ = tr.RenderSums(W, H, psc, partition=(0, 0, 10), sample_count=22)
    # visibility contributions are added with float atomics: order-dependent rounding only
    assert np.allclose(a + b, full, rtol=2e-3, atol=2e-3)


def test_bidir_errors(built):
    from model3d_b200 import render3d as R
    from model3d_b200 import _native as N
    spec = scenes.cornell_box()
    psc = scenes.build_product(spec)
    tr = scenes.product_bidir(spec, psc, 40, 4)
    with pytest.raises(N.M3DError) as ei:
        tr.RenderSums(4, 4, psc)
    assert ei.value.code == 2  # depth above the GPU limit: UNSUPPORTED, no fallback
    tr = scenes.product_bidir(spec, psc, 4, 4)
    tr.Light = R.NewSphereAreaLight(R.Sphere((0, 0, 0), 1.0), (1, 1, 1))
    with pytest.raises(ValueError):
        tr.RenderSums(4, 4, psc)  # the light is not part of the scene


def test_bidir_adaptive_sampling(built, oracle):
    """TestBidirPathTracer's own configuration (bidir_test.go:36-43): NumSamples 200000,
    MinSamples 1000, MaxStddev 0.0015 -- early stop per pixel, image within 0.02 of the path
    tracer's ground truth like the reference test demands."""
    from model3d_b200 import render3d as R
    spec = scenes.testing_scene()
    osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
    cam = spec["camera"]
    ocam = oracle.camera_at(cam["src"], cam["dst"], cam["fov"])
    pp = scenes.oracle_path_params(spec, osc, 10, 60000, seed=8)
    truth = osc.render_path(ocam, [], pp, 4, 4, threads=4)["mean"]
    tr = scenes.product_bidir(spec, psc, 10, 200000, seed=2)
    tr.MinSamples, tr.MaxStddev = 1000, 0.0015
    img = R.Image(4, 4)
    stats = tr.Render(img, psc)
    assert np.isfinite(img.Data).all()
    assert np.linalg.norm(img.Data - truth, axis=2).max() < 0.02
    assert 16 * 1000 < stats["samples"] < 16 * 200000
