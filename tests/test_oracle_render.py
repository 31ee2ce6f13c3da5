"""Pins the oracle's render3d restatement (oracle/scene.hpp, oracle/render.hpp) with the
reference's own self-consistency properties (the reference holds no golden vectors for
this path):
  material sampler == density == BSDF integrals    render3d/material_test.go:117-176
  RefractMaterial asymmetry                        render3d/material_test.go:56-80
  BidirPathTracer == RecursiveRayTracer            render3d/bidir_test.go:12-65
CPU only."""
import math

import zlib

import numpy as np
import pytest

import scenes


def rand_unit(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def one_material_scene(oracle, spec):
    sc = scenes.build_oracle(dict(objects=[dict(kind="sphere", center=(0, 0, 0), radius=1.0, material=spec)]))
    return sc, sc.material_index(spec)


def source_color(src):
    # sourceColorFunc (material_test.go:118-124)
    return np.stack([src[:, 0] + 2 * src[:, 1] ** 2 + 3 * src[:, 2] ** 3, src[:, 2] - src[:, 0] + src[:, 1],
                     np.ones(len(src))], axis=1)


MATS = {
    "lambert": scenes.lambert(diffuse=(1.0, 0.9, 0.5)),
    "phong0": scenes.phong(0.0, specular=(1.0, 0.9, 0.5)),
    "phong0.5": scenes.phong(0.5, specular=(1.0, 0.9, 0.5)),
    "phong2": scenes.phong(2.0, specular=(1.0, 0.9, 0.5)),
    "phong2_diffuse": scenes.phong(2.0, specular=(1.0, 0.9, 0.5), diffuse=(0.3, 0.2, 0.5)),
}


@pytest.mark.parametrize("name", sorted(MATS))
def test_material_sampling(oracle, name):
    """testMaterialSampling (material_test.go:117-152): the uniform-sphere integral of
    BSDF*f equals the importance-sampled one, BSDF*f/SourceDensity, within 1 %."""
    sc, mi = one_material_scene(oracle, MATS[name])
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000 + 5)  # str hashes vary per process
    normal = rand_unit(rng, 1)[0]
    dest = rand_unit(rng, 1)[0]
    while dest @ normal < 0.1:
        dest = rand_unit(rng, 1)[0]
    n = 2_000_000
    src = rand_unit(rng, n)
    bsdf, _, _ = sc.material_eval_batch(mi, normal, src, dest)
    actual = (bsdf * source_color(src)).mean(axis=0)
    smp = sc.material_sample_source(mi, 99, normal, dest, n)
    bsdf, sd, _ = sc.material_eval_batch(mi, normal, smp, dest)
    assert (sd > 0).all()
    expected = (bsdf * source_color(smp) / sd[:, None]).mean(axis=0)
    assert np.linalg.norm(actual - expected) <= np.linalg.norm(actual) * 0.01, (actual, expected)


@pytest.mark.parametrize("spec", [scenes.lambert(diffuse=(1, 1, 1)), scenes.phong(1e5, specular=(1, 1, 1)),
                                  scenes.phong(0.0, diffuse=(1, 1, 1))], ids=["lambert", "mirror", "phong_diffuse"])
def test_material_energy_conservation(oracle, spec):
    """testMaterialEnergyConservation (material_test.go:154-176)."""
    sc, mi = one_material_scene(oracle, spec)
    rng = np.random.default_rng(3)
    normal = rand_unit(rng, 1)[0]
    dest = np.zeros(3)
    while abs(normal @ dest - 0.8) > 0.1:
        dest = rand_unit(rng, 1)[0]
    n = 1_000_000
    smp = sc.material_sample_source(mi, 1337, normal, dest, n)
    bsdf, sd, _ = sc.material_eval_batch(mi, normal, smp, dest)
    e = (np.abs(smp @ normal) * bsdf[:, 0] / sd).mean()
    assert abs(e - 1) < 1e-2, e


@pytest.mark.parametrize("specular", [0.0, 1.0])
def test_refract_material_asym(oracle, specular):
    """TestRefractMaterialAsym (material_test.go:56-80): a sampled destination always has
    non-zero dest density, source density and BSDF."""
    sc, mi = one_material_scene(oracle, scenes.refract(1.3, (1, 1, 1), specular=(specular,) * 3))
    rng = np.random.default_rng(1337)
    for i in range(300):
        normal, source = rand_unit(rng, 2)
        dest = sc.material_sample_dest(mi, 1000 + i, normal, source, 1)[0]
        bsdf, sd, dd = sc.material_eval(mi, normal, source, dest)
        assert dd != 0 and sd != 0 and bsdf[0] != 0


def _testing_scene_oracle(oracle):
    spec = scenes.testing_scene()
    sc = scenes.build_oracle(spec)
    cam = oracle.camera_at(spec["camera"]["src"], spec["camera"]["dst"], spec["camera"]["fov"])
    return spec, sc, cam


def test_bidir_matches_path_tracer(oracle):
    """TestBidirPathTracer (bidir_test.go:12-65): on testingScene, 4x4 pixels, the BDPT image
    equals the RecursiveRayTracer image within 0.02 per pixel (Euclidean over RGB)."""
    spec, sc, cam = _testing_scene_oracle(oracle)
    pp = oracle.PathParams()
    pp.max_depth, pp.num_samples, pp.min_samples, pp.max_stddev = 10, 100000, 1000, 0.0015
    pp.num_focus_points = 2
    for i, f in enumerate(spec["focus"]):
        pp.focus[i].kind = 1
        pp.focus[i].target[:] = f["target"]
        pp.focus[i].radius = f["radius"]
        pp.focus[i].prob = f["prob"]
        pp.focus[i].material_mask = 0xFFFFFFFFFFFFFFFF
    pp.seed = 5
    truth = sc.render_path(cam, [], pp, 4, 4, threads=4)["mean"]
    assert np.isfinite(truth).all() and truth.min() > 0.01

    lights = []
    for l in spec["area_lights"]:
        a = oracle.AreaLight()
        a.object = l["object"]
        a.emission[:] = l["emission"]
        lights.append(a)
    bp = oracle.BidirParams()
    bp.max_depth, bp.num_samples, bp.seed = 10, 60000, 11
    got = sc.render_bidir(cam, lights, bp, 4, 4, threads=4)["mean"]
    assert np.isfinite(got).all()
    d = np.linalg.norm(got - truth, axis=2)
    assert d.max() < 0.02, (d.max(), got[0, 0], truth[0, 0])
    # RoulettePath variant (bidir_test.go:62-64)
    bp.min_depth, bp.num_samples = 1, 150000
    got = sc.render_bidir(cam, lights, bp, 4, 4, threads=4)["mean"]
    d = np.linalg.norm(got - truth, axis=2)
    assert d.max() < 0.02, d.max()


def test_path_tracer_early_stop_quirk(oracle):
    """ray_renderer.go:134-150: with a convergence criterion the mean uses the loop index."""
    spec, sc, cam = _testing_scene_oracle(oracle)
    pp = oracle.PathParams()
    pp.max_depth, pp.num_samples, pp.min_samples, pp.max_stddev, pp.seed = 3, 2000, 10, 1e9, 2
    a = sc.render_path(cam, [], pp, 2, 2, threads=1)["mean"]
    assert np.isfinite(a).all()


def test_sample_around_uniform(oracle):
    """TestSampleAroundUniform (focus_point_test.go:10-47): the uniform-sphere average of a test
    function restricted to the cone equals its importance-sampled average with weights 1/density,
    within 1e-3 (the SphereFocusPoint sampler of the showcase scene)."""
    rng = np.random.default_rng(1337)
    direction = rand_unit(rng, 1)[0]
    min_cos = 0.83

    def f(c):
        out = np.stack([c[:, 0] * c[:, 1] - c[:, 2],
                        c[:, 2] * c[:, 2] - 0.7 * c[:, 1] * c[:, 1] + 0.3 * c[:, 0] * c[:, 0],
                        0.6 * c[:, 0] + 0.5 * c[:, 1] + 0.3 * c[:, 2]], axis=1)
        return np.where((c @ direction < min_cos)[:, None], 0.0, out)

    n = 4_000_000
    expected = f(rand_unit(rng, n)).mean(axis=0)
    smp, dens = oracle.sample_around_uniform(1337, min_cos, direction, n)
    assert np.abs(np.linalg.norm(smp, axis=1) - 1).max() < 1e-12
    assert (dens > 0).all() and np.allclose(dens, 2 / (1 - min_cos))
    actual = (f(smp) / dens[:, None]).mean(axis=0)
    # the importance-sampled side is nearly exact; the uniform side carries ~3e-4 of noise at 4M
    assert np.linalg.norm(actual - expected) < 1e-3, (actual, expected)
