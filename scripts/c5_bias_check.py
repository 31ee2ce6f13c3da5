"""Is there a systematic difference between the GPU BidirPathTracer and the float64 oracle at the
C5 parameters (MaxDepth 10, MinDepth 3, RouletteDelta 0.2, PowerHeuristic 2)?  Image means with
standard errors for: oracle BDPT, GPU BDPT, oracle path tracer, GPU path tracer (same scene, the
light as an emissive object), and the same at MaxDepth 6.
  python scripts/c5_bias_check.py [size] [gpu_spp] [oracle_spp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from oracle import pyoracle as O

size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
gpu_spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
ora_spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
W = H = size
spec = scenes.cornell_box()
osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
cam = spec["camera"]
ocam = O.camera_at(cam["src"], cam["dst"], cam["fov"])


def stats(mean, var_of_mean):
    return mean.mean(), np.sqrt(var_of_mean.sum()) / mean.size


for depth in (10, 6, 3):
    kw = dict(max_depth=depth, min_depth=3, roulette_delta=0.2, power_heuristic=2.0, antialias=1.0, cutoff=1e-4)
    for variant, kwv in (("C5 params", kw), ("no roulette", dict(kw, roulette_delta=0.0, min_depth=0)),
                         ("balance heuristic", dict(kw, power_heuristic=0.0))):
        bp, lights = scenes.oracle_bidir_params(spec, num_samples=ora_spp, seed=11, **kwv)
        r = osc.render_bidir(ocam, lights, bp, W, H, threads=O.hardware_threads())
        om, ose = stats(r["mean"], r["var_of_mean"])
        bd = scenes.product_bidir(spec, psc, num_samples=gpu_spp, seed=7, **kwv)
        rgb, sq, _ = bd.RenderSums(W, H, psc, sample_count=gpu_spp, variance=True)
        mean = rgb.astype(np.float64) / gpu_spp
        var = np.maximum(sq.astype(np.float64) / gpu_spp - mean * mean, 0) / (gpu_spp - 1)
        gm, gse = stats(mean, var)
        print("depth %2d %-18s BDPT  oracle %.5f +- %.5f   GPU %.5f +- %.5f   diff %+.2f%% (%.1f sigma)" % (
            depth, variant, om, ose, gm, gse, 100 * (gm - om) / om, (gm - om) / np.hypot(ose, gse)), flush=True)
    pp = scenes.oracle_path_params(spec, osc, depth, ora_spp * 2, cutoff=1e-4, antialias=1.0, seed=5)
    r = osc.render_path(ocam, [], pp, W, H, threads=O.hardware_threads())
    om, ose = stats(r["mean"], r["var_of_mean"])
    tr = scenes.product_tracer(spec, psc, depth, gpu_spp, cutoff=1e-4, antialias=1.0, seed=3)
    rgb, sq, _ = tr.RenderSums(W, H, psc, sample_count=gpu_spp, variance=True)
    mean = rgb.astype(np.float64) / gpu_spp
    var = np.maximum(sq.astype(np.float64) / gpu_spp - mean * mean, 0) / (gpu_spp - 1)
    gm, gse = stats(mean, var)
    print("depth %2d %-18s PATH  oracle %.5f +- %.5f   GPU %.5f +- %.5f   diff %+.2f%% (%.1f sigma)" % (
        depth, "", om, ose, gm, gse, 100 * (gm - om) / om, (gm - om) / np.hypot(ose, gse)), flush=True)
