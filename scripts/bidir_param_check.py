"""GPU BDPT vs the float64 oracle for unusual parameter sets: prints the mean z-score and the image means
(a bias grows with the sample count, a skew artefact of few samples shrinks).
  python scripts/bidir_param_check.py [n_ref] [n_gpu]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
from test_gpu_path import z_scores
from test_gpu_bidir import gpu_bidir, oracle_bidir
from oracle import pyoracle as O

n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
spec = scenes.cornell_box()
osc, psc = scenes.build_oracle(spec), scenes.build_product(spec)
W = H = 24
for kw in (dict(max_depth=4, max_light_depth=2, min_depth=2, power_heuristic=3.0),
           dict(max_depth=4, max_light_depth=2, min_depth=2, power_heuristic=2.0),
           dict(max_depth=4, max_light_depth=2, min_depth=2),
           dict(max_depth=5, max_light_depth=3, min_depth=2, power_heuristic=3.0),
           dict(max_depth=5, max_light_depth=3, min_depth=2)):
    ref = oracle_bidir(O, spec, osc, W, H, n_ref, **kw)
    mean, var, _ = gpu_bidir(spec, psc, W, H, n_gpu, **kw)
    z = z_scores(mean, var, ref["mean"], ref["var_of_mean"])
    tot_s = np.sqrt(var.sum() + ref["var_of_mean"].sum()) / mean.size
    print(kw, "mean z %.3f  |z|>3 %.4f  image mean gpu %.5f oracle %.5f (sigma %.5f)" % (
        z.mean(), (np.abs(z) > 3).mean(), mean.mean(), ref["mean"].mean(), tot_s), flush=True)
