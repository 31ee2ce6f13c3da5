"""Renders the C5 scene (cornell_box, BidirPathTracer at its literal parameters) at a small size with a
fixed seed and saves the per-pixel sums: two builds of the library (M3D_LIB=...) that only differ in the
order of float additions must agree to ~1e-5.   python scripts/bidir_image_dump.py OUT.npy [size] [spp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes
out = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 256
res = {}
for name, spec in (("cornell", scenes.cornell_box()), ("testing", scenes.testing_scene())):
    psc = scenes.build_product(spec)
    tr = scenes.product_bidir(spec, psc, num_samples=spp, seed=77, max_depth=10, min_depth=3, roulette_delta=0.2,
                              power_heuristic=2.0, antialias=1.0, cutoff=1e-4)
    rgb, sq, stats = tr.RenderSums(size, size, psc, sample_count=spp, variance=True)
    res[name] = rgb
    res[name + "_sq"] = sq
    print(name, "mean", float(rgb.mean()) / spp, "rays", stats["rays"])
np.savez(out, **res)
