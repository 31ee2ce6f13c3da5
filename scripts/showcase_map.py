"""Block-error map of our showcase HD frame against the reference's committed output_hd.png
(tuning aid for tests/test_reference_golden.py): python scripts/showcase_map.py [spp]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import scenes
from test_reference_golden import read_png_rgb8, srgb_expand, GOLD

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
spec = scenes.showcase(hd=True)
psc = scenes.build_product(spec)
tr = scenes.product_tracer(spec, psc, 10, n, cutoff=1e-4, antialias=1.0, seed=31)
rgb, _, _ = tr.RenderSums(960, 640, psc, sample_count=n)
mean = np.clip(rgb.astype(np.float64) / n, 0, 1)
lin = srgb_expand(read_png_rgb8(os.path.join(GOLD, "ref_showcase_output_hd.png")))
B = 32
lb = lin.reshape(640 // B, B, 960 // B, B, 3).mean(axis=(1, 3))
mb = mean.reshape(640 // B, B, 960 // B, B, 3).mean(axis=(1, 3))
rel = np.abs((mb - lb) / np.maximum(lb, 0.02)).max(axis=2)
np.set_printoptions(linewidth=250)
print((rel * 100).astype(int))
mask = np.ones(rel.shape, bool)
mask[6:, 21:] = False
mask[14:19, 16:] = False
r = rel[mask]
print("median %.4f p90 %.4f max %.3f far %s" % (np.median(r), np.percentile(r, 90), r.max(), np.argwhere((rel > 0.2) & mask).tolist()))
print("global", lin[:, :672].mean(axis=(0, 1)), mean[:, :672].mean(axis=(0, 1)))
