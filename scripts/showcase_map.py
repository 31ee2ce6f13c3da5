import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import scenes
from test_reference_golden import read_png_rgb8, srgb_expand, GOLD
spec = scenes.showcase(); psc = scenes.build_product(spec); n = 1024
tr = scenes.product_tracer(spec, psc, 10, n, cutoff=1e-4, antialias=1.0, seed=23)
rgb, _, _ = tr.RenderSums(480, 320, psc, sample_count=n)
mean = np.clip(rgb.astype(np.float64) / n, 0, 1)
np.save('gpurun_out/showcase_gpu_mean.npy', mean.astype(np.float32))
ref8 = read_png_rgb8(os.path.join(GOLD, "ref_showcase_output.png")); lin = srgb_expand(ref8)
B = 16
lb = lin.reshape(320 // B, B, 480 // B, B, 3).mean(axis=(1, 3)); mb = mean.reshape(320 // B, B, 480 // B, B, 3).mean(axis=(1, 3))
rel = np.abs((mb - lb) / np.maximum(lb, 0.02)).max(axis=2)
np.set_printoptions(linewidth=250)
print((rel * 100).astype(int))
