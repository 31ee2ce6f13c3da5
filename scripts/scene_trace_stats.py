"""Nodes / triangle tests per ray of the example scenes (primary camera rays and random interior
rays, the secondary-ray distribution of the path tracers): python scripts/scene_trace_stats.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from model3d_b200 import examples, render3d as R

rng = np.random.default_rng(3)
for name, spec in (("cornell_box (C3/C5)", examples.cornell_box()), ("showcase (C4)", examples.showcase(hd=True))):
    psc = examples.build_product(spec)
    cam = spec["camera"]
    c = R.NewCameraAt(cam["src"], cam["dst"], cam["fov"])
    W, H = (960, 640) if "showcase" in name else (1024, 1024)
    d = R.CasterRays(c, W, H).astype(np.float32)
    o = np.tile(np.asarray(cam["src"], np.float32), (d.shape[0], 1))
    r = psc.Cast(o, d, counters=True)
    n = d.shape[0]
    print("%s primary: nodes/ray %.2f tris/ray %.2f hit %.3f kernel %.3f ms (%.2f Grays/s)" % (
        name, r["stats"]["nodes_visited"] / n, r["stats"]["tris_tested"] / n, (r["obj"] >= 0).mean(),
        r["stats"]["kernel_ms"], n / r["stats"]["kernel_ms"] / 1e6), flush=True)
    # secondary rays: from the primary hit points in cosine-ish random directions
    hit = r["obj"] >= 0
    p = (o + d * r["t"][:, None])[hit]
    nrm = r["normal"][hit]
    v = rng.normal(size=p.shape).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v = np.where(((v * nrm).sum(1) < 0)[:, None], -v, v)
    p = p + nrm * 1e-3
    r2 = psc.Cast(p, v, counters=True)
    n2 = p.shape[0]
    print("%s secondary: nodes/ray %.2f tris/ray %.2f hit %.3f kernel %.3f ms (%.2f Grays/s)" % (
        name, r2["stats"]["nodes_visited"] / n2, r2["stats"]["tris_tested"] / n2, (r2["obj"] >= 0).mean(),
        r2["stats"]["kernel_ms"], n2 / r2["stats"]["kernel_ms"] / 1e6), flush=True)
