"""Scene specs shared by tests: model3d_b200.examples holds the neutral specs and the product
builder; this module adds the second builder, for the CPU oracle (oracle/pyoracle.py)."""
import numpy as np

from model3d_b200.examples import *  # noqa: F401,F403
from model3d_b200.examples import all_materials


def build_oracle(spec):
    from oracle import pyoracle as O
    sc = O.Scene()
    cache = {}

    def mat(m):
        if id(m) in cache:
            return cache[id(m)]
        d = O.MaterialDesc()
        if m["kind"] == "lambert":
            d.kind = 0
            d.diffuse[:], d.ambient[:], d.emission[:] = m["diffuse"], m["ambient"], m["emission"]
        elif m["kind"] == "phong":
            d.kind = 1
            d.alpha = m["alpha"]
            d.specular[:], d.diffuse[:], d.ambient[:], d.emission[:] = m["specular"], m["diffuse"], m["ambient"], m["emission"]
        elif m["kind"] == "refract":
            d.kind = 2
            d.index_of_refraction = m["ior"]
            d.refract[:], d.specular[:] = m["refract"], m["specular"]
        elif m["kind"] == "joined":
            d.kind = 3
            d.num_sub = len(m["mats"])
            for i, s in enumerate(m["mats"]):
                d.sub[i] = mat(s)
                d.sub_prob[i] = m["probs"][i]
        elif m["kind"] == "checker":
            d.kind = 0
            d.flags = 2
            d.diffuse[:], d.diffuse2[:] = m["color1"], m["color2"]
        elif m["kind"] == "zgradient":
            d.kind = 1
            d.flags = 4
            d.alpha = m["alpha"]
            d.specular[:], d.diffuse[:], d.diffuse2[:] = m["specular"], m["color1"], m["color2"]
            d.proc_param = m["max_z"]
        idx = sc.add_material(d)
        cache[id(m)] = idx
        return idx

    for o in spec["objects"]:
        mi = mat(o["material"])
        flags = 1 if o.get("flip") else 0
        xf = o.get("xf")
        if o["kind"] == "mesh":
            vn = o.get("vnormals")
            sc.add_mesh(np.asarray(o["tris"], np.float32), mi, vnormals=None if vn is None else np.asarray(vn, np.float32),
                        flags=flags, xf=xf)
        elif o["kind"] == "sphere":
            sc.add_sphere(o["center"], o["radius"], mi, flags=flags, xf=xf)
        elif o["kind"] == "rect":
            sc.add_rect(o["min"], o["max"], mi, flags=flags, xf=xf)
        elif o["kind"] == "cylinder":
            sc.add_cylinder(o["p1"], o["p2"], o["radius"], mi, flags=flags, xf=xf)
    sc.material_index = lambda m: cache[id(m)]
    return sc


def oracle_path_params(spec, osc, max_depth, num_samples, cutoff=0.0, antialias=0.0, seed=1):
    from oracle import pyoracle as O
    pp = O.PathParams()
    pp.max_depth, pp.num_samples, pp.cutoff, pp.antialias, pp.seed = max_depth, num_samples, cutoff, antialias, seed
    focus = spec.get("focus", [])
    pp.num_focus_points = len(focus)
    for i, f in enumerate(focus):
        pp.focus[i].kind = 0 if f["kind"] == "phong" else 1
        pp.focus[i].target[:] = f["target"]
        pp.focus[i].alpha = f.get("alpha", 0.0)
        pp.focus[i].radius = f.get("radius", 0.0)
        pp.focus[i].prob = f["prob"]
        mask = 0
        for m in all_materials(spec):
            if f["applies"](m):
                mask |= 1 << osc.material_index(m)
        pp.focus[i].material_mask = mask
    return pp




def oracle_bidir_params(spec, max_depth, num_samples, min_depth=0, roulette_delta=0.0, power_heuristic=0.0,
                        cutoff=0.0, antialias=0.0, max_light_depth=0, seed=1):
    from oracle import pyoracle as O
    bp = O.BidirParams()
    bp.max_depth, bp.max_light_depth, bp.min_depth, bp.num_samples = max_depth, max_light_depth, min_depth, num_samples
    bp.roulette_delta, bp.power_heuristic, bp.cutoff, bp.antialias, bp.seed = (
        roulette_delta, power_heuristic, cutoff, antialias, seed)
    lights = []
    for l in spec["area_lights"]:
        a = O.AreaLight()
        a.object = l["object"]
        a.emission[:] = l["emission"]
        lights.append(a)
    return bp, lights
