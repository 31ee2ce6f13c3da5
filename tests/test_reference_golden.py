"""Pins against OUTPUTS OF THE REFERENCE ITSELF.

The reference ships no golden vectors for this path, but it does commit renderings that its own
Go code produced: examples/renderings/cornell_box/output.png (the exact configuration of
cornell_box/main.go: 200x200, MaxDepth 5, NumSamples 400, Antialias 1, Cutoff 1e-4,
PhongFocusPoint prob 0.3) and output_hd.png (500x500; README.md: "MaxDepth to maybe 15, NumSamples
to 20000").  They are byte-copied to tests/golden/ (make_fixtures.py).  The scene exercises the whole
path: BVH first hits on the diamond mesh and the room rectangles, analytic spheres, Lambert / Phong /
refractive / joined materials, the focus point, the recursive tracer.

The comparison is statistical (the reference's random stream cannot be reproduced): every 8-bit
sRGB pixel is expanded to linear light and compared with our estimate; the reference's own noise
is modelled from our per-pixel sample variance at its sample count, plus ours, plus quantisation.
A correct restatement gives z-scores of mean 0 and rms 1 and equal image means."""
import os
import struct
import zlib

import numpy as np
import pytest

import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_png_rgb8(path):
    """Minimal 8-bit RGB(A) / palette PNG decoder (all five filter types)."""
    d = open(path, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, plte = 8, b"", None
    while pos < len(d):
        n, typ = struct.unpack(">I4s", d[pos:pos + 8])
        body = d[pos + 8:pos + 8 + n]
        pos += 12 + n
        if typ == b"IHDR":
            W, H, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert depth == 8 and ctype in (2, 3, 6)
        elif typ == b"PLTE":
            plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif typ == b"IDAT":
            idat += body
    bpp = {2: 3, 3: 1, 6: 4}[ctype]
    raw = zlib.decompress(idat)
    stride = W * bpp
    out = np.zeros((H, stride), np.int64)
    prev = np.zeros(stride, np.int64)
    for y in range(H):
        f = raw[y * (stride + 1)]
        line = np.frombuffer(raw[y * (stride + 1) + 1:(y + 1) * (stride + 1)], np.uint8).astype(np.int64)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:
            cur = np.zeros(stride, np.int64)
            for x in range(stride):
                a = cur[x - bpp] if x >= bpp else 0
                b = prev[x]
                c = prev[x - bpp] if x >= bpp else 0
                if f == 1:
                    p = a
                elif f == 3:
                    p = (a + b) // 2
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + p) & 255
        out[y] = cur
        prev = cur
    if ctype == 3:
        return plte[out.reshape(H, W)]
    return out.reshape(H, W, bpp)[:, :, :3].astype(np.uint8)


def srgb_expand(u8):
    """Inverse of render3d's gammaCompress (light.go:41-55) on 8-bit values (image.go:125-145)."""
    u = u8.astype(np.float64) / 255.0
    return np.where(u <= 0.04045, u / 12.92, ((u + 0.055) / 1.055) ** 2.4)


def compare(ref8, mean, var_of_mean, n_ours, n_ref):
    lin = srgb_expand(ref8)
    ok = (ref8 < 250).all(axis=2) & (mean < 0.95).all(axis=2)  # clamped (light source) pixels carry no information
    var_pix = var_of_mean * n_ours
    sigma = np.sqrt(var_pix / n_ref + var_of_mean + (0.5 / 255) ** 2)
    z = ((lin - mean) / np.maximum(sigma, 1e-4))[ok]
    rel = (lin[ok].mean(axis=0) - mean[ok].mean(axis=0)) / mean[ok].mean(axis=0)
    return dict(used=ok.mean(), mean_z=z.mean(), rms_z=np.sqrt((z ** 2).mean()), frac3=(np.abs(z) > 3).mean(),
                rel=np.abs(rel).max())


def cornell(oracle):
    spec = scenes.cornell_box()
    cam = spec["camera"]
    return spec, oracle.camera_at(cam["src"], cam["dst"], cam["fov"])


def test_oracle_matches_the_references_own_rendering(oracle):
    ref8 = read_png_rgb8(os.path.join(GOLD, "ref_cornell_box_output.png"))
    assert ref8.shape == (200, 200, 3)
    spec, ocam = cornell(oracle)
    osc = scenes.build_oracle(spec)
    n = 600
    pp = scenes.oracle_path_params(spec, osc, 5, n, cutoff=1e-4, antialias=1.0, seed=11)
    r = osc.render_path(ocam, [], pp, 200, 200, threads=8)
    c = compare(ref8, r["mean"], r["var_of_mean"], n, 400)
    assert c["used"] > 0.9
    assert abs(c["mean_z"]) < 0.15, c
    assert 0.85 < c["rms_z"] < 1.3, c
    assert c["frac3"] < 0.03, c
    assert c["rel"] < 0.004, c  # image means agree to a fraction of a percent


@pytest.mark.gpu
def test_gpu_matches_the_references_own_rendering(built):
    ref8 = read_png_rgb8(os.path.join(GOLD, "ref_cornell_box_output.png"))
    spec = scenes.cornell_box()
    psc = scenes.build_product(spec)
    n = 4096
    tr = scenes.product_tracer(spec, psc, 5, n, cutoff=1e-4, antialias=1.0, seed=21)
    rgb, sq, _ = tr.RenderSums(200, 200, psc, sample_count=n, variance=True)
    mean = rgb.astype(np.float64) / n
    vom = np.maximum(sq.astype(np.float64) / n - mean * mean, 0.0) / (n - 1)
    c = compare(ref8, mean, vom, n, 400)
    assert abs(c["mean_z"]) < 0.15, c
    assert 0.85 < c["rms_z"] < 1.3, c
    assert c["frac3"] < 0.03, c
    assert c["rel"] < 0.004, c


@pytest.mark.gpu
def test_gpu_matches_the_references_hd_rendering(built):
    """output_hd.png: 20000 spp leave almost no noise, so this is a bias test at the 8-bit
    quantisation level; MaxDepth is only documented as "maybe 15"."""
    ref8 = read_png_rgb8(os.path.join(GOLD, "ref_cornell_box_output_hd.png"))
    assert ref8.shape == (500, 500, 3)
    spec = scenes.cornell_box()
    psc = scenes.build_product(spec)
    n = 16384
    tr = scenes.product_tracer(spec, psc, 15, n, cutoff=1e-4, antialias=1.0, seed=22)
    rgb, sq, _ = tr.RenderSums(500, 500, psc, sample_count=n, variance=True)
    mean = rgb.astype(np.float64) / n
    vom = np.maximum(sq.astype(np.float64) / n - mean * mean, 0.0) / (n - 1)
    c = compare(ref8, mean, vom, n, 20000)
    # measured on B200: mean z -0.17, rms z 1.27, channel means within 0.8 % (MaxDepth 10 / 15 / 20
    # give the same numbers; the exact settings of this image are not recorded upstream)
    assert c["rel"] < 0.012, c
    assert abs(c["mean_z"]) < 0.4 and c["rms_z"] < 1.6, c
    # 8-bit agreement after our own gamma compression (light.go:41-47, image.go:136-141): both
    # images carry ~1/255 of noise in the dark regions, where the sRGB curve is steepest
    lin = np.clip(mean, 0, 1)
    ours8 = np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055) * 255.999
    d8 = np.abs(ours8.astype(np.int64) - ref8.astype(np.int64)).max(axis=2)
    ok = (ref8 < 250).all(axis=2)
    assert (d8[ok] <= 3).mean() > 0.85 and (d8[ok] <= 5).mean() > 0.94, ((d8[ok] <= 3).mean(), (d8[ok] <= 5).mean())


# ---- showcase (BASELINE config C4): examples/renderings/showcase/output.png -----------------------
def _check_showcase(mean, hd=False):
    """mean: linear 320x480x3 estimate of showcase/main.go's frame (480x320, MaxDepth 10, Antialias 1,
    Cutoff 1e-4, SphereFocusPoint 0.3).  The reference image was committed as a 244-colour palette
    PNG at 50 spp, so single pixels carry several 8-bit levels of quantisation on top of the noise:
    16x16 block means in linear light are compared.  The vase (mesh missing from the checkout) and
    the floor / base it shades are masked."""
    ref8 = read_png_rgb8(os.path.join(GOLD, "ref_showcase_output_hd.png" if hd else "ref_showcase_output.png"))
    k = 2 if hd else 1  # output_hd.png: 960x640, same framing, 32x32 blocks
    assert ref8.shape == (320 * k, 480 * k, 3)
    lin, ours = srgb_expand(ref8), np.clip(mean, 0, 1)
    B = 16 * k
    lb = lin.reshape(320 * k // B, B, 480 * k // B, B, 3).mean(axis=(1, 3))
    mb = ours.reshape(320 * k // B, B, 480 * k // B, B, 3).mean(axis=(1, 3))
    mask = np.ones(lb.shape[:2], bool)
    mask[6:, 21:] = False    # the vase and its shadow (x >= 336, y >= 96)
    mask[14:19, 16:] = False  # the base of the curvy thing, shaded by the vase in the reference
    rel_map = np.abs((mb - lb) / np.maximum(lb, 0.02)).max(axis=2)
    rel = rel_map[mask]
    assert mask.sum() >= 400
    # measured: oracle at 64 spp median 1.2 %, 90th percentile 3.5 %; GPU at 1024 spp 1.3 % / 4.4 %
    assert np.median(rel) < 0.02, np.median(rel)
    assert np.percentile(rel, 90) < 0.06, np.percentile(rel, 90)
    # The only blocks far off are caustics in, under and beside the wine glass (x < 192): single
    # 50-sample frames are fireflies there (the reference is one such frame); everywhere else every
    # block agrees within 20 %.
    far = (rel_map > 0.2) & mask
    assert far[:, 12:].sum() == 0, np.argwhere(far[:, 12:])
    assert far.sum() <= 10, far.sum()
    g_ref, g_ours = lin[:, :336 * k].mean(axis=(0, 1)), ours[:, :336 * k].mean(axis=(0, 1))
    assert np.abs(g_ours / g_ref - 1).max() < 0.025, (g_ref, g_ours)  # measured +1.3 % (no vase)


# The reference clamps each pixel's 50-sample mean to [0, 1] (image.go:125-145); under the glass and
# on the specular highlights those means are heavy tailed, so E[clip(mean_50)] < clip(E[mean]).  Both
# tests therefore average frames rendered at exactly 50 spp and clamped one by one -- the
# reference's own estimator -- instead of clamping one converged frame (which reads 3 % brighter).
def test_oracle_matches_the_references_showcase_rendering(oracle):
    spec = scenes.showcase()
    osc = scenes.build_oracle(spec)
    cam = spec["camera"]
    ocam = oracle.camera_at(cam["src"], cam["dst"], cam["fov"])
    acc = np.zeros((320, 480, 3))
    for seed in (7, 8):
        pp = scenes.oracle_path_params(spec, osc, 10, 50, cutoff=1e-4, antialias=1.0, seed=seed)
        acc += np.clip(osc.render_path(ocam, [], pp, 480, 320, threads=8)["mean"], 0, 1)
    _check_showcase(acc / 2)


@pytest.mark.gpu
def test_gpu_matches_the_references_showcase_rendering(built):
    spec = scenes.showcase()
    psc = scenes.build_product(spec)
    acc = np.zeros((320, 480, 3))
    frames = 24
    for k in range(frames):
        tr = scenes.product_tracer(spec, psc, 10, 50, cutoff=1e-4, antialias=1.0, seed=100 + k)
        rgb, _, _ = tr.RenderSums(480, 320, psc, sample_count=50)
        acc += np.clip(rgb.astype(np.float64) / 50, 0, 1)
    _check_showcase(acc / frames)


# ---- RayCaster: examples/decoration/rose/rendering.png ---------------------------------------
def _rose_setup():
    import math
    z = np.load(os.path.join(GOLD, "showcase_models.npz"))
    tris = z["rose_v"][z["rose_f"]].astype(np.float32)
    spec = scenes.c1_scene(4)  # SaveRendering's material, light and camera rules (helpers.go:101-128)
    spec["objects"][0]["tris"] = tris
    origin = np.array([0.0, -2.0, 4.0])  # rose/main.go:31
    v = tris.reshape(-1, 3).astype(np.float64)
    center = (v.min(0) + v.max(0)) / 2
    return spec, tris, origin, center, math.pi / 3.6


def _check_rose(img_1000):
    """img_1000: linear 1000x1000x3 frame (SaveRendering renders at 2x and box-filters)."""
    ref8 = read_png_rgb8(os.path.join(GOLD, "ref_rose_rendering.png"))
    lin = np.clip(img_1000.reshape(500, 2, 500, 2, 3).mean(axis=(1, 3)), 0, 1)
    ours8 = (np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055) * 255.999).astype(np.int64)
    bg_ref, bg_ours = ref8.sum(axis=2) == 0, ours8.sum(axis=2) == 0
    # the mesh in showcase/models is a later export of the same model: the silhouette is the same,
    # individual facets differ, so the foreground is compared on average
    assert (bg_ref == bg_ours).mean() > 0.995, (bg_ref == bg_ours).mean()
    fg = ~bg_ref & ~bg_ours
    assert 0.2 < fg.mean() < 0.3
    ratio = srgb_expand(ours8[fg]).mean(axis=0) / srgb_expand(ref8[fg]).mean(axis=0)
    assert np.abs(ratio[:2] - 1).max() < 0.01, ratio  # red / green (blue is ~6/255: quantisation)
    d = np.abs(ours8 - ref8.astype(np.int64)).max(axis=2)
    assert (d <= 3).mean() > 0.88, (d <= 3).mean()


def test_oracle_raycaster_matches_the_references_rose_rendering(oracle):
    spec, tris, origin, center, fov = _rose_setup()
    osc = scenes.build_oracle(spec)
    ocam = oracle.camera_at(tuple(origin), tuple(center), fov)
    ol = oracle.PointLight()
    ol.origin[:] = tuple(center + (origin - center) * 1000)
    ol.color[:] = (1.0, 1.0, 1.0)
    _check_rose(osc.render_raycast(ocam, [ol], 1000, 1000, threads=8)["img"])


@pytest.mark.gpu
def test_gpu_raycaster_matches_the_references_rose_rendering(built):
    from model3d_b200 import render3d as R
    spec, tris, origin, center, fov = _rose_setup()
    psc = scenes.build_product(spec)
    img = R.Image(1000, 1000)
    R.RayCaster(Camera=R.NewCameraAt(tuple(origin), tuple(center), fov),
                Lights=[R.PointLight(tuple(center + (origin - center) * 1000), (1.0, 1.0, 1.0))]).Render(img, psc)
    _check_rose(img.Data.astype(np.float64))


@pytest.mark.gpu
def test_gpu_matches_the_references_showcase_hd_rendering(built):
    """showcase/output_hd.png (960x640; main.go's HighRes mode: adaptive 1000...100000 spp until
    MaxStddev 0.02): a nearly converged frame, so one fixed-spp frame of ours is compared with it
    (measured on B200 at 1024 spp: median block difference 1.3 %, 90th percentile 3.3 %, two glass-rim
    blocks at 26 %, image mean +1.3 % -- the same offset the low-resolution pin shows without the vase)."""
    spec = scenes.showcase(hd=True)
    psc = scenes.build_product(spec)
    n = 1024
    tr = scenes.product_tracer(spec, psc, 10, n, cutoff=1e-4, antialias=1.0, seed=31)
    rgb, _, _ = tr.RenderSums(960, 640, psc, sample_count=n)
    _check_showcase(rgb.astype(np.float64) / n, hd=True)


# ---- flat vs interpolated normals: examples/renderings/smooth_shading/rendering.png -----------------
def _smooth_shading_spec():
    """smooth_shading/main.go:13-55 without the two text labels (they need model2d)."""
    from model3d_b200 import meshes
    ico = meshes.NewMeshIcosphere((0, 0, 0), 1.0, 4)
    assert ico.shape[0] == 320
    mat = scenes.phong(10.0, specular=scenes.gray(0.15), diffuse=scenes.gray(0.75), ambient=scenes.gray(0.1))
    flat = (ico + np.array([-1.3, 0.0, 0.0])).astype(np.float32)
    smooth = (ico + np.array([1.3, 0.0, 0.0])).astype(np.float32)
    return dict(objects=[dict(kind="mesh", tris=flat, material=mat),
                         dict(kind="mesh", tris=smooth, material=mat, vnormals=meshes.VertexNormals(smooth))],
                camera=dict(src=(0.0, -8.0, -0.6), dst=(0.0, 0.0, -0.6), fov=0.8),
                lights=[dict(origin=(2.0, -10.0, 4.0), color=scenes.gray(1.0))])


def _check_smooth_shading(img, tol8, frac):
    """img: linear 1728x3072x3 frame; main.go renders at 4x and calls Image.Downsample(4)."""
    ref8 = read_png_rgb8(os.path.join(GOLD, "ref_smooth_shading_rendering.png"))
    assert ref8.shape == (432, 768, 3)
    lin = np.clip(img.reshape(432, 4, 768, 4, 3).mean(axis=(1, 3)), 0, 1)
    ours8 = (np.where(lin <= 0.0031308, 12.92 * lin, 1.055 * lin ** (1 / 2.4) - 0.055) * 255.999).astype(np.int64)
    rows = slice(0, 268)  # the labels start below the spheres
    r, o = ref8[rows].astype(np.int64), ours8[rows]
    lit_r, lit_o = r.sum(axis=2) > 0, o.sum(axis=2) > 0
    assert 0.3 < lit_r.mean() < 0.5
    assert (lit_r == lit_o).mean() > 0.9995, (lit_r == lit_o).mean()  # same silhouettes
    d = np.abs(r - o).max(axis=2)
    both = lit_r & lit_o
    assert (d[both] <= tol8).mean() > frac, ((d[both] <= tol8).mean(), d[both].max())
    # each sphere separately: flat facets on the left, smooth interpolation on the right
    for cols in (slice(0, 384), slice(384, 768)):
        m = both[:, cols]
        ratio = srgb_expand(o[:, cols][m]).mean() / srgb_expand(r[:, cols][m]).mean()
        assert abs(ratio - 1) < 0.003, ratio


def test_vertex_normals_restatement():
    """meshes.VertexNormals against a plain-loop restatement of mesh_ops.go:146-169."""
    import math
    from model3d_b200 import meshes
    tris = meshes.NewMeshIcosphere((0.3, -0.1, 0.2), 1.5, 3)
    got = meshes.VertexNormals(tris)
    sums = {}
    for t in tris:
        e = [t[0] - t[1], t[1] - t[2], t[2] - t[0]]
        e = [v / math.sqrt(float(v @ v)) for v in e]
        nrm = np.cross(t[1] - t[0], t[2] - t[0])
        nrm = nrm / math.sqrt(float(nrm @ nrm))
        for i in range(3):
            theta = math.acos(max(-1.0, min(1.0, -float(e[(i + 2) % 3] @ e[i]))))
            k = tuple(t[i].tolist())
            sums[k] = sums.get(k, np.zeros(3)) + nrm * theta
    for ti, t in enumerate(tris):
        for i in range(3):
            v = sums[tuple(t[i].tolist())]
            assert np.allclose(got[ti, i], v / math.sqrt(float(v @ v)), atol=1e-12)
    # on a sphere the vertex normals point away from the centre
    radial = (tris - np.array([0.3, -0.1, 0.2])) / 1.5
    assert np.abs(got - radial).max() < 0.02


def test_oracle_matches_the_references_smooth_shading_rendering(oracle):
    spec = _smooth_shading_spec()
    osc = scenes.build_oracle(spec)
    cam = spec["camera"]
    ocam = oracle.camera_at(cam["src"], cam["dst"], cam["fov"])
    ol = oracle.PointLight()
    ol.origin[:] = spec["lights"][0]["origin"]
    ol.color[:] = (1.0, 1.0, 1.0)
    _check_smooth_shading(osc.render_raycast(ocam, [ol], 3072, 1728, threads=8)["img"], 1, 0.995)


@pytest.mark.gpu
def test_gpu_matches_the_references_smooth_shading_rendering(built):
    from model3d_b200 import render3d as R
    spec = _smooth_shading_spec()
    psc = scenes.build_product(spec)
    cam = spec["camera"]
    img = R.Image(3072, 1728)
    rc = R.RayCaster(Camera=R.NewCameraAt(cam["src"], cam["dst"], cam["fov"]),
                     Lights=[R.PointLight(Origin=spec["lights"][0]["origin"], Color=(1.0, 1.0, 1.0))])
    rc.Render(img, psc)
    _check_smooth_shading(np.asarray(img.Data, np.float64).reshape(1728, 3072, 3), 1, 0.99)
