"""render3d's convenience renderers on the GPU RayCaster: the most-used consumers
of the hot path in the reference (render3d/helpers.go:71-266, cli/render_stl).

  Objectify          helpers.go:56-99   (default yellow Phong material)
  SaveRendering      helpers.go:101-128
  SaveRandomGrid     helpers.go:130-167 (rows x cols views of ONE device scene: the BVH is
                                         built once and every view is a batch of camera rays)
  SaveRotatingGIF    helpers.go:169-236 (the object is not re-uploaded per frame: rotating the
                                         object about its centre equals counter-rotating camera
                                         and light, which leaves RayCaster shading unchanged)
  DirectionalCamera  helpers.go:238-266

ColorFunc closures cannot cross the C ABI: colorFunc must be None (UnsupportedError otherwise).
"""
import math
import struct
import zlib

import numpy as np

from . import render3d as R
from .model3d import MeshCollider, UnsupportedError

helperFieldOfView = math.pi / 3.6
helperAmbient, helperDiffuse, helperSpecular, helperAntialias = 0.1, 0.8, 0.2, 2


def Objectify(obj, colorFunc=None):
    """render3d.Objectify: triangle array / MeshCollider / analytic collider -> Object with the
    default yellow Phong material; an Object is returned unchanged."""
    if colorFunc is not None:
        raise UnsupportedError("ColorFunc closures are not supported on the GPU path")
    if isinstance(obj, (R.ColliderObject, R.JoinedObject, R.Scene, R._Transformed)):
        return obj
    if isinstance(obj, (MeshCollider, np.ndarray, R.Sphere, R.Rect, R.Cylinder)):
        yellow = R.NewColorRGB(224.0 / 255, 209.0 / 255, 0)
        mat = R.PhongMaterial(Alpha=10, SpecularColor=R.NewColor(helperSpecular),
                              DiffuseColor=tuple(helperDiffuse * c for c in yellow),
                              AmbientColor=tuple(helperAmbient * c for c in yellow))
        return R.ColliderObject(Collider=obj, Material=mat)
    raise TypeError("type not recognized")


def _scene_of(obj, colorFunc):
    return R._as_scene(R.JoinedObject([Objectify(obj, colorFunc)]) if not isinstance(obj, R.Scene) else obj)


def _bounds(sc):
    return np.asarray(sc.Min(), np.float64), np.asarray(sc.Max(), np.float64)


def SaveRendering(path, obj, origin, width, height, colorFunc=None):
    """render3d.SaveRendering: camera at `origin` facing the bounding-box centre, one far point
    light behind the camera, 2x supersampling; returns the downsampled Image (also saved when
    path is not None)."""
    sc = _scene_of(obj, colorFunc)
    mn, mx = _bounds(sc)
    center = (mn + mx) / 2
    origin = np.asarray(origin, np.float64)
    caster = R.RayCaster(Camera=R.NewCameraAt(tuple(origin), tuple(center), helperFieldOfView),
                         Lights=[R.PointLight(Origin=tuple(center + (origin - center) * 1000), Color=R.NewColor(1.0))])
    img = R.Image(width * helperAntialias, height * helperAntialias)
    caster.Render(img, sc)
    out = img.Downsample(helperAntialias)
    if path is not None:
        out.Save(path)
    return out


def DirectionalCamera(obj, direction, fov):
    """render3d.DirectionalCamera: bisection on the distance at which the bounding box fits."""
    mn, mx = (np.asarray(obj.Min(), np.float64), np.asarray(obj.Max(), np.float64))
    baseline = float(np.linalg.norm(mx - mn))
    center = (mn + mx) / 2
    direction = np.asarray(direction, np.float64)
    margin = 0.05
    lo, hi = baseline * 1e-4, baseline * 1e4
    corners = [(x, y, z) for x in (mn[0], mx[0]) for y in (mn[1], mx[1]) for z in (mn[2], mx[2])]
    for _ in range(32):
        d = (lo + hi) / 2
        cam = R.NewCameraAt(tuple(center + direction * d), tuple(center), helperFieldOfView)
        unc = R.Uncaster(cam, 1, 1)
        contained = True
        for c in corners:
            sx, sy = unc(c)
            if sx < margin or sy < margin or sx >= 1 - margin or sy >= 1 - margin:
                contained = False
        if contained:
            hi = d
        else:
            lo = d
    return R.NewCameraAt(tuple(center + direction * hi), tuple(center), fov)


def SaveRandomGrid(path, obj, rows, cols, imgSize, colorFunc=None, seed=None):
    """render3d.SaveRandomGrid: rows x cols renderings from random unit directions
    (model3d.NewCoord3DRandUnit -> numpy Generator; the reference uses the global math/rand)."""
    sc = _scene_of(obj, colorFunc)
    mn, mx = _bounds(sc)
    center = (mn + mx) / 2
    rng = np.random.default_rng(seed)
    full = R.Image(cols * imgSize, rows * imgSize)
    casters = []
    for _ in range(rows * cols):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        casters.append(R.RayCaster(Camera=DirectionalCamera(sc, d, helperFieldOfView),
                                   Lights=[R.PointLight(Origin=tuple(center + d * 1000), Color=R.NewColor(1.0))]))
    # all rows x cols views in one call: one BVH, the views' launch chains back to back, the 2x
    # supersampled frames box-filtered on the device, one copy back (spread over the GPUs of a
    # multi-device context)
    views, _ = R.RayCaster.RenderViews(casters, imgSize * helperAntialias, imgSize * helperAntialias, sc,
                                       downsample=helperAntialias)
    for i in range(rows):
        for j in range(cols):
            full.Data[i * imgSize:(i + 1) * imgSize, j * imgSize:(j + 1) * imgSize] = views[i * cols + j]
    if path is not None:
        full.Save(path)
    return full


def _rotation(axis, angle):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    k = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(angle) * k + (1 - math.cos(angle)) * (k @ k)


def SaveRotatingGIF(path, obj, axis, cameraDir, imgSize, frames, fps, colorFunc=None):
    """render3d.SaveRotatingGIF: grayscale animation of the object rotating about `axis`.
    Returns the list of uint8 frames (also written as an animated GIF when path is not None)."""
    sc = _scene_of(obj, colorFunc)
    mn, mx = _bounds(sc)
    center = (mn + mx) / 2
    cam_dir = np.asarray(cameraDir, np.float64)
    cam_dir = cam_dir / np.linalg.norm(cam_dir)
    corners = np.array([(x, y, z) for x in (mn[0], mx[0]) for y in (mn[1], mx[1]) for z in (mn[2], mx[2])])

    class _Box:  # bounds of the rotated object (transform.go:60-74: bounds of the rotated corners)
        def __init__(self, pts):
            self.mn, self.mx = pts.min(axis=0), pts.max(axis=0)

        def Min(self):
            return tuple(self.mn)

        def Max(self):
            return tuple(self.mx)

    furthest, rots = None, []
    for i in range(frames):
        rot = _rotation(axis, math.pi * 2 * i / frames)
        rots.append(rot)
        box = _Box((corners - center) @ rot.T + center)
        cam = DirectionalCamera(box, cam_dir, helperFieldOfView)
        if furthest is None or np.linalg.norm(np.asarray(cam.Origin) - center) > \
                np.linalg.norm(np.asarray(furthest.Origin) - center):
            furthest = cam
    offset = np.asarray(furthest.Origin, np.float64) - center
    light = center + offset * 1000
    casters = []
    for rot in rots:
        inv = rot.T  # counter-rotate camera and light about the centre instead of the object

        def back(p, lin=False):
            p = np.asarray(p, np.float64)
            return tuple(inv @ p) if lin else tuple(inv @ (p - center) + center)

        cam = R.Camera(Origin=back(furthest.Origin), ScreenX=back(furthest.ScreenX, True),
                       ScreenY=back(furthest.ScreenY, True), FieldOfView=furthest.FieldOfView)
        casters.append(R.RayCaster(Camera=cam, Lights=[R.PointLight(Origin=back(light), Color=R.NewColor(1.0))]))
    views, _ = R.RayCaster.RenderViews(casters, imgSize, imgSize, sc)  # all frames in one call
    out = []
    for v in views:
        img = R.Image(imgSize, imgSize)
        img.Data = np.array(v)
        out.append(img.Gray8())
    if path is not None:
        write_gif(path, out, int(math.ceil(100 / fps)))
    return out


# ---- minimal encoders (image/png and image/gif stand-ins; off the hot path) -------------------
def write_png(path, rgb8):
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def _gif_lzw_uncompressed(idx):
    """GIF image data for 8-bit indices without compression: 9-bit codes with a clear code
    often enough that the code size never grows."""
    out = bytearray()
    acc = bits = 0
    codes = []
    for k, v in enumerate(idx):
        if k % 254 == 0:
            codes.append(256)  # clear
        codes.append(int(v))
    codes.append(257)  # end of information
    for c in codes:
        acc |= c << bits
        bits += 9
        while bits >= 8:
            out.append(acc & 0xff)
            acc >>= 8
            bits -= 8
    if bits:
        out.append(acc & 0xff)
    blocks = bytearray()
    for i in range(0, len(out), 255):
        b = out[i:i + 255]
        blocks.append(len(b))
        blocks += b
    blocks.append(0)
    return bytes(blocks)


def write_gif(path, gray_frames, delay_cs):
    h, w = gray_frames[0].shape
    pal = b"".join(bytes((i, i, i)) for i in range(256))
    data = bytearray(b"GIF89a" + struct.pack("<HHBBB", w, h, 0xF7, 0, 0) + pal)
    data += b"\x21\xff\x0bNETSCAPE2.0\x03\x01\x00\x00\x00"  # loop forever
    for fr in gray_frames:
        data += b"\x21\xf9\x04\x00" + struct.pack("<H", delay_cs) + b"\x00\x00"
        data += b"\x2c" + struct.pack("<HHHHB", 0, 0, w, h, 0) + b"\x08"
        data += _gif_lzw_uncompressed(np.ascontiguousarray(fr, np.uint8).ravel())
    data += b"\x3b"
    with open(path, "wb") as f:
        f.write(bytes(data))
