"""render_stl: render a 3D model in an STL file to a PNG (a grid of randomized views) or a GIF (a
rotating view) on the GPU -- the counterpart of the reference's cli/render_stl/main.go:17-78, with
the same flags and defaults.

    python -m model3d_b200.cli.render_stl [flags] <model.stl[.gz]> <output.png | output.gif>

Pipeline: fileformats.ReadSTL (binary / ASCII, gzip detected) -> MeshCollider (device-resident wide
BVH; --device-build builds it on the GPU) -> helpers.SaveRandomGrid / SaveRotatingGIF (RayCaster).
There is no CPU fallback: without a CUDA device the command fails with the library's error.
"""
import argparse
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog="render_stl", description=__doc__.split("\n\n")[0])
    ap.add_argument("--grid-size", type=int, default=3, help="grid size (used for rows and columns)")
    ap.add_argument("--image-size", type=int, default=300, help="size of each image in the grid")
    ap.add_argument("--fps", type=float, default=10.0, help="FPS for GIF outputs")
    ap.add_argument("--frames", type=int, default=20, help="total number of frames for GIF outputs")
    ap.add_argument("--verbose", action="store_true", help="run in verbose mode")
    ap.add_argument("--seed", type=int, default=None, help="seed of the random view directions (reference: global math/rand)")
    ap.add_argument("--device-build", action="store_true", help="build the BVH on the GPU (LBVH + wide collapse)")
    ap.add_argument("model", help="<model.stl> (binary or ASCII, optionally gzip-compressed)")
    ap.add_argument("output", help="[output.png | output.gif]")
    args = ap.parse_args(argv)

    from .. import fileformats, helpers
    from ..model3d import MeshCollider

    def log(*a):
        if args.verbose:
            print(time.strftime("%Y/%m/%d %H:%M:%S"), *a, file=sys.stderr, flush=True)

    log("Loading model from", args.model, "...")
    tris = fileformats.ReadSTL(args.model)
    log("Converting mesh to collider ... (%d triangles)" % tris.shape[0])
    collider = MeshCollider(np.asarray(tris, np.float32), device_build=args.device_build)
    log("Rendering mesh to", args.output, "...")
    if args.output.endswith(".gif"):
        d = np.array([0.0, -1.0, 0.1])
        helpers.SaveRotatingGIF(args.output, collider, (0.0, 0.0, 1.0), tuple((d / np.linalg.norm(d)).tolist()),
                                args.image_size, args.frames, args.fps, None)
    else:
        helpers.SaveRandomGrid(args.output, collider, args.grid_size, args.grid_size, args.image_size, None,
                               seed=args.seed)
    return 0


if __name__ == "__main__":
    sys.exit(main())
