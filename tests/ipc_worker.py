"""Helper process of test_gpu_multi.py::test_ipc_accumulator_shared_between_processes: maps the
parent's accumulator (m3d_ipc_open) and flushes a sample shard into it with M3D_PART_ATOMIC."""
import sys

import scenes
from model3d_b200 import _native as N


def main():
    handle = bytes.fromhex(sys.argv[1])
    W, H, s_begin, s_count = (int(x) for x in sys.argv[2:6])
    ctx = N.default_context(0)
    acc = N.ipc_open(ctx, handle)
    spec = scenes.cornell_box()
    psc = scenes.build_product(spec)
    tr = scenes.product_tracer(spec, psc, 5, 32, cutoff=1e-4, antialias=1.0, seed=11)
    tr.RenderSumsDevice(W, H, psc, acc, partition=(0, 0, s_begin, N.PART_ATOMIC), sample_count=s_count)
    ctx.synchronize()
    N.ipc_close(ctx, acc)


if __name__ == "__main__":
    main()
