"""L2 -> SM bandwidth of this B200 (m3d_measure_l2_bandwidth): the denominator of roofline.frac_l2.
  python scripts/l2_peak.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from model3d_b200 import _native as N


def measure(ctx, mode, mb, repeats=5):
    g = C.c_double()
    N.check(N.lib().m3d_measure_l2_bandwidth(ctx.h, C.c_int32(mode), C.c_int64(mb << 20), C.c_int32(repeats), C.byref(g)))
    return g.value


if __name__ == "__main__":
    ctx = N.Context(0)
    for mb in (24, 56, 96, 160, 512):
        print("working set %4d MB: stream %.0f GB/s, 80-byte node gather %.0f GB/s" % (mb, measure(ctx, 0, mb), measure(ctx, 1, mb)), flush=True)
    names = {2: "96 B stride, 2xLDG.256+LDG.128 (80 useful)", 3: "64 B records, 2xLDG.256",
             4: "128 B records, 4xLDG.256", 5: "96 B records, 3xLDG.256"}
    for mode, nm in names.items():
        print("56 MB gather, %s: %.0f GB/s" % (nm, measure(ctx, mode, 56)), flush=True)
