"""Host<->device copy bandwidth of this box with pinned memory (the bound of the host-buffer
first-hit call): H2D alone, D2H alone, both at once on two streams."""
import torch
n_in, n_out = (1 << 24) * 24, (1 << 24) * 20
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device="cuda")
dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    s1.synchronize()
    s2.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


import time
for name, a, b in (("H2D alone", True, False), ("D2H alone", False, True), ("both", True, True)):
    run(a, b, 2)
    t0 = time.perf_counter()
    run(a, b, 10)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print("%s: %.3f ms per step; H2D %.1f GB/s D2H %.1f GB/s" % (name, ms, n_in / ms / 1e6 if a else 0, n_out / ms / 1e6 if b else 0))
