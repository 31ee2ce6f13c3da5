#!/usr/bin/env python
"""Registers / stack / spills per kernel of one .cu file of model3d_b200/csrc:
    python scripts/ptxas_summary.py path_kernels.cu [-DX=Y ...]"""
import os
import re
import subprocess
import sys

src = sys.argv[1]
extra = sys.argv[2:]
csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "model3d_b200", "csrc")
cmd = ["nvcc", "-DM3D_HAVE_SCENE", "-DM3D_HAVE_RAYCAST", "-DM3D_HAVE_PATH", "-DM3D_HAVE_BIDIR", "-gencode",
       "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr", "-Xptxas", "-v",
       "-c", os.path.join(csrc, src), "-o", "/tmp/_ptxas_summary.o"] + extra
out = subprocess.run(cmd, capture_output=True, text=True).stderr
filt = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout.splitlines()
for i, l in enumerate(filt):
    m = re.search(r"Compiling entry function '(.*)' for", l)
    if not m:
        continue
    name = re.sub(r"\(anonymous namespace\)::|m3d::|void ", "", m.group(1)).split("(")[0]
    stack = spill = regs = None
    for k in filt[i + 1:i + 8]:
        s = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", k)
        if s and stack is None:
            stack, spill = s.group(1), s.group(2) + "/" + s.group(3)
        r = re.search(r"Used (\d+) registers", k)
        if r and regs is None:
            regs = r.group(1)
            break
    print("%-50s regs %4s  stack %5s  spill st/ld %s" % (name, regs, stack, spill))
