#!/bin/bash
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:path_resolve_kernel -s 7 -c 1 -f -o $O/prof_resolve_r2c python bench.py --workload c3 --spp 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_r2c_a.log 2>&1
$NCU --set full --import-source on -k regex:path_sample_kernel -s 14 -c 3 -f -o $O/prof_sample_r2c python bench.py --workload c3 --spp 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_r2c_b.log 2>&1
ls -la $O/*.ncu-rep
