#!/bin/bash
# ncu evidence of round 2 (one GPU): launch lists of the bench commands and --set full captures of the
# dominant kernels.  Outputs under gpurun_out/; summaries are made from them with scripts/profile_summaries.py.
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $O/ncu_r2_a.log 2>&1
for wl in c1 c3 c4 c5; do
  spp=4; [ $wl = c5 ] && spp=1
  $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/launches_${wl}_r2.csv python bench.py --workload $wl --spp $spp --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_r2_$wl.log 2>&1
done
$NCU --set full --import-source on -k regex:trace_first_hit_kernel -s 4 -c 1 -f -o $O/prof_trace_r2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e > $O/ncu_r2_b.log 2>&1
$NCU --set full --import-source on -k regex:finish_mesh_hits -s 4 -c 1 -f -o $O/prof_finish_r2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e > $O/ncu_r2_c.log 2>&1
$NCU --set full --import-source on -k regex:path_resolve_kernel -s 2 -c 1 -f -o $O/prof_resolve_r2 python bench.py --workload c3 --spp 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_r2_d.log 2>&1
$NCU --set full --import-source on -k regex:bidir_connect_kernel -s 2 -c 1 -f -o $O/prof_connect_r2 python bench.py --workload c5 --spp 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_r2_e.log 2>&1
$NCU --set full --import-source on -k regex:path_flush -s 1 -c 2 -f -o $O/prof_flush_r2 python scripts/flush_mode_timing.py c3 128 > $O/ncu_r2_f.log 2>&1
ls -la $O/*.ncu-rep $O/launches*_r2.csv
