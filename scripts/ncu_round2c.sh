#!/bin/bash
# ncu evidence at the end of round 2 (one GPU): launch lists of the bench commands and --set full captures
# of the dominant kernels.  Outputs under gpurun_out/; summaries are made with scripts/profile_summaries.py
# and scripts/ncu_brief.py.
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_r2c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $O/ncu_r2c_l0.log 2>&1
for wl in c1 c3 c4 c5; do
  spp=4; [ $wl = c5 ] && spp=1
  $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/launches_${wl}_r2c.csv python bench.py --workload $wl --spp $spp --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_r2c_l$wl.log 2>&1
done
F="--set full --import-source on -f"
B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$NCU $F -k regex:trace_first_hit_kernel -s 4 -c 1 -o $O/prof_trace_r2c python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e > $O/ncu_r2c_1.log 2>&1
$NCU $F -k regex:finish_mesh_hits -s 4 -c 1 -o $O/prof_finish_r2c python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e > $O/ncu_r2c_2.log 2>&1
$NCU $F -k regex:path_resolve_kernel -s 7 -c 1 -o $O/prof_resolve_c3_r2c python bench.py --workload c3 --spp 16 $B > $O/ncu_r2c_3.log 2>&1
$NCU $F -k regex:path_sample_kernel -s 14 -c 3 -o $O/prof_sample_c3_r2c python bench.py --workload c3 --spp 16 $B > $O/ncu_r2c_4.log 2>&1
$NCU $F -k regex:trace_first_hit_kernel -s 7 -c 1 -o $O/prof_trace_c3_r2c python bench.py --workload c3 --spp 16 $B > $O/ncu_r2c_5.log 2>&1
$NCU $F -k regex:path_resolve_kernel -s 12 -c 1 -o $O/prof_resolve_c4_r2c python bench.py --workload c4 --spp 16 $B > $O/ncu_r2c_6.log 2>&1
$NCU $F -k regex:trace_first_hit_kernel -s 12 -c 1 -o $O/prof_trace_c4_r2c python bench.py --workload c4 --spp 16 $B > $O/ncu_r2c_7.log 2>&1
$NCU $F -k regex:bidir_shade_kernel -s 6 -c 1 -o $O/prof_bshade_r2c python bench.py --workload c5 --spp 8 $B > $O/ncu_r2c_8.log 2>&1
$NCU $F -k regex:bidir_connect_kernel -s 12 -c 1 -o $O/prof_connect_r2c python bench.py --workload c5 --spp 8 $B > $O/ncu_r2c_9.log 2>&1
$NCU $F -k regex:bidir_prefix_kernel -s 1 -c 1 -o $O/prof_prefix_r2c python bench.py --workload c5 --spp 8 $B > $O/ncu_r2c_10.log 2>&1
ls -la $O/*_r2c.ncu-rep $O/launches*_r2c.csv
