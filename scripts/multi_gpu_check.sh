#!/bin/bash
# Multi-GPU validation on an N-GPU box: the two-device tests, then C3/C4/C5 with the fused flush,
# the NCCL reduce (A/B) and the single-process multi-device context.  Usage: multi_gpu_check.sh N [quick]
N=${1:-2}
Q=${2:-}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/r2_tmulti_n$N.log 2>&1
tail -4 $OUT/r2_tmulti_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for wl in c3 c4 c5; do
  steps=5; [ "$wl" = c5 ] && steps=3
  for mode in fused nccl; do
    $TR bench.py --gpus $N --workload $wl --steps $steps --warmup 3 --reduce $mode > $OUT/r2_${wl}_n${N}_$mode.json 2> $OUT/r2_${wl}_n${N}_$mode.err
    tail -2 $OUT/r2_${wl}_n${N}_$mode.err | cut -c1-300
    cut -c1-200 $OUT/r2_${wl}_n${N}_$mode.json
  done
  python bench.py --gpus $N --workload $wl --steps $steps --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2_${wl}_n${N}_lib.json 2> $OUT/r2_${wl}_n${N}_lib.err
  tail -2 $OUT/r2_${wl}_n${N}_lib.err | cut -c1-300
  cut -c1-200 $OUT/r2_${wl}_n${N}_lib.json
  [ -n "$Q" ] && break
done
