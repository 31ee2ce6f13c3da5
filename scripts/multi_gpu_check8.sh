#!/bin/bash
# 8-GPU validation: two-device tests on all devices, the headline bench under torchrun (its path-tracing
# block runs C3/C4/C5 at BASELINE spp with the fused flush), then NCCL-reduce and library-mode A/B lines.
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py tests/test_c_abi.py -m gpu -q > $OUT/r2_tmulti_n$N.log 2>&1
tail -3 $OUT/r2_tmulti_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err
tail -2 $OUT/r2_bench_n$N.err | cut -c1-300
cut -c1-300 $OUT/r2_bench_n$N.json
for wl in c3 c4 c5; do
  steps=5; [ "$wl" = c5 ] && steps=3
  $TR bench.py --gpus $N --workload $wl --steps $steps --warmup 3 --reduce nccl > $OUT/r2_${wl}_n${N}_nccl.json 2> $OUT/r2_${wl}_n${N}_nccl.err
  cut -c1-160 $OUT/r2_${wl}_n${N}_nccl.json
  python bench.py --gpus $N --workload $wl --steps $steps --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2_${wl}_n${N}_lib.json 2> $OUT/r2_${wl}_n${N}_lib.err
  tail -1 $OUT/r2_${wl}_n${N}_lib.err | cut -c1-300
  cut -c1-160 $OUT/r2_${wl}_n${N}_lib.json
done
nvidia-smi topo -m > $OUT/r2_topo_n$N.txt 2>&1
