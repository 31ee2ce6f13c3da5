#!/bin/bash
# Final multi-GPU check of a round: the multi-device tests, then the driver's own bench command at N GPUs
# (torchrun, default steps) and the single-process multi-device context on C3.  Usage: multi_gpu_final.sh N
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_multi.py tests/test_c_abi.py -m gpu -q > $OUT/r2c_tmulti_n$N.log 2>&1
tail -3 $OUT/r2c_tmulti_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_r2c_n$N.json 2> $OUT/bench_r2c_n$N.err
tail -2 $OUT/bench_r2c_n$N.err | cut -c1-300
python - $OUT/bench_r2c_n$N.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("N", d["n_gpus"], "C2", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "frac_of_copy_bound", d["e2e"].get("frac_of_copy_bound"))
for k, v in d.get("path_tracing", {}).items():
    print(k, round(v["Msamples_per_s"], 1), "Msamples/s", "render", v.get("render_ms_per_rank"), "wait", v.get("reduce_wait_ms_per_rank"))
PY
python bench.py --gpus $N --workload c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_r2c_c3_n${N}_lib.json 2> $OUT/bench_r2c_c3_n${N}_lib.err
cut -c1-160 $OUT/bench_r2c_c3_n${N}_lib.json
