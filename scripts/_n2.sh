#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for extra in "" ""; do
M3D_DEBUG_T=1 $TR bench.py --gpus 2 --steps 20 --warmup 5 $extra 2>gpurun_out/n2_err.log | grep '^{' | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('extra [$extra] C2', round(d['value']), 'e2e', d.get('e2e', {}).get('value'))
for k, v in d.get('path_tracing', {}).items():
    print(' ', k, round(v['Msamples_per_s'], 1), 'render', v.get('render_ms_per_rank'), 'kernel', v.get('kernel_ms_per_rank'), 'wait', v.get('reduce_wait_ms_per_rank'), v.get('clocks'))
"
grep "render_path" gpurun_out/n2_err.log | head -24 | cut -c1-120
done
