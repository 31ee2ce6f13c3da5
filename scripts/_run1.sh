M3D_LIB=$PWD/variants/libm3dgpu_chk.so python -m pytest tests/test_gpu_bidir.py -x -q -s -k "general_power or balance" 2>&1 | grep -E "MIS|passed|failed|Error|assert" | sort | uniq -c | sort -rn | head -20
python -m pytest tests/test_gpu_bidir.py tests/test_gpu_multi.py tests/test_c_abi.py -x -q 2>&1 | tail -3
python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2b_c5_tables.json
