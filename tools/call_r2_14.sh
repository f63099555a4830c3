#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "one_kernel or reproducible or golden or config or prefix or decisive or teacher or greedy" > gpurun_out/r2_tests14.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_tests14.log | tail -n 10
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro > gpurun_out/r2_bench14_c4.json 2> gpurun_out/r2_bench14_c4.err
tail -n 2 gpurun_out/r2_bench14_c4.err | cut -c1-200; cat gpurun_out/r2_bench14_c4.json | cut -c1-330
timeout 200 python bench.py --config 2 --no-cpu-baseline --no-micro > gpurun_out/r2_bench14_c2.json 2> gpurun_out/r2_bench14_c2.err
tail -n 2 gpurun_out/r2_bench14_c2.err | cut -c1-200; cat gpurun_out/r2_bench14_c2.json | cut -c1-330
