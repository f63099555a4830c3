#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "beam or config3 or prefix or coalesced" > gpurun_out/r2_tests13.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_tests13.log | tail -n 10
timeout 300 python bench.py --config 3 --no-cpu-baseline --no-micro > gpurun_out/r2_bench13_c3.json 2> gpurun_out/r2_bench13_c3.err
tail -n 2 gpurun_out/r2_bench13_c3.err | cut -c1-200; cat gpurun_out/r2_bench13_c3.json | cut -c1-330
