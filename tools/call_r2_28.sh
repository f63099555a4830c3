#!/bin/bash
# Round 2, call 28 (last of the round): flag chain for the beam steps + flash_attn_tc_long_kernel -- full GPU suite, bench config 3.
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests28.log 2>&1
tail -n 3 gpurun_out/r2_tests28.log
timeout 100 python bench.py --config 3 --no-cpu-baseline --no-micro > gpurun_out/r2_bench28_c3.json 2> gpurun_out/r2_bench28_c3.err
tail -n 2 gpurun_out/r2_bench28_c3.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench28_c3.json
