#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "one_kernel or teacher_forced or semantics_replay or eos_forcing or reproducible or config2 or config4 or host_and_tensor or long_max" > gpurun_out/r2_tests3.log 2>&1
tail -n 40 gpurun_out/r2_tests3.log
timeout 300 python tools/mega_ab.py 64 16 > gpurun_out/r2_mega_ab.txt 2>&1
cat gpurun_out/r2_mega_ab.txt
timeout 600 python bench.py --steps 16 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -n 3 gpurun_out/r2_bench3.err; cat gpurun_out/r2_bench3.json
