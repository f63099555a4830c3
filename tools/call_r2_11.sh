#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "flash_attention" > gpurun_out/r2_fa_tests11.log 2>&1
tail -n 5 gpurun_out/r2_fa_tests11.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "image_features or config2" > gpurun_out/r2_tests11.log 2>&1
tail -n 4 gpurun_out/r2_tests11.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:flash_attn -c 20 --csv --log-file gpurun_out/r2_fa_launches11.csv python tools/one_call.py 64 > gpurun_out/r2_ncu11.log 2>&1
grep -c flash_attn gpurun_out/r2_fa_launches11.csv; grep "gpu__time_duration" gpurun_out/r2_fa_launches11.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | head -6
timeout 200 python bench.py --steps 8 --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
cat gpurun_out/r2_bench11.json | cut -c1-300
