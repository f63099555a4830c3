#!/bin/bash
# Round 2, call 26: tcgen05 attention for S > 512 (flash_attn_tc_long_kernel) -- unit tests, the parity tests that reach it
# (video prefill, VQA geometry), launch times inside a config-4 call, bench config 4.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "flash_attention" > gpurun_out/r2_ktests26.log 2>&1
tail -n 6 gpurun_out/r2_ktests26.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config4 or vatex or vqa or image_features" > gpurun_out/r2_tests26.log 2>&1
tail -n 3 gpurun_out/r2_tests26.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:flash_attn --csv --log-file gpurun_out/r2_fa_launches26.csv python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-micro --no-serving --ncu-range > gpurun_out/r2_ncu26.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_fa_launches26.csv
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro > gpurun_out/r2_bench26_c4.json 2> gpurun_out/r2_bench26_c4.err
tail -n 2 gpurun_out/r2_bench26_c4.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench26_c4.json
