#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "flash_attention" > gpurun_out/r2_fa_tests.log 2>&1
tail -n 15 gpurun_out/r2_fa_tests.log
timeout 600 python tools/mega_debug.py > gpurun_out/r2_mega_debug.txt 2>&1
cat gpurun_out/r2_mega_debug.txt
GITB200_TIMELINE=1 timeout 300 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild.log 2>&1
timeout 300 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline.txt 2>&1
tail -n 100 gpurun_out/r2_mega_timeline.txt
