#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -q -s > gpurun_out/r2_tests6.log 2>&1
grep -E "passed|failed|FAILED|one-kernel|decisive|config|rows " gpurun_out/r2_tests6.log | tail -n 40
STEPS=6 timeout 600 python tools/mega_debug.py > gpurun_out/r2_mega_debug6.txt 2>&1
grep -E "rows (32|64)" gpurun_out/r2_mega_debug6.txt
timeout 300 python tools/mega_ab.py 64 > gpurun_out/r2_mega_ab6.txt 2>&1
cat gpurun_out/r2_mega_ab6.txt
GITB200_TIMELINE=1 timeout 300 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild.log 2>&1
timeout 300 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline6.txt 2>&1
grep -E "L2 |lm_head|step total" gpurun_out/r2_mega_timeline6.txt
