#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "one_kernel or reproducible or config2" > gpurun_out/r2_tests8.log 2>&1
grep -E "passed|failed|FAILED|one-kernel|rows |Error|error" gpurun_out/r2_tests8.log | tail -n 20
timeout 120 python tools/mega_ab.py 64 > gpurun_out/r2_mega_ab8.txt 2>&1
cat gpurun_out/r2_mega_ab8.txt | tail -5
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline8.txt 2>&1
grep -E "L2 |lm_head|step total" gpurun_out/r2_mega_timeline8.txt; tail -3 gpurun_out/r2_mega_timeline8.txt
