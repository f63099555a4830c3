#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "one_kernel or reproducible or config2 or config4 or eos_forcing or long_max or prefix or coalesced or semantics_replay or teacher_forced" > gpurun_out/r2_tests7.log 2>&1
grep -E "passed|failed|FAILED|one-kernel|row |rows |Error" gpurun_out/r2_tests7.log | tail -n 40
timeout 300 python tools/mega_ab.py 64 16 > gpurun_out/r2_mega_ab7.txt 2>&1
cat gpurun_out/r2_mega_ab7.txt
NSEED=60 timeout 900 python tools/decisive_pick.py > gpurun_out/r2_decisive_pick.txt 2>&1
cat gpurun_out/r2_decisive_pick.txt
GITB200_TIMELINE=1 timeout 300 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild.log 2>&1
timeout 300 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline7.txt 2>&1
grep -E "L2 |lm_head|step total" gpurun_out/r2_mega_timeline7.txt
