#!/bin/bash
# Round 2, call 23: L2 prefetch of the image K/V a layer ahead (A/B), LM-head statistics pipelined -- parity subset, bench, timeline.
mkdir -p gpurun_out
timeout 200 python tools/mega_opt_ab.py mega_l2_prefetch 64 > gpurun_out/r2_l2_prefetch_ab23.txt 2>&1
cat gpurun_out/r2_l2_prefetch_ab23.txt | tail -n 5
VIDEO=1 timeout 200 python tools/mega_opt_ab.py mega_l2_prefetch 16 > gpurun_out/r2_l2_prefetch_ab23_video.txt 2>&1
cat gpurun_out/r2_l2_prefetch_ab23_video.txt | tail -n 5
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config2 or config4 or long_max_steps or decisive or one_kernel or reproducible or teacher" > gpurun_out/r2_tests23.log 2>&1
tail -n 2 gpurun_out/r2_tests23.log
timeout 200 python bench.py --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench23_c2.json 2> gpurun_out/r2_bench23_c2.err
tail -n 2 gpurun_out/r2_bench23_c2.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench23_c2.json
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild23.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline23.txt 2>&1
grep -E "L3 |lm_head|step total" gpurun_out/r2_mega_timeline23.txt | cut -c1-90; grep -A8 "qkv    barrier released" gpurun_out/r2_mega_timeline23.txt | cut -c1-90; grep -A10 "LM head of that step" gpurun_out/r2_mega_timeline23.txt | cut -c1-90
