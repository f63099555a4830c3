#!/bin/bash
# Round 2, call 22: tail boxes with a single call site of the attention unit body (instruction-cache footprint).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config2 or config4 or long_max_steps or decisive" > gpurun_out/r2_tests22.log 2>&1
tail -n 2 gpurun_out/r2_tests22.log
timeout 200 python bench.py --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench22_c2.json 2> gpurun_out/r2_bench22_c2.err
tail -n 2 gpurun_out/r2_bench22_c2.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench22_c2.json
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench22_c4.json 2> gpurun_out/r2_bench22_c4.err
cut -c1-330 gpurun_out/r2_bench22_c4.json
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild22.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline22.txt 2>&1
grep -E "L3 |step total" gpurun_out/r2_mega_timeline22.txt | cut -c1-90; grep -A8 "qkv    barrier released" gpurun_out/r2_mega_timeline22.txt | cut -c1-90
