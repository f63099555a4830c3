#!/bin/bash
# Round 2, call 21: 16-row tail boxes for the decode attention units -- parity subset (incl. video, VQA geometry, long
# captions), bench config 2 / 4, timeline with LM-head marks.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_kernel or reproducible or config2 or config4 or decisive or teacher or eos_forcing or prefix or long_max_steps or vqa or semantics" > gpurun_out/r2_tests21.log 2>&1
tail -n 3 gpurun_out/r2_tests21.log
timeout 200 python bench.py --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench21_c2.json 2> gpurun_out/r2_bench21_c2.err
tail -n 2 gpurun_out/r2_bench21_c2.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench21_c2.json
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench21_c4.json 2> gpurun_out/r2_bench21_c4.err
cut -c1-330 gpurun_out/r2_bench21_c4.json
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild21.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline21.txt 2>&1
grep -E "L3 |step total" gpurun_out/r2_mega_timeline21.txt | cut -c1-90; grep -A45 "LM head of that step" gpurun_out/r2_mega_timeline21.txt | cut -c1-90
