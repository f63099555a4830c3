#!/bin/bash
mkdir -p gpurun_out
ROWS=32 timeout 300 python tools/mega_debug2.py > gpurun_out/r2_mega_debug2.txt 2>&1
cat gpurun_out/r2_mega_debug2.txt
ROWS=64 timeout 300 python tools/mega_debug2.py > gpurun_out/r2_mega_debug2_64.txt 2>&1
cat gpurun_out/r2_mega_debug2_64.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "image_features or teacher_forced" > gpurun_out/r2_tests5.log 2>&1
tail -n 12 gpurun_out/r2_tests5.log
timeout 300 python tools/mega_ab.py 64 > gpurun_out/r2_mega_ab5.txt 2>&1
cat gpurun_out/r2_mega_ab5.txt
timeout 300 python bench.py --steps 8 --no-cpu-baseline --no-micro > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
cat gpurun_out/r2_bench5.json | cut -c1-400
GITB200_TIMELINE=1 timeout 300 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild.log 2>&1
timeout 300 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline5.txt 2>&1
grep -E "L2 |lm_head|step total" gpurun_out/r2_mega_timeline5.txt
