#!/bin/bash
# Round 2, call 19: 256-bit A loads + release-only grid barrier in decode_mega_kernel -- parity subset, bench, timeline, barrier bench.
mkdir -p gpurun_out
timeout 60 tools/_bin/barrier_bench > gpurun_out/r2_barrier_bench19.txt 2>&1
cat gpurun_out/r2_barrier_bench19.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "one_kernel or reproducible or config2 or config4 or decisive or teacher or eos_forcing or prefix" > gpurun_out/r2_tests19.log 2>&1
tail -n 3 gpurun_out/r2_tests19.log
timeout 200 python bench.py --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench19_c2.json 2> gpurun_out/r2_bench19_c2.err
tail -n 2 gpurun_out/r2_bench19_c2.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench19_c2.json
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench19_c4.json 2> gpurun_out/r2_bench19_c4.err
cut -c1-330 gpurun_out/r2_bench19_c4.json
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild19.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline19.txt 2>&1
tail -n 58 gpurun_out/r2_mega_timeline19.txt
