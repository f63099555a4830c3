#!/bin/bash
# Round 2, call 27: ncu launch list of one config-3 call (GIT_LARGE, batch 32, beam 4: the kernel-chain decode path).
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches27_c3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --no-micro --no-serving --ncu-range > gpurun_out/r2_ncu27.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches27_c3.csv | head -n 30
