#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/launches_r02_config3.csv python bench.py --config 3 --steps 1 --warmup 3 --no-cpu-baseline --no-micro --no-serving --ncu-range > gpurun_out/r2_ncu12.log 2>&1
tail -n 2 gpurun_out/r2_ncu12.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file gpurun_out/launches_r02_config4.csv python bench.py --config 4 --steps 1 --warmup 3 --no-cpu-baseline --no-micro --no-serving --ncu-range > gpurun_out/r2_ncu12b.log 2>&1
tail -n 2 gpurun_out/r2_ncu12b.log | cut -c1-200
