#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/_bin/barrier_bench > gpurun_out/r2_barrier_bench.txt 2>&1
cat gpurun_out/r2_barrier_bench.txt
