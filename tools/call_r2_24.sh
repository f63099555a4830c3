#!/bin/bash
# Round 2, call 24 (8 GPUs): BASELINE.json config 5 -- GIT_LARGE, 8192 synthetic images, greedy, image-parallel over 8 x B200
# with ONE all_gather of the finished captions -- and config 2 at 8 GPUs.
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2_bench24_c5_n8.json 2> gpurun_out/r2_bench24_c5_n8.err
tail -n 3 gpurun_out/r2_bench24_c5_n8.err | cut -c1-300; cut -c1-600 gpurun_out/r2_bench24_c5_n8.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 16 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2_bench24_c2_n8.json 2> gpurun_out/r2_bench24_c2_n8.err
tail -n 3 gpurun_out/r2_bench24_c2_n8.err | cut -c1-300; cut -c1-600 gpurun_out/r2_bench24_c2_n8.json
