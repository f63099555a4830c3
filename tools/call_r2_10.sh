#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2_bench10_n2.json 2> gpurun_out/r2_bench10_n2.err
tail -n 3 gpurun_out/r2_bench10_n2.err | cut -c1-300; cat gpurun_out/r2_bench10_n2.json | cut -c1-500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2_bench10_c5_n2.json 2> gpurun_out/r2_bench10_c5_n2.err
tail -n 3 gpurun_out/r2_bench10_c5_n2.err | cut -c1-300; cat gpurun_out/r2_bench10_c5_n2.json | cut -c1-500
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench10_ref_n2.json 2> gpurun_out/r2_bench10_ref_n2.err
cat gpurun_out/r2_bench10_ref_n2.json | cut -c1-300
