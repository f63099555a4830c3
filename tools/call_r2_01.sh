#!/bin/bash
# Round 2, call 1: settle what round 1 left unmeasured + CPU thread scaling of the reference arm.
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt; lscpu | head -20 >> gpurun_out/r2_nproc.txt
GITB200_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_parity.py -q -k experimental > gpurun_out/r2_experimental_tests.log 2>&1
tail -5 gpurun_out/r2_experimental_tests.log
timeout 120 python tools/attn_ab.py kv_head_major > gpurun_out/r2_kv_head_major_ab.txt 2>&1
cat gpurun_out/r2_kv_head_major_ab.txt
timeout 200 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:decode_attn -s 6 -c 1 \
    -f -o gpurun_out/r2_decode_attn_256 python tools/one_call.py 256 > gpurun_out/r2_ncu_decode_attn.log 2>&1
tail -2 gpurun_out/r2_ncu_decode_attn.log
timeout 200 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:flash_attn -s 2 -c 1 \
    -f -o gpurun_out/r2_flash_attn_64 python tools/one_call.py 64 > gpurun_out/r2_ncu_flash.log 2>&1
tail -2 gpurun_out/r2_ncu_flash.log
timeout 400 python tools/cpu_threads_probe.py > gpurun_out/r2_cpu_threads_probe.txt 2>&1
cat gpurun_out/r2_cpu_threads_probe.txt
