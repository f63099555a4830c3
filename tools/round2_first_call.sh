#!/bin/bash
# One gpurun call that settles what round 1 left unmeasured:
#   gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
# 1. the gated tests of switches written blind at the end of round 1
GITB200_TEST_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_gpu_parity.py -q -k experimental > gpurun_out/experimental_tests.log 2>&1
tail -3 gpurun_out/experimental_tests.log
# 2. does a contiguous (head-major) image K/V slice lift the decode attention off 0.4 of HBM peak?
timeout 90 python tools/attn_ab.py kv_head_major > gpurun_out/kv_head_major_ab.txt 2>&1
cat gpurun_out/kv_head_major_ab.txt
# 3. where the kernel stalls (read here with: ncu -i gpurun_out/decode_attn_256.ncu-rep --page raw --csv)
timeout 150 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:decode_attn -s 6 -c 1 \
    -f -o gpurun_out/decode_attn_256 python tools/one_call.py 256 > gpurun_out/ncu_decode_attn.log 2>&1
tail -2 gpurun_out/ncu_decode_attn.log
