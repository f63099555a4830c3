#!/bin/bash
# Round-2 starting point: where does decode_attn_kernel spend its cycles at 256 rows?  (profiles/decode_256rows_r01.md)
#   gpurun --timeout 300 -- 'bash tools/prof_decode_attn.sh'
# then here:  ncu -i gpurun_out/decode_attn_256.ncu-rep --page raw --csv | grep -E "smsp__average_warp|stall|dram__|lts__|sm__inst"
#             ncu -i gpurun_out/decode_attn_256.ncu-rep --page source --csv | sort by stall samples
set -e
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:decode_attn -s 6 -c 1 \
    -o gpurun_out/decode_attn_256 python tools/one_call.py 256
# A/B of the layouts / variants on whole calls (engine options): attn_pipe 0/1 is in tools/attn_ab.py; head-major copy:
GITB200_KV_HEAD_MAJOR=1 python tools/one_call.py 256
GITB200_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -q -k experimental
