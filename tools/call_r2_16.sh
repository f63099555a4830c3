#!/bin/bash
# Round 2, call 16: HEAD evidence -- bench lines (config 2 with the CPU baseline, config 4), the ncu launch
# list of one config-2 call and one full ncu capture each of decode_mega_kernel, flash_attn_tc_kernel and the c_fc GEMM.
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2_bench16_c2.json 2> gpurun_out/r2_bench16_c2.err
tail -n 2 gpurun_out/r2_bench16_c2.err | cut -c1-200; cut -c1-400 gpurun_out/r2_bench16_c2.json
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro > gpurun_out/r2_bench16_c4.json 2> gpurun_out/r2_bench16_c4.err
cut -c1-300 gpurun_out/r2_bench16_c4.json
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches16.csv python tools/one_call.py 64 > gpurun_out/r2_ncu16_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches16.csv | head -n 12
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_mega -s 5 -c 1 -f -o gpurun_out/r2_mega16 python tools/one_call.py 64 > gpurun_out/r2_ncu16_mega.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash_attn_tc -s 3 -c 1 -f -o gpurun_out/r2_fa16 python tools/one_call.py 64 > gpurun_out/r2_ncu16_fa.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm2_bf16 -s 8 -c 3 -f -o gpurun_out/r2_gemm16 python tools/one_call.py 64 > gpurun_out/r2_ncu16_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
