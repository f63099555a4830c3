#!/bin/bash
# Round 2, call 25: evidence for the final code -- full GPU suite, bench lines of configs 2-5 and the reference arm, ncu
# launch list of one config-2 call, ncu --set full of decode_mega_kernel and flash_attn_tc_kernel.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests25.log 2>&1
tail -n 3 gpurun_out/r2_tests25.log
timeout 400 python bench.py > gpurun_out/r2_bench25_c2.json 2> gpurun_out/r2_bench25_c2.err
tail -n 2 gpurun_out/r2_bench25_c2.err | cut -c1-200; cut -c1-300 gpurun_out/r2_bench25_c2.json
for c in 3 4 5; do
  timeout 300 python bench.py --config $c --no-cpu-baseline --no-micro > gpurun_out/r2_bench25_c$c.json 2> gpurun_out/r2_bench25_c$c.err
  cut -c1-260 gpurun_out/r2_bench25_c$c.json
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench25_ref.json 2> gpurun_out/r2_bench25_ref.err
cut -c1-300 gpurun_out/r2_bench25_ref.json
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches25.csv python tools/one_call.py 64 > gpurun_out/r2_ncu25_list.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches25.csv 2>/dev/null | head -n 8
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_mega -s 5 -c 1 -f -o gpurun_out/r2_mega25 python tools/one_call.py 64 > gpurun_out/r2_ncu25_mega.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash_attn_tc -s 3 -c 1 -f -o gpurun_out/r2_fa25 python tools/one_call.py 64 > gpurun_out/r2_ncu25_fa.log 2>&1
ls -la gpurun_out/*25.ncu-rep
