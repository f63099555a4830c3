#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_tests9.log 2>&1
grep -E "passed|failed|FAILED|decisive|prefix|parity\]" gpurun_out/r2_tests9.log | tail -n 40
for c in 2 3 4 5; do
  timeout 400 python bench.py --config $c > gpurun_out/r2_bench9_c$c.json 2> gpurun_out/r2_bench9_c$c.err
  tail -n 2 gpurun_out/r2_bench9_c$c.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2_bench9_c$c.json'))
    print('config $c: value %.1f  e2e %.1f  median ms %.2f  serving %s  whole_call %s  cpu %s' % (d['value'], d['e2e']['value'], d['median_ms_per_step'], (d.get('serving') or {}).get('value'), d['whole_call']['frac_of_roofline'], (d.get('cpu_baseline') or {}).get('value')))
except Exception as e:
    print('config $c: no line', e)
PY
done
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench9_ref.json 2> gpurun_out/r2_bench9_ref.err
cat gpurun_out/r2_bench9_ref.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-micro --no-serving --ncu-range > gpurun_out/r2_ncu_launches.log 2>&1
tail -n 2 gpurun_out/r2_ncu_launches.log | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:decode_mega -s 5 -c 1 -f -o gpurun_out/r2_prof_mega python tools/one_call.py 64 > gpurun_out/r2_ncu_mega.log 2>&1
tail -n 2 gpurun_out/r2_ncu_mega.log | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:flash_attn_tc -s 3 -c 1 -f -o gpurun_out/r2_prof_fa python tools/one_call.py 64 > gpurun_out/r2_ncu_fa.log 2>&1
tail -n 2 gpurun_out/r2_ncu_fa.log | cut -c1-200
