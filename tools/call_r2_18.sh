#!/bin/bash
# Round 2, call 18: flash_attn_tc softmax rewrite (unit + feature parity, launch times), bench config 2, fine-grained decode timeline.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2_ktests18.log 2>&1
tail -n 3 gpurun_out/r2_ktests18.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "image_features or config2 or decisive" > gpurun_out/r2_tests18.log 2>&1
tail -n 3 gpurun_out/r2_tests18.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:flash_attn -c 20 --csv --log-file gpurun_out/r2_fa_launches18.csv python tools/one_call.py 64 > gpurun_out/r2_ncu18.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_fa_launches18.csv
timeout 200 python bench.py --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench18_c2.json 2> gpurun_out/r2_bench18_c2.err
tail -n 2 gpurun_out/r2_bench18_c2.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench18_c2.json
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild18.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline18.txt 2>&1
tail -n 60 gpurun_out/r2_mega_timeline18.txt
