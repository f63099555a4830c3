#!/bin/bash
# Round 2, call 20: decode attention dealt by 64-key units over all 8 warps + one LayerNorm row per CTA; trie / sampling
# decoders -- full GPU suite, bench config 2 / 4, timeline.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests20.log 2>&1
tail -n 4 gpurun_out/r2_tests20.log
timeout 200 python bench.py --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench20_c2.json 2> gpurun_out/r2_bench20_c2.err
tail -n 2 gpurun_out/r2_bench20_c2.err | cut -c1-200; cut -c1-330 gpurun_out/r2_bench20_c2.json
timeout 200 python bench.py --config 4 --no-cpu-baseline --no-micro --no-serving > gpurun_out/r2_bench20_c4.json 2> gpurun_out/r2_bench20_c4.err
cut -c1-330 gpurun_out/r2_bench20_c4.json
GITB200_TIMELINE=1 timeout 200 python -c "from generativeimage2text_b200 import build; build.build(force=True)" > gpurun_out/r2_tlbuild20.log 2>&1
timeout 120 python tools/mega_timeline.py > gpurun_out/r2_mega_timeline20.txt 2>&1
tail -n 58 gpurun_out/r2_mega_timeline20.txt | cut -c1-90
