#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_config4_vatex_batch16_against_reference \
  --deselect tests/test_gpu_parity.py::test_config3_large_beam_batch32_against_reference_trajectory \
  --deselect tests/test_gpu_parity.py::test_decisive_checkpoint_free_running_token_identity \
  --deselect "tests/test_gpu_parity.py::test_parity_mode_logits_within_1e3_of_the_fp32_reference[base_decisive]" \
  -s > gpurun_out/r2_tests2.log 2>&1
tail -n 30 gpurun_out/r2_tests2.log
timeout 600 python bench.py --steps 16 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -n 3 gpurun_out/r2_bench2.err; cat gpurun_out/r2_bench2.json
