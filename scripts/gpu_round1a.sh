#!/bin/bash
# first GPU contact: parity tests, then a quick sweep of the mat-vec kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -15
timeout 300 python scripts/gemv_sweep.py --generic > gpurun_out/sweep_default.jsonl 2>gpurun_out/sweep_default.err
tail -3 gpurun_out/sweep_default.err
cat gpurun_out/sweep_default.jsonl
for cfg in "18 4 4" "18 3 8" "9 4 4" "36 3 8" "9 6 4" "18 4 8"; do
  set -- $cfg
  GGML_B200_GEMV_STAGE_KB=$1 GGML_B200_GEMV_STAGES=$2 GGML_B200_GEMV_WARPS=$3 timeout 120 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x4096 >> gpurun_out/sweep_tun.jsonl 2>>gpurun_out/sweep_tun.err
done
cat gpurun_out/sweep_tun.jsonl
