#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 1" "1 1" "0 0"; do set -- $cfg
  GGML_B200_SB_DEBUG=1 GGML_B200_NO_PDL=$1 STATIC=$2 timeout 60 python scripts/pdl_trace.py 2>&1 | tail -16
done
timeout 900 python -m pytest tests -q -m gpu --timeout 180 > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_full.log | tail -20
grep -E "assert|Error|cpu:|gpu:" gpurun_out/pytest_gpu_full.log | cut -c1-260 | head -40
bash scripts/gpt2_bench.sh 2>&1 | tail -20
