#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 2" "0 1"; do set -- $cfg
  GGML_B200_SB_DEBUG=1 GGML_B200_NO_PDL=$1 STATIC=$2 timeout 60 python scripts/pdl_trace.py 2>&1 | tail -15
done
timeout 900 python -m pytest tests -q -m gpu --timeout 180 -x > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_full.log | tail -12
grep -E "^E  " gpurun_out/pytest_gpu_full.log | cut -c1-260 | head -20
for cfg in "36 2 1" "18 4 1" "18 3 1" "36 2 2" "18 2 1"; do set -- $cfg
  GGML_B200_SB_STAGE_KB=$1 GGML_B200_SB_STAGES=$2 GGML_B200_SB_CTAS=$3 timeout 120 python scripts/gemv_sweep.py --independent --types q4_K,q8_0 --shapes 11008x4096,4096x4096,32000x4096 2>&1 | cut -c1-100,140-260
done
timeout 300 python bench.py 2>gpurun_out/bench_default.err > gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-1500
bash scripts/gpt2_bench.sh 2>&1 | tail -14
