#!/bin/bash
# pair GEMM v7 (parallel relay lanes, split cluster sync): stress, failing q6_K test verbose, timings (L2-resident and streaming), accounts
for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_BN=64"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 16 --big 2>&1 | grep -v "^ok" | tail -8 | cut -c1-330
done
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm" 2>&1 | grep -E "assert|Error|passed|failed|^E " | head -20
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q6_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096" "q8_0 4096 128 4096"; do
  timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1
done
GEMM_PROF_NBUF=16 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1
GEMM_PROF_NBUF=16 GGML_B200_TC_PAIR=0 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1
GEMM_PROF_NBUF=28 timeout 120 python scripts/gemm_prof.py q4_K 4096 512 4096 2>&1 | tail -1
for cfg in "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=1" "GGML_B200_TC2_STAGES=3" "GGML_B200_TC2_STAGES=5"; do env $cfg timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1; done
echo "== accounts"
GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19
