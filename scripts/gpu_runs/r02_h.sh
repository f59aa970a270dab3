#!/bin/bash
# pair GEMM v3 (no run-time divisions, merged fences, conflict-free Q8_0 unit loads) with TMA / cp.async raw producers: stress, timings, accounts
for cfg in "X=0" "GGML_B200_TC2_RAWMODE=1" "GGML_B200_TC2_BN=128 GGML_B200_TC2_RAWMODE=1" "GGML_B200_TC2_BN=128"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 8 2>&1 | grep -v "^ok" | tail -8 | cut -c1-330
done
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -1
GGML_B200_TC2_RAWMODE=1 timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -1
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q6_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096"; do
  for cfg in "X=0" "GGML_B200_TC2_RAWMODE=1" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_BN=128 GGML_B200_TC2_RAWMODE=1"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
echo "== accounts"
for cfg in "X=0" "GGML_B200_TC2_RAWMODE=1" "GGML_B200_TC2_BN=128 GGML_B200_TC2_RAWMODE=1"; do echo "-- $cfg"; env $cfg GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19; done
GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q4_K 4096 512 4096 --trace 2>&1 | tail -19
