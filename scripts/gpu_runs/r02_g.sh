#!/bin/bash
# per-thread arrivals (both tcgen05 kernels): stress, parity, timings, cycle accounts
for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_BN=64" "GGML_B200_TC_SPLITK=3 GGML_B200_TC2_RAW=2" "GGML_B200_TC_PAIR=0"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 12 2>&1 | tail -14 | cut -c1-330
done
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm" 2>&1 | tail -3
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q8_0 32000 512 4096"; do
  for cfg in "GGML_B200_TC_PAIR=0" "X=0" "GGML_B200_TC2_BN=128"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
echo "== accounts"
for cfg in "X=0" "GGML_B200_TC2_BN=128"; do echo "-- $cfg"; env $cfg GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19; done
GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q4_K 4096 512 4096 --trace 2>&1 | tail -10
