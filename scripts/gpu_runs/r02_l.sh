#!/bin/bash
# GEMM: solo mode (cta_group::1 per CTA + activation multicast) vs pair mode (relay): stress, parity, timings, accounts; small-n dispatch check
for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_BN=64" "GGML_B200_TC_SPLITK=3"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 12 2>&1 | grep -v "^ok" | tail -8 | cut -c1-330
done
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -1
GGML_B200_MMID_GROUPED=1 timeout 300 python tests/gpu_mmid_grouped_check.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm" 2>&1 | grep -E "assert|Error|passed|failed|^E " | head -20
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q6_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096" "q8_0 4096 128 4096"; do
  for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_SOLO=0" "GGML_B200_TC_PAIR=0"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
GEMM_PROF_NBUF=16 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1
for cfg in "GGML_B200_TC_SPLITK=1" "GGML_B200_TC2_BN=128 GGML_B200_TC_SPLITK=2" "GGML_B200_TC2_STAGES=4 GGML_B200_TC2_RAW=2" "GGML_B200_TC2_TMA_EPI=0"; do env $cfg timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1; done
echo "== accounts"
GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19
GGML_B200_TC2_BN=128 GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19
echo "== n = 5..8 long rows now on tensor cores"
timeout 200 python scripts/gemv_sweep.py --types q4_K,q8_0 --shapes 4096x14336 --n 4,5,8 2>&1 | cut -c1-120 | tail -7
