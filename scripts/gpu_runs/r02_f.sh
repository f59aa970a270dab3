#!/bin/bash
# pair-kernel race hunt, second pass: MMA delay, per-thread arrivals, raw-ring cross-check; one-CTA kernel in the MMA-starved regime; timeline trace
for cfg in "GGML_B200_TC_PAIR=0" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=8" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=16" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=32 GGML_B200_TC2_TRACE=1" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=24" "GGML_B200_TC2_BN=128 GGML_B200_TC2_STAGES=2" "GGML_B200_TC2_BN=128 GGML_B200_TC2_STAGES=8" "X=0"; do
  echo "-- $cfg"
  env $cfg timeout 200 python tests/gpu_tc2_stress.py 10 --big 2>&1 | tail -22 | cut -c1-330
done
echo "== timeline"
for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=1"; do echo "-- $cfg"; env $cfg GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -10; done
GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 32000 512 4096 --trace 2>&1 | tail -10
