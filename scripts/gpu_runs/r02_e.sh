#!/bin/bash
# pair-kernel race hunt: repeatability / exactness stress under kernel variants
for cfg in "GGML_B200_TC_PAIR=0" "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=3 GGML_B200_TC2_RAW=2" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=1" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=2" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=4" "GGML_B200_TC2_BN=128 GGML_B200_TC2_DBG=7" "GGML_B200_TC_SPLITK=3 GGML_B200_TC2_RAW=2 GGML_B200_TC2_DBG=7" "GGML_B200_TC2_BN=128 GGML_B200_NO_PDL=1" "GGML_B200_TC2_BN=64"; do
  echo "-- $cfg"
  env $cfg timeout 200 python tests/gpu_tc2_stress.py 8 2>&1 | tail -14
done
