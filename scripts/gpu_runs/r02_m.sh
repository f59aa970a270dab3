#!/bin/bash
# solo-mode race localisation (q8_0 / q4_K 4096 x 512 x 4096 and friends, 10 launches each)
for cfg in "X=0" "GGML_B200_TC2_DBG=1" "GGML_B200_TC2_DBG=2" "GGML_B200_TC2_DBG=4" "GGML_B200_TC2_TMA_EPI=0" "GGML_B200_TC2_DBG=7 GGML_B200_TC2_TMA_EPI=0" "GGML_B200_TC2_RAW=1" "GGML_B200_NO_PDL=1" "GGML_B200_TC2_SOLO=0"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 10 --big 2>&1 | grep -E "BAD|CLEAN|raw-ring" | cut -c1-200
done
echo "== dense fp16 path: next-formats check (i-quants n >= 9 on tensor cores) and f16 weights through the plug-in"
timeout 900 python -m pytest tests/test_gpu_next_formats.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_backend_plugin.py -q -m gpu -x -k "mul_mat or MUL_MAT" 2>&1 | tail -4
