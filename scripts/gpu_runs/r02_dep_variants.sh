#!/bin/bash
# dependent-chain mat-vec experiments (Q4_K 4096 -> 11008): two rows per lane group sharing the activation loads, warps / residency / stage size
for cfg in "X=0" "GGML_B200_SB_WARPS=4" "GGML_B200_SB_TWOROW=1 GGML_B200_SB_WARPS=4" "GGML_B200_SB_TWOROW=1 GGML_B200_SB_WARPS=4 GGML_B200_SB_STAGE_KB=72" \
           "GGML_B200_SB_TWOROW=1" "GGML_B200_SB_TWOROW=1 GGML_B200_SB_WARPS=4 GGML_B200_SB_RESIDENT=3 GGML_B200_SB_STAGE_KB=18" "GGML_B200_SB_WARPS=4 GGML_B200_SB_RESIDENT=3 GGML_B200_SB_STAGE_KB=18" \
           "GGML_B200_SB_L2_MB=0" "GGML_B200_SB_TWOROW=1 GGML_B200_SB_WARPS=4 GGML_B200_SB_L2_MB=0"; do
  echo "== $cfg"
  env $cfg timeout 120 python scripts/gemv_sweep.py --types q4_K --shapes 11008x4096 --both 2>&1 | grep -v "^$" | cut -c1-160 | tail -4
done
