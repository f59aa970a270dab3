#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "gemv or golden or full_size or fused" 2>&1 | tail -3
export GGML_B200_SB_DEBUG=1
while read w c r kb st; do
  [ -z "$w" ] && continue
  GGML_B200_SB_WARPS=$w GGML_B200_SB_CTAS=$c GGML_B200_SB_RESIDENT=$r GGML_B200_SB_STAGE_KB=$kb GGML_B200_SB_STAGES=$st STATIC=1 timeout 100 python scripts/pdl_trace.py 2>&1 | grep SUMMARY
done <<CFG
8 1 2 36 2
8 1 1 36 3
8 1 1 36 5
8 2 2 36 2
4 2 2 18 5
4 2 4 18 2
4 4 4 18 2
4 1 2 18 5
CFG
unset GGML_B200_SB_DEBUG
timeout 200 python scripts/gemv_sweep.py --types q4_K --shapes 11008x4096,4096x11008 --both 2>&1 | cut -c1-120
