#!/bin/bash
export GGML_B200_SB_DEBUG=1
echo "--- dependent, L2 prefetch on"; STATIC=1 timeout 100 python scripts/pdl_trace.py 2>&1 | tail -14
echo "--- dependent, L2 prefetch off"; GGML_B200_SB_L2_MB=0 STATIC=1 timeout 100 python scripts/pdl_trace.py 2>&1 | tail -14
echo "--- independent"; STATIC=2 timeout 100 python scripts/pdl_trace.py 2>&1 | tail -14
