#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "gemm or golden" 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -30
for h in 0 1 2; do echo "== HALVES=$h"; GGML_B200_TC_HALVES=$h timeout 300 python scripts/gemm_sweep.py 2>&1 | cut -c1-120; done
