#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backend_plugin.py -q -m gpu --timeout 600 -k "gemv or small_batch or golden or MUL_MAT or mul_mat or plugin" 2>&1 | grep -E "^E  |^FAILED|passed|failed" | head -20
timeout 200 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0,q6_K --shapes 11008x4096 --n 2,3,4,8 2>&1 | cut -c1-120
