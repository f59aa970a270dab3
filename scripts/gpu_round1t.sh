#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backend_plugin.py -x -q -m gpu --timeout 600 2>&1 | tail -3
timeout 300 python scripts/gemv_sweep.py --types q4_K,q5_K,q6_K,q4_0,q8_0 --shapes 11008x4096,4096x4096,4096x11008,32000x4096 --both 2>&1 | cut -c1-110
timeout 300 python bench.py --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['roofline']['frac'], d['config'].get('dependent_chain'), d['e2e'])"
