#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpt2_bench.sh 2>&1 | tail -12
echo "== graphs disabled"; D=/tmp/ggml_b200_gpt2_v2; LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref GGML_B200_DISABLE_GRAPHS=1 oracle/_ref/gpt-2-backend-b200 -m $D/gpt2_q4_0.bin -s 1234 -n 128 --ignore-eos --top_k 1 -p a_b_c -t 8 -ngl 12 2>&1 | grep -E "predict time"
timeout 900 python -m pytest tests -q -m gpu --timeout 180 > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_full.log | tail -8
grep -E "^E  " gpurun_out/pytest_gpu_full.log | cut -c1-240 | head -10
timeout 100 python scripts/gemm_sweep.py 2>&1 | grep -E '"N": 512' | cut -c1-120
