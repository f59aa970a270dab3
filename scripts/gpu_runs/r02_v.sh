#!/bin/bash
# mma small-batch kernel: two independent consumer groups per CTA (vs one), Q6_K on the mma path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_batch" 2>&1 | tail -4
GGML_B200_MMA_GROUPS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_batch_mma" 2>&1 | tail -2
echo "== two groups"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0,q5_K,q6_K --shapes 11008x4096,4096x14336,4096x4096 --n 2,8 2>&1 | cut -c1-110
echo "== one group"
GGML_B200_MMA_GROUPS=1 timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q6_K --shapes 11008x4096,4096x14336,4096x4096 --n 8 2>&1 | cut -c1-140
echo "== n = 1 on the mma kernel, two groups"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q6_K --shapes 11008x4096,4096x14336,4096x4096 --n 1 --mma 2>&1 | cut -c1-110
echo "== q6_K dp4a reference"
timeout 300 python scripts/gemv_sweep.py --types q6_K --shapes 11008x4096,4096x14336 --n 2,8 --dp4a 2>&1 | cut -c1-110
