#!/bin/bash
# round-robin tile / chunk hand-out instead of one global atomic per stage: mma small-batch kernel and the dependent n = 1 mat-vec
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_batch or matvec or full_size" 2>&1 | tail -4
echo "== mma, static (16 warps / 8 warps) and dynamic"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336,4096x4096 --n 2,8 2>&1 | cut -c1-110
GGML_B200_MMA_WARPS=8 timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336,4096x4096 --n 8 2>&1 | cut -c1-140
GGML_B200_MMA_DYNAMIC=1 timeout 300 python scripts/gemv_sweep.py --types q4_K --shapes 11008x4096,4096x14336 --n 8 2>&1 | cut -c1-140
echo "== n = 1 dp4a kernel: dynamic vs round-robin (dependent / independent)"
for st in 0 1; do GGML_B200_SB_STATIC=$st timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0,q6_K --shapes 11008x4096,4096x4096,32000x4096 --n 1 --both 2>&1 | cut -c1-140; done
echo "== n = 1 on the mma kernel"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0 --shapes 11008x4096,4096x14336,4096x4096 --n 1 --mma 2>&1 | cut -c1-110
