#!/bin/bash
# mma small-batch kernel, 16 consumer warps (vs 8) + in-place high nibbles: parity, timings, one full ncu capture; Q6_K GEMM under the bench pattern
mkdir -p gpurun_out
echo "== mma small-batch kernel: parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_batch" 2>&1 | tail -4
echo "== timings, 16 warps"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0,q5_K --shapes 11008x4096,4096x14336,4096x4096 --n 2,8 2>&1 | cut -c1-110
echo "== timings, 8 warps"
GGML_B200_MMA_WARPS=8 timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0 --shapes 11008x4096,4096x14336 --n 8 2>&1 | cut -c1-140
echo "== ncu, q4_K 11008 x 8 x 4096"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmvq_mma -s 5 -c 1 -o gpurun_out/r02_mma_q4k_n8 -f python scripts/mma_one.py q4_K 11008 8 4096 > gpurun_out/ncu_mma.log 2>&1; echo "capture rc=$?"
echo "== Q6_K GEMM, bench pattern"
for cfg in "X=0 --serial" "GGML_B200_NO_PDL=1" "GGML_B200_TC_PAIR=0 --serial" "GGML_B200_TC2_DBG=4"; do
  set -- $cfg; env $1 timeout 200 python scripts/gemm_bench_parity.py q6_K 4096 512 4096 ${@:2} 2>&1 | tail -6 | cut -c1-300
done
timeout 200 python scripts/gemm_bench_parity.py q5_K 4096 512 4096 2>&1 | tail -2 | cut -c1-300
