#!/bin/bash
# (1) why does the bench's GEMM leg fail parity while the stress test is clean: the bench's exact launch pattern, every element checked, per kernel variant
# (2) first run of the int8 mma.sync small-batch kernel (mmvq_mma.cu): parity tests + timings against the dp4a path
mkdir -p gpurun_out
echo "== GEMM bench-pattern parity"
for cfg in "X=0" "X=0 --plain" "X=0 --noflags" "GGML_B200_NO_PDL=1" "GGML_B200_TC2_SOLO=0" "GGML_B200_TC_PAIR=0"; do
  set -- $cfg; env $1 timeout 200 python scripts/gemm_bench_parity.py q8_0 4096 512 4096 ${@:2} 2>&1 | tail -4 | cut -c1-330
done
timeout 200 python scripts/gemm_bench_parity.py q4_K 4096 512 4096 2>&1 | tail -3 | cut -c1-330
echo "== mma small-batch kernel: parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_batch" 2>&1 | tail -15
echo "== mma small-batch kernel: timings (dependent launches, weights from HBM)"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336,4096x4096 --n 2,4,8 2>&1 | cut -c1-150
echo "-- n = 1 on the mma kernel (dependent / independent) vs the dp4a kernel"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336,4096x4096 --n 1 --mma --both 2>&1 | cut -c1-150
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336,4096x4096 --n 1 --both 2>&1 | cut -c1-150
