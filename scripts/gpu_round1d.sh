#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep_tun.jsonl gpurun_out/sweep_tun.err
for cfg in "36 2 1 0" "36 2 1 1" "36 3 1 0" "36 2 2 0" "18 4 1 0" "18 5 1 0" "72 2 1 0"; do
  set -- $cfg
  GGML_B200_SB_STAGE_KB=$1 GGML_B200_SB_STAGES=$2 GGML_B200_SB_CTAS=$3 GGML_B200_NO_PDL=$4 timeout 120 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x4096,32000x4096 >> gpurun_out/sweep_tun.jsonl 2>>gpurun_out/sweep_tun.err
done
tail -3 gpurun_out/sweep_tun.err
timeout 300 python scripts/gemm_sweep.py > gpurun_out/gemm_sweep.jsonl 2> gpurun_out/gemm_sweep.err; tail -3 gpurun_out/gemm_sweep.err; cat gpurun_out/gemm_sweep.jsonl
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
