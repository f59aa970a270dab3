#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gemv_sweep.py --v1 > gpurun_out/sweep_default.jsonl 2>gpurun_out/sweep_default.err
tail -3 gpurun_out/sweep_default.err; cut -c1-120 gpurun_out/sweep_default.jsonl
rm -f gpurun_out/sweep_tun.jsonl gpurun_out/sweep_tun.err
for cfg in "36 2 2 0" "36 2 2 1" "18 4 2 0" "18 3 2 0" "36 3 1 0" "54 2 1 0" "18 6 2 0" "72 2 1 0"; do
  set -- $cfg
  GGML_B200_SB_STAGE_KB=$1 GGML_B200_SB_STAGES=$2 GGML_B200_SB_CTAS=$3 GGML_B200_NO_PDL=$4 timeout 120 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x4096 >> gpurun_out/sweep_tun.jsonl 2>>gpurun_out/sweep_tun.err
done
tail -3 gpurun_out/sweep_tun.err
cut -c1-30,60-250 gpurun_out/sweep_tun.jsonl
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmvq_sb -s 20 -c 2 -o gpurun_out/prof_gemv_q4k_v2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
