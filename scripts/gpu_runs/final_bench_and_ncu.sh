#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/gemv_sweep.py --types q4_K,q8_0 --shapes 11008x4096 --n 2,3,4,8 2>&1 | cut -c1-120
timeout 300 python bench.py > gpurun_out/bench_final1.json 2> gpurun_out/bench_final1.err; tail -c 2500 gpurun_out/bench_final1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_v3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmvq_sb -s 30 -c 2 -o gpurun_out/prof_gemv_q4k_r01c -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc -s 0 -c 1 -o gpurun_out/prof_gemm_r01c -f python scripts/gemm_one.py q8_0 32000 512 4096 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
