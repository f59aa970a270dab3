#!/bin/bash
# final round-1 profiles of the bench command: launch list + full capture of the dominant kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_v2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmvq_sb -s 30 -c 2 -o gpurun_out/prof_gemv_q4k_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01_v2.csv
