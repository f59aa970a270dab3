#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python scripts/gemv_sweep.py --generic > gpurun_out/sweep_default.jsonl 2>gpurun_out/sweep_default.err
tail -3 gpurun_out/sweep_default.err; cat gpurun_out/sweep_default.jsonl | cut -c1-120
rm -f gpurun_out/sweep_tun.jsonl gpurun_out/sweep_tun.err
for cfg in "18 4 4" "18 3 8" "9 4 4" "36 3 8" "9 6 4" "18 4 8" "36 4 4" "27 3 4" "18 2 4"; do
  set -- $cfg
  GGML_B200_GEMV_STAGE_KB=$1 GGML_B200_GEMV_STAGES=$2 GGML_B200_GEMV_WARPS=$3 timeout 120 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x4096 >> gpurun_out/sweep_tun.jsonl 2>>gpurun_out/sweep_tun.err
done
cut -c1-30,60-250 gpurun_out/sweep_tun.jsonl
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmvq_tma -s 20 -c 2 -o gpurun_out/prof_gemv_q4k -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -15
