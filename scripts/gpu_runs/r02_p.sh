#!/bin/bash
# round 2, session 2 baseline: GEMM stress (short), both bench arms, GEMM timings + trace, small-n sweep, launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
echo "== GEMM stress (default modes, 10 launches per case)"
timeout 400 python tests/gpu_tc2_stress.py 10 2>&1 | grep -E "BAD|CLEAN|run " | cut -c1-250 | tail -8
echo "== bench"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_r02_p.json 2> gpurun_out/bench_r02_p.err; echo "bench rc=$? after ${SECONDS}s"; tail -c 600 gpurun_out/bench_r02_p.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_p.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "parity", d["parity"])
    print("e2e", json.dumps(d["e2e"])[:900])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:420])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_p_ref.json 2> gpurun_out/bench_r02_p_ref.err; echo "reference arm rc=$? after ${SECONDS}s"; tail -c 500 gpurun_out/bench_r02_p_ref.json
echo "== GEMM timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q6_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096" "q8_0 4096 128 4096" "q4_0 4096 512 14336"; do
  timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1
done
GGML_B200_TC2_SOLO=0 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1
GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19
echo "== small n"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336 --n 1,2,4,8 2>&1 | cut -c1-140 | tail -30
echo "== ncu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmq_tc2 -s 8 -c 1 -o gpurun_out/r02_gemm_q8_0_solo -f python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > gpurun_out/ncu_gemm.log 2>&1; echo "gemm capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmvq_sb -s 40 -c 2 -o gpurun_out/r02_gemv_q4k -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_gemv.log 2>&1; echo "gemv capture rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
