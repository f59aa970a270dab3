#!/bin/bash
# round 2 final single-GPU bundle: small-batch parity (n = 1 long rows now on the mma kernel), both bench arms, launch list of the bench command,
# this backend in the reference's perf harness (the reference CUDA backend's numbers for the same cases: profiles/r02_tbo_perf_reference_cuda.log)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backend_plugin.py -q -m gpu -x 2>&1 | tail -3
echo "== bench"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_r02_w.json 2> gpurun_out/bench_r02_w.err; echo "bench rc=$? after ${SECONDS}s"; tail -c 600 gpurun_out/bench_r02_w.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_w.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "parity", d["parity"], "launches", d.get("gpu_launches"))
    print("e2e", json.dumps(d["e2e"])[:900])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:330])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
SECONDS=0
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_w_ref.json 2> gpurun_out/bench_r02_w_ref.err; echo "reference arm rc=$? after ${SECONDS}s"; tail -c 400 gpurun_out/bench_r02_w_ref.json
echo "== launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --sweeps-per-step 2 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
echo "== reference perf harness, this backend"
export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref:/usr/local/cuda/lib64
GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so timeout 900 oracle/_ref/test-backend-ops perf -o MUL_MAT -b B2000 > gpurun_out/r02_tbo_perf_b200_final.log 2>&1; echo "rc=$?"
grep -E "type_a=(q4_0|q8_0|q4_K|q5_K|q6_K),type_b=f32,m=4096,n=(1|2|4|8|512),k=14336" gpurun_out/r02_tbo_perf_b200_final.log | sed 's/  */ /g' | cut -c1-150
