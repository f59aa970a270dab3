#!/bin/bash
# round 2 final single-GPU bundle: GEMM stress, the whole GPU suite, smoke, both bench arms, our side of the reference perf harness
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
echo "== GEMM stress (default modes, 30 launches per case)"
timeout 600 python tests/gpu_tc2_stress.py 30 2>&1 | grep -E "BAD|CLEAN|run " | cut -c1-250 | tail -12
GGML_B200_TC2_BN=128 timeout 600 python tests/gpu_tc2_stress.py 20 --big 2>&1 | grep -E "BAD|CLEAN" | cut -c1-200
echo "== suite"
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_r02_final.log; tail -25 gpurun_out/pytest_r02_final.log
timeout 120 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02_final.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_final.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "parity", d["parity"])
    print("e2e", json.dumps(d["e2e"])[:900])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:420])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_final_ref.json 2> gpurun_out/bench_r02_final_ref.err; echo "reference arm rc=$?"; tail -c 500 gpurun_out/bench_r02_final_ref.json
echo "== reference perf harness, this backend"
export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref:/usr/local/cuda/lib64
GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so timeout 900 oracle/_ref/test-backend-ops perf -o MUL_MAT -b B2000 > gpurun_out/tbo_perf_b200_final.log 2>&1; echo "rc=$?"
grep -E "type_a=(q4_0|q8_0|q4_K|q6_K|iq2_xxs|f16),type_b=f32,m=4096,n=(1|4|5|8|512),k=14336" gpurun_out/tbo_perf_b200_final.log | sed 's/  */ /g' | cut -c1-160
echo "== gpt-2 profile"
D=/tmp/ggml_b200_gpt2_v2
GGML_B200_PROFILE=1 timeout 200 oracle/_ref/gpt-2-backend-b200 -m $D/gpt2_q4_0.bin -s 1234 -n 128 -t 8 --ignore-eos --top_k 1 -p "a b c" -ngl 12 2>&1 | grep -E "profile|predict time|sample time"
