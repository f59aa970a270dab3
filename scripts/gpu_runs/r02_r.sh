#!/bin/bash
# mma small-batch kernel with the activation pre-kernel: parity + timings; GEMM default (pair + relay) under the bench pattern; the full bench
mkdir -p gpurun_out
echo "== mma small-batch kernel: parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_batch" 2>&1 | tail -5
echo "== timings (dependent launches, weights from HBM)"
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0,q4_0 --shapes 11008x4096,4096x14336,4096x4096 --n 2,4,8 2>&1 | cut -c1-110
timeout 300 python scripts/gemv_sweep.py --types q4_K,q8_0 --shapes 11008x4096,4096x14336,4096x4096 --n 1 --mma 2>&1 | cut -c1-110
echo "== GEMM bench-pattern parity, default mode"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q6_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096" "q4_0 4096 128 4096"; do
  timeout 200 python scripts/gemm_bench_parity.py $sh 2>&1 | tail -3 | cut -c1-300
done
echo "== bench"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_r02_r.json 2> gpurun_out/bench_r02_r.err; echo "bench rc=$? after ${SECONDS}s"; tail -c 600 gpurun_out/bench_r02_r.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_r.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "parity", d["parity"])
    print("e2e", json.dumps(d["e2e"])[:900])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:420])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
