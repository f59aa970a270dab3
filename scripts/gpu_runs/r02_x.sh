#!/bin/bash
# mma small-batch path for Q5_0 / Q4_1 / Q5_1 / IQ4_NL / IQ4_XS / Q2_K: parity + timings; split-buffer test; refreshed bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_formats.py -q -m gpu -x -k "small_batch or next" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_backend_plugin.py -q -m gpu -x -k "split_buffer or MUL_MAT or mul_mat" 2>&1 | tail -3
echo "== timings, new mma formats (and their dp4a times)"
timeout 300 python scripts/gemv_sweep.py --types q5_0,q4_1,q5_1,iq4_nl,iq4_xs,q2_K --shapes 4096x14336,11008x4096 --n 2,8 2>&1 | cut -c1-110
timeout 300 python scripts/gemv_sweep.py --types q5_0,q4_1,iq4_xs,q2_K --shapes 4096x14336 --n 8 --dp4a 2>&1 | cut -c1-110
echo "== bench"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_r02_x.json 2> gpurun_out/bench_r02_x.err; echo "bench rc=$? after ${SECONDS}s"; tail -c 400 gpurun_out/bench_r02_x.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_x.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "e2e", d["e2e"]["value"])
    for k, v in d["extra"].items():
        print(k, v.get("us_per_matvec", v.get("us_per_mul_mat", v.get("b200_ms_per_token"))), v.get("roofline", {}).get("frac"))
except Exception as e:
    print("no bench line:", e)
PY
