#!/bin/bash
# round 2: the whole GPU suite, smoke, and the full bench line (all BASELINE configurations + parity + plug-in e2e)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
nproc
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -12
timeout 120 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02a.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02a.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "parity", d["parity"])
    print("e2e", json.dumps(d["e2e"])[:900])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:420])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
