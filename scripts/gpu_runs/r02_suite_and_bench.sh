#!/bin/bash
# round 2: pair-kernel sanity (isolated), the whole GPU suite, smoke, the full bench line (all BASELINE configurations + parity + plug-in e2e),
# dependent-chain experiments, and the reference's own CUDA backend through the reference's perf harness
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
nproc
echo "== pair kernel check"
if timeout 150 python tests/gpu_tc2_check.py --time > gpurun_out/tc2_check.log 2>&1; then
  tail -14 gpurun_out/tc2_check.log
  for cfg in "128 0" "256 1" "128 2" "64 0"; do set -- $cfg; GGML_B200_TC2_BN=$1 GGML_B200_TC_SPLITK=$2 timeout 100 python tests/gpu_tc2_check.py --time 2>&1 | grep "^time"; done
  for st in 3 4; do GGML_B200_TC2_STAGES=$st timeout 100 python tests/gpu_tc2_check.py --time 2>&1 | grep "^time"; done
else
  echo "PAIR KERNEL CHECK FAILED (rc=$?): continuing with GGML_B200_TC_PAIR=0"; tail -25 gpurun_out/tc2_check.log
  export GGML_B200_TC_PAIR=0
  nvidia-smi --query-gpu=name --format=csv,noheader | head -1
fi
echo "== suite"
timeout 2000 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_r02.log; tail -30 gpurun_out/pytest_r02.log
timeout 120 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02a.err; python - <<'PY'
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
echo "== dependent-chain variants"
bash scripts/gpu_runs/r02_dep_variants.sh 2>&1 | tail -60
echo "== reference CUDA backend vs this backend, reference harness"
bash scripts/gpu_runs/r02_reference_cuda_perf.sh 2>&1 | tail -50
