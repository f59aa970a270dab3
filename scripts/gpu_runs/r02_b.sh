#!/bin/bash
# round 2, second GPU bundle: node-by-node gpt-2 graph parity, pair-kernel timings, grouped MUL_MAT_ID, full suite with the pair kernel on, bench, gpt-2 launch list
mkdir -p gpurun_out
export LD_LIBRARY_PATH_SAVE=$LD_LIBRARY_PATH
D=/tmp/ggml_b200_gpt2_v2; mkdir -p $D
[ -f $D/gpt2_q4_0.bin ] || { python scripts/make_gpt2_synth.py $D/gpt2_f16.bin > /dev/null; LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref oracle/_ref/gpt-2-quantize $D/gpt2_f16.bin $D/gpt2_q4_0.bin 2 > /dev/null 2>&1; rm -f $D/gpt2_f16.bin; }
echo "== gpt-2 graph, node by node (CPU vs B200)"
LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so timeout 300 oracle/_ref/gpt2-compare $D/gpt2_q4_0.bin B2000 5 > gpurun_out/gpt2_compare.log 2>gpurun_out/gpt2_compare.err
echo "rc=$?"; grep summary gpurun_out/gpt2_compare.log; awk '$NF+0 > 1e-9' gpurun_out/gpt2_compare.log | head -40
echo "== pair kernel check + timings"
timeout 200 python tests/gpu_tc2_check.py --time 2>&1 | tail -16
for cfg in "128 0" "256 1" "128 2" "64 0"; do set -- $cfg; GGML_B200_TC2_BN=$1 GGML_B200_TC_SPLITK=$2 timeout 100 python tests/gpu_tc2_check.py --time 2>&1 | grep "^time"; done
for st in 3 4; do GGML_B200_TC2_STAGES=$st timeout 100 python tests/gpu_tc2_check.py --time 2>&1 | grep "^time"; done
GGML_B200_TC_PAIR=0 timeout 100 python -c "
import os, sys; sys.argv=['x','--time']
" ; echo "== grouped mul_mat_id"
GGML_B200_MMID_GROUPED=1 timeout 200 python tests/gpu_mmid_grouped_check.py --time 2>&1 | tail -8
echo "== suite"
timeout 2000 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_r02b.log; tail -25 gpurun_out/pytest_r02b.log
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r02b.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02b.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "clocks", d["clocks"], "parity", d["parity"])
    print("e2e", json.dumps(d["e2e"])[:1200])
    for k, v in d["extra"].items():
        print(k, json.dumps(v)[:420])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
echo "== gpt-2 launch list (ncu, 6 tokens)"
LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_gpt2_launches.csv oracle/_ref/gpt-2-backend-b200 -m $D/gpt2_q4_0.bin -s 1234 -n 6 -t 8 --ignore-eos --top_k 1 -p "a b c" -ngl 12 > gpurun_out/ncu_gpt2.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_gpt2_launches.csv")) if len(r) > 10 and r[0].isdigit()]
names = collections.Counter(); dur = collections.Counter()
for r in rows:
    k = r[4].split("(")[0][:60]; names[k] += 1
    try: dur[k] += float(r[-1])
    except: pass
print("launches", len(rows))
for k, c in names.most_common(20): print(c, k, round(dur[k] / 1000, 1), "us total")
PY
