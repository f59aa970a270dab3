#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 180 > gpurun_out/pytest_gpu_full.log 2>&1
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu_full.log | tail -8
grep -E "^E  " gpurun_out/pytest_gpu_full.log | cut -c1-240 | head -12
bash scripts/gpt2_bench.sh 2>&1 | tail -9 | cut -c1-160
echo "== fusion off"; D=/tmp/ggml_b200_gpt2_v2; LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref GGML_B200_DISABLE_FUSION=1 oracle/_ref/gpt-2-backend-b200 -m $D/gpt2_q4_0.bin -s 1234 -n 128 --ignore-eos --top_k 1 -p a_b_c -t 8 -ngl 12 2>&1 | grep -E "predict time|^a_b_c" | cut -c1-160
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; python -c "
import json;d=json.load(open('gpurun_out/bench_default.json'));print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']),'dep',d['config']['dependent_chain'],'cpu',d['cpu_baseline']['value'])"
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 | cut -c1-300
