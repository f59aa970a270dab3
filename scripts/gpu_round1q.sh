#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 300 -k "gemv or golden or full_size or fused" 2>&1 | tail -3
out=gpurun_out/sb_tune2.jsonl; : > $out
for cfg in "48 0" "0 0" "48 1"; do set -- $cfg
  GGML_B200_SB_L2_MB=$1 SWEEP_NO_SRC0_STATIC=$2 timeout 200 python scripts/gemv_sweep.py --types q4_K,q8_0,q6_K --shapes 11008x4096,4096x4096,4096x11008,32000x4096 --both >> $out 2>gpurun_out/sb_tune2.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sb_tune2.jsonl'):
    d=json.loads(l); t=d['tun']; print(d['type'], d['M'], d['K'], d['kernel'], d['us'], d['GBps'], {k[10:]:v for k,v in t.items()})
PY
timeout 300 python bench.py --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['roofline']['frac'], d['config'].get('dependent_chain'), d['e2e'])"
bash scripts/gpt2_bench.sh 2>&1 | tail -8
