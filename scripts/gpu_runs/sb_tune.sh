#!/bin/bash
# tunable sweep of the superblock mat-vec kernel (developer tool): WARPS CTAS RESIDENT STAGE_KB STAGES
mkdir -p gpurun_out; out=gpurun_out/sb_tune.jsonl; : > $out
while read w c r kb st; do
  [ -z "$w" ] && continue
  GGML_B200_SB_WARPS=$w GGML_B200_SB_CTAS=$c GGML_B200_SB_RESIDENT=$r GGML_B200_SB_STAGE_KB=$kb GGML_B200_SB_STAGES=$st \
    timeout 120 python scripts/gemv_sweep.py --types ${TYPES:-q4_K} --shapes ${SHAPES:-11008x4096} --both >> $out 2>gpurun_out/sb_tune.err || echo "{\"fail\": \"$w $c $r $kb $st\"}" >> $out
done <<CFG
8 1 2 36 0
8 2 2 36 0
8 1 1 72 2
4 1 2 18 0
4 1 3 18 0
4 1 4 18 0
4 2 2 18 0
4 2 3 18 0
4 2 4 18 0
4 3 3 18 0
4 4 4 18 0
4 2 2 36 0
4 1 2 36 0
CFG
python - <<'PY'
import json
for l in open('gpurun_out/sb_tune.jsonl'):
    d=json.loads(l)
    if 'fail' in d: print(d); continue
    t=d['tun']; print(d['kernel'], d['us'], d['GBps'], {k[13:]:v for k,v in t.items()})
PY
