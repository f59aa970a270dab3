#!/bin/bash
NP=$(nvidia-smi -L | wc -l)
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 50 --warmup 5 --exchange fused --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['n_gpus'], 'value', round(d['value']), 'GB/s  us/matvec', round(d['roofline']['us_per_launch'],2))"; }
run default
GGML_B200_SB_WARPS=8 GGML_B200_SB_RESIDENT=2 GGML_B200_SB_STAGE_KB=36 run w8_res2
GGML_B200_SB_WARPS=4 GGML_B200_SB_RESIDENT=3 GGML_B200_SB_STAGE_KB=18 run w4_res3
GGML_B200_SB_WARPS=4 GGML_B200_SB_RESIDENT=2 GGML_B200_SB_STAGE_KB=18 run w4_res2
