#!/bin/bash
mkdir -p gpurun_out
NP=2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 200 --warmup 10 --exchange fused 2>gpurun_out/bench${NP}_fused.err | tail -1 > gpurun_out/bench${NP}_fused.json; grep -v "^W0\|\*\*\*\|OMP" gpurun_out/bench${NP}_fused.err | tail -4; python -c "
import json;d=json.load(open('gpurun_out/bench${NP}_fused.json'));print('fused', d['n_gpus'], 'value', round(d['value']), 'GB/s  us/matvec', round(d['roofline']['us_per_launch'],2))"
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -k "gemm or golden" 2>&1 | tail -5
timeout 100 python scripts/gemm_sweep.py 2>&1 | grep -E '"N": (512|128)' | cut -c1-120
