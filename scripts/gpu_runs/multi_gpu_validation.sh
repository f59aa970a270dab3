#!/bin/bash
mkdir -p gpurun_out
NP=$(nvidia-smi -L | wc -l)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29511 tests/gpu_multi_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -9
for ex in fused nccl; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 50 --warmup 5 --exchange $ex 2>gpurun_out/bench${NP}_$ex.err | tail -1 > gpurun_out/bench${NP}_$ex.json; grep -v "^W0\|\*\*\*\|OMP" gpurun_out/bench${NP}_$ex.err | tail -4; python -c "
import json;d=json.load(open('gpurun_out/bench${NP}_$ex.json'));print('$ex', d['n_gpus'], 'value', round(d['value']), 'GB/s  us/matvec', round(d['roofline']['us_per_launch'],2))"
done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $NP --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-200
