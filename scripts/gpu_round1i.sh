#!/bin/bash
mkdir -p gpurun_out
NP=2
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NP --steps 200 --warmup 10 --exchange fused 2>gpurun_out/bench${NP}_fused.err | tail -1 > gpurun_out/bench${NP}_fused.json; grep -v "^W0\|\*\*\*\|OMP" gpurun_out/bench${NP}_fused.err | tail -4; python -c "
import json;d=json.load(open('gpurun_out/bench${NP}_fused.json'));print('fused', d['n_gpus'], 'value', round(d['value']), 'GB/s  us/matvec', round(d['roofline']['us_per_launch'],2), d['config']['parallelism'])"
fi
for sk in 1 2 4; do echo "== splitk $sk"; GGML_B200_TC_SPLITK=$sk timeout 100 python scripts/gemm_sweep.py 2>&1 | grep -E '"N": 512' | grep -E "q8_0|q4_K" | head -6; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc -s 2 -c 1 -o gpurun_out/prof_gemm_q8 -f python scripts/gemm_sweep.py > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log | cut -c1-200
