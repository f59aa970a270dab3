#!/bin/bash
# 2-GPU run: fused NVLink gather vs NCCL all-gather; then the bench at N=2 in both exchange modes
mkdir -p gpurun_out
nvidia-smi topo -m | head -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 2>gpurun_out/bench2_fused.err | tail -1 > gpurun_out/bench2_fused.json; tail -3 gpurun_out/bench2_fused.err; cut -c1-1200 gpurun_out/bench2_fused.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 --exchange nccl 2>gpurun_out/bench2_nccl.err | tail -1 > gpurun_out/bench2_nccl.json; tail -3 gpurun_out/bench2_nccl.err; cut -c1-600 gpurun_out/bench2_nccl.json
echo "== gpt-2 token trajectories: default kernels vs generic mat-vec (both parity-green)"
D=/tmp/ggml_b200_gpt2_v2; mkdir -p $D; export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref
[ -f $D/gpt2_q4_0.bin ] || { python scripts/make_gpt2_synth.py $D/gpt2_f16.bin >/dev/null; oracle/_ref/gpt-2-quantize $D/gpt2_f16.bin $D/gpt2_q4_0.bin 2 > /dev/null; }
ARGS="-m $D/gpt2_q4_0.bin -s 1234 -n 24 --ignore-eos --top_k 1 -p a_b_c -t 8"
oracle/_ref/gpt-2-backend $ARGS 2>&1 | grep "^a_b_c" | cut -c1-240
oracle/_ref/gpt-2-backend-b200 $ARGS -ngl 12 2>&1 | grep "^a_b_c" | cut -c1-240
GGML_B200_FORCE_GENERIC=1 oracle/_ref/gpt-2-backend-b200 $ARGS -ngl 12 2>&1 | grep "^a_b_c" | cut -c1-240
