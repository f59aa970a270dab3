#!/bin/bash
# round 2, multi-GPU bundle (run with gpurun --gpus N): fused-gather parity on every rank, split buffers across devices, weak-scaled bench (fused vs NCCL)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== multi-GPU parity (torchrun, $N ranks)"
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --timeout 600 -s 2>&1 | grep -E "OK|passed|failed|Error|assert|skipped" | tail -12
echo "== split buffers over $N devices (single process, the reference's own entry point)"
timeout 600 python -m pytest tests/test_gpu_backend_plugin.py -q -m gpu -x --timeout 600 -k "split_buffer" -s 2>&1 | tail -6
echo "== bench --gpus $N (fused exchange), then the NCCL baseline"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N > gpurun_out/bench_r02_n${N}_fused.json 2> gpurun_out/bench_r02_n${N}_fused.err; tail -c 1800 gpurun_out/bench_r02_n${N}_fused.json; tail -3 gpurun_out/bench_r02_n${N}_fused.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --exchange nccl > gpurun_out/bench_r02_n${N}_nccl.json 2> gpurun_out/bench_r02_n${N}_nccl.err; tail -c 600 gpurun_out/bench_r02_n${N}_nccl.json; tail -3 gpurun_out/bench_r02_n${N}_nccl.err
