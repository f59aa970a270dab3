#!/bin/bash
# SURVEY §8d "reference kernel to beat": the reference's own CUDA backend (oracle/_ref/cuda/libggml-cuda.so, src/ggml-cuda compiled unmodified for
# sm_100 by oracle/Makefile refcuda) timed by the reference's own harness on this B200, next to this repo's backend through the same harness.
mkdir -p gpurun_out
export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref:/usr/local/cuda/lib64
for be in cuda b200; do
  if [ $be = cuda ]; then export GGML_BACKEND_PATH=$PWD/oracle/_ref/cuda/libggml-cuda.so; dev=CUDA0; else export GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so; dev=B2000; fi
  timeout 600 oracle/_ref/test-backend-ops perf -o MUL_MAT -b $dev > gpurun_out/tbo_perf_$be.log 2>&1
  echo "== $be ($dev) rc=$?"; grep -E "type_a=(q4_0|q8_0|q4_K|q5_K|q6_K),type_b=f32,m=4096,n=(1|8|512),k=14336" gpurun_out/tbo_perf_$be.log | sed 's/  */ /g' | cut -c1-200
done
