#!/bin/bash
# GPT-2 117M (synthetic weights, Q4_0) decode: reference ggml-cpu vs the B200 backend, unmodified example programs.
set -e
D=/tmp/ggml_b200_gpt2_v2; mkdir -p $D gpurun_out
export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref
[ -f $D/gpt2_q4_0.bin ] || { python scripts/make_gpt2_synth.py $D/gpt2_f16.bin; oracle/_ref/gpt-2-quantize $D/gpt2_f16.bin $D/gpt2_q4_0.bin 2 > /dev/null; }
ARGS="-m $D/gpt2_q4_0.bin -s 1234 -n 128 --ignore-eos --top_k 1 -p a_b_c"
for t in 4 8 16; do echo "== cpu gpt-2-backend -t $t"; oracle/_ref/gpt-2-backend $ARGS -t $t 2>&1 | grep -E "predict time|^a_b_c" | cut -c1-200; done
echo "== b200 gpt-2-backend-b200"; oracle/_ref/gpt-2-backend-b200 $ARGS -t 8 -ngl 12 2>&1 | grep -E "predict time|^a_b_c|CUDA" | cut -c1-200
echo "== b200 gpt-2-sched-b200 -ngl 99"; oracle/_ref/gpt-2-sched-b200 $ARGS -t 8 -ngl 99 2>&1 | grep -E "predict time|^a_b_c" | cut -c1-200
