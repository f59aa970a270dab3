#!/bin/bash
D=/tmp/ggml_b200_gpt2_v2; mkdir -p $D
export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref
[ -f $D/gpt2_q4_0.bin ] || { python scripts/make_gpt2_synth.py $D/gpt2_f16.bin; oracle/_ref/gpt-2-quantize $D/gpt2_f16.bin $D/gpt2_q4_0.bin 2 > /dev/null; }
ARGS="-m $D/gpt2_q4_0.bin -s 1234 -n 24 -t 8 --ignore-eos --top_k 1"
echo "cpu  : $(oracle/_ref/gpt-2-backend $ARGS -p 'a b c' 2>/dev/null | grep '^a b c' | cut -c1-330)"
echo "b200 : $(oracle/_ref/gpt-2-backend-b200 $ARGS -p 'a b c' -ngl 12 2>/dev/null | grep '^a b c' | cut -c1-330)"
echo "sched: $(oracle/_ref/gpt-2-sched-b200 $ARGS -p 'a b c' -ngl 99 2>/dev/null | grep '^a b c' | cut -c1-330)"
echo "gen  : $(GGML_B200_FORCE_GENERIC=1 oracle/_ref/gpt-2-backend-b200 $ARGS -p 'a b c' -ngl 12 2>/dev/null | grep '^a b c' | cut -c1-330)"
