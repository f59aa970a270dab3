#!/bin/bash
# pair GEMM v4 (bulk-store epilogue) + gpt-2 host-side profile
D=/tmp/ggml_b200_gpt2_v2; mkdir -p $D gpurun_out
for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_TMA_EPI=0" "GGML_B200_TC_SPLITK=3" "GGML_B200_TC2_RAWMODE=1"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 6 2>&1 | grep -v "^ok" | tail -8 | cut -c1-330
done
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -1
GGML_B200_MMID_GROUPED=1 timeout 300 python tests/gpu_mmid_grouped_check.py 2>&1 | tail -1
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096"; do
  for cfg in "X=0" "GGML_B200_TC2_TMA_EPI=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=1" "GGML_B200_TC_PAIR=0"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
echo "== accounts"
for cfg in "X=0" "GGML_B200_TC2_BN=128"; do echo "-- $cfg"; env $cfg GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19; done
echo "== gpt-2 host profile"
[ -f $D/gpt2_q4_0.bin ] || { python scripts/make_gpt2_synth.py $D/gpt2_f16.bin > /dev/null; LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref oracle/_ref/gpt-2-quantize $D/gpt2_f16.bin $D/gpt2_q4_0.bin 2 > /dev/null 2>&1; rm -f $D/gpt2_f16.bin; }
for cfg in "X=0" "GGML_B200_GRAPH_MAX_UPDATES=100000" "GGML_B200_DISABLE_GRAPHS=1" "GGML_B200_DISABLE_FUSION=1"; do
  echo "-- $cfg"
  env $cfg GGML_B200_PROFILE=1 LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref timeout 200 oracle/_ref/gpt-2-backend-b200 -m $D/gpt2_q4_0.bin -s 1234 -n 128 -t 8 --ignore-eos --top_k 1 -p "a b c" -ngl 12 2>&1 | grep -E "profile|predict time|sample time|total time|load time"
done
