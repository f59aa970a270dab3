#!/bin/bash
# round 2, third GPU bundle: gpt-2 per-node parity (sync / free, also the reference CUDA backend for context), new gpt-2 tests, GEMM kernel timings + ncu
mkdir -p gpurun_out
D=/tmp/ggml_b200_gpt2_v2; mkdir -p $D
[ -f $D/gpt2_q4_0.bin ] || { python scripts/make_gpt2_synth.py $D/gpt2_f16.bin > /dev/null; LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref oracle/_ref/gpt-2-quantize $D/gpt2_f16.bin $D/gpt2_q4_0.bin 2 > /dev/null 2>&1; rm -f $D/gpt2_f16.bin; }
echo "== gpt-2 graph node by node"
for mode in sync free; do
  LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so timeout 300 oracle/_ref/gpt2-compare $D/gpt2_q4_0.bin B2000 5 $mode 2>/dev/null > gpurun_out/gpt2_compare_$mode.log; grep summary gpurun_out/gpt2_compare_$mode.log
done
sort -g -k 9 gpurun_out/gpt2_compare_sync.log 2>/dev/null | grep "^node" | tail -5
LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref:/usr/local/cuda/lib64 GGML_BACKEND_PATH=$PWD/oracle/_ref/cuda/libggml-cuda.so timeout 300 oracle/_ref/gpt2-compare $D/gpt2_q4_0.bin CUDA0 5 free 2>/dev/null > gpurun_out/gpt2_compare_refcuda.log; echo "reference CUDA backend vs ggml-cpu:"; grep summary gpurun_out/gpt2_compare_refcuda.log
echo "== gpt-2 tests"
timeout 900 python -m pytest tests/test_gpu_gpt2.py -q -m gpu --timeout 600 -s 2>&1 | grep -E "gpt-2|passed|failed|Error|assert" | tail -20
echo "== GEMM timings (graph replay)"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096"; do
  GGML_B200_TC_PAIR=0 timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1
  timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1
done
for cfg in "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=1" "GGML_B200_TC2_BN=128 GGML_B200_TC_SPLITK=2" "GGML_B200_NO_PDL=1" "GGML_B200_TC2_STAGES=3"; do env $cfg timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1; done
echo "== ncu launch list (q8_0 4096x512x4096, pair kernel then one-CTA kernel)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_gemm_launches_pair.csv python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > /dev/null 2>&1
GGML_B200_TC_PAIR=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_gemm_launches_onecta.csv python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > /dev/null 2>&1
python - <<'PY'
import csv
for name in ("pair", "onecta"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r02_gemm_launches_{name}.csv")) if len(r) > 10 and r[0].isdigit()]
    print(name, [(r[4].split("(")[0][-40:], r[-1]) for r in rows[-8:]])
PY
echo "== ncu full: pair kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc2 -s 12 -c 1 -o gpurun_out/r02_gemm_pair -f python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
GGML_B200_TC_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc_kernel -s 12 -c 1 -o gpurun_out/r02_gemm_onecta -f python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > gpurun_out/ncu_onecta.log 2>&1; tail -2 gpurun_out/ncu_onecta.log
echo "== mat-vec: dependent variants + long rows"
for cfg in "X=0" "GGML_B200_SB_TWOROW=0" "GGML_B200_SB_TWOROW=1 GGML_B200_SB_RESIDENT=1" "GGML_B200_SB_TWOROW=1 GGML_B200_SB_RESIDENT=1 GGML_B200_SB_STAGES=3"; do echo "-- $cfg"; env $cfg timeout 120 python scripts/gemv_sweep.py --types q4_K --shapes 11008x4096,4096x4096,32000x4096 2>&1 | cut -c1-110 | tail -3; done
timeout 200 python scripts/gemv_sweep.py --types q4_K,q8_0 --shapes 4096x14336 --n 1,4,8 2>&1 | cut -c1-120 | tail -6
