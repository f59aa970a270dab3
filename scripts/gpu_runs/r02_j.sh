#!/bin/bash
# pair GEMM v5 (Q8_0 8-block units, early weight requests, bulk-store epilogue): stress, parity, timings, accounts, ncu
mkdir -p gpurun_out
for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC2_BN=128 GGML_B200_TC2_STAGES=8" "GGML_B200_TC_SPLITK=3" "GGML_B200_TC2_BN=64" "GGML_B200_TC2_BN=64 GGML_B200_TC2_TMA_EPI=0"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 20 --big 2>&1 | grep -v "^ok" | tail -8 | cut -c1-330
done
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -1
GGML_B200_MMID_GROUPED=1 timeout 300 python tests/gpu_mmid_grouped_check.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_formats.py -q -m gpu -x -k "gemm or next" 2>&1 | tail -3
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q5_K 4096 512 4096" "q6_K 4096 512 4096" "q4_0 4096 512 4096" "q8_0 32000 512 4096" "q4_K 32000 512 4096" "q4_K 11008 512 4096" "q8_0 4096 128 4096" "q4_K 4096 2048 4096"; do
  for cfg in "X=0" "GGML_B200_TC2_TMA_EPI=0"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
for cfg in "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=1" "GGML_B200_TC_SPLITK=3" "GGML_B200_TC_SPLITK=4" "GGML_B200_TC_PAIR=0"; do env $cfg timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1; done
echo "== accounts"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096"; do GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py $sh --trace 2>&1 | tail -19; done
echo "== ncu"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_gemm_launches.csv python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02j_gemm_launches.csv")) if len(r) > 10 and r[0].isdigit()]
print([(r[4].split("(")[0][-40:], r[-1]) for r in rows[-8:]])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc2 -s 12 -c 1 -o gpurun_out/r02j_gemm_pair -f python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > gpurun_out/ncu_pair_j.log 2>&1; tail -2 gpurun_out/ncu_pair_j.log
