#!/bin/bash
# round 2, fourth GPU bundle: reworked CTA-pair GEMM (cheap fences, decoupled producers, deeper raw ring, inv_scale staging): parity, tuning sweep, ncu
mkdir -p gpurun_out
echo "== pair kernel parity"
timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -12
GGML_B200_TC2_BN=128 timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -2
GGML_B200_TC_SPLITK=3 GGML_B200_TC2_RAW=2 timeout 300 python tests/gpu_tc2_check.py 2>&1 | tail -2
GGML_B200_MMID_GROUPED=1 timeout 300 python tests/gpu_mmid_grouped_check.py 2>&1 | tail -4
echo "== sweep"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096" "q6_K 4096 512 4096"; do
  GGML_B200_TC_PAIR=0 timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1
  for cfg in "X=0" "GGML_B200_TC2_BN=128" "GGML_B200_TC_SPLITK=1" "GGML_B200_TC2_BN=128 GGML_B200_TC_SPLITK=2"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
for cfg in "GGML_B200_TC2_BN=128 GGML_B200_TC2_STAGES=3" "GGML_B200_TC2_BN=128 GGML_B200_TC2_STAGES=6" "GGML_B200_TC2_BN=128 GGML_B200_TC2_RAW=2" "GGML_B200_TC2_BN=128 GGML_B200_TC2_RAW=4" "GGML_B200_TC2_STAGES=3" "GGML_B200_TC2_RAW=2" "GGML_B200_TC2_BN=128 GGML_B200_NO_PDL=1"; do env $cfg timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 2>&1 | tail -1; done
echo "== ncu launch list + full capture (q8_0 4096x512x4096, BN=128 no split-K; then default)"
GGML_B200_TC2_BN=128 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_gemm_launches_bn128.csv python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_gemm_launches_bn256.csv python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > /dev/null 2>&1
python - <<'PY'
import csv
for name in ("bn128", "bn256"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r02d_gemm_launches_{name}.csv")) if len(r) > 10 and r[0].isdigit()]
    print(name, [(r[4].split("(")[0][-40:], r[-1]) for r in rows[-8:]])
PY
GGML_B200_TC2_BN=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc2 -s 12 -c 1 -o gpurun_out/r02d_gemm_pair_bn128 -f python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > gpurun_out/ncu_pair_d.log 2>&1; tail -2 gpurun_out/ncu_pair_d.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmq_tc2 -s 12 -c 1 -o gpurun_out/r02d_gemm_pair_bn256 -f python scripts/gemm_prof.py q8_0 4096 512 4096 --ncu > gpurun_out/ncu_pair_d2.log 2>&1; tail -2 gpurun_out/ncu_pair_d2.log
