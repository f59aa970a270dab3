#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=oracle/_ref/native:oracle/_ref
echo "=== small op via test-backend-ops"; GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so timeout 60 oracle/_ref/test-backend-ops test -o ADD -b B2000 2>&1 | tail -12
GGML_BACKEND_PATH=$PWD/ggml_b200/libggml-b200.so timeout 60 oracle/_ref/test-backend-ops test -o GET_ROWS -b B2000 2>&1 | grep -v "not supported" | tail -8
unset LD_LIBRARY_PATH
for t in 2 8 12 13 14; do echo "=== shapes type $t"; timeout 90 python scripts/diag_shapes.py $t 2>&1 | tail -30; echo "rc=$?"; done
for m in plain graph; do for pdl in 0 1; do
  echo "=== mode=$m NO_PDL=$pdl"; GGML_B200_NO_PDL=$pdl timeout 60 python scripts/diag_sb.py $m 2>&1 | tail -8; echo "rc=$?"
done; done
echo "=== bench"; timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 | cut -c1-600
