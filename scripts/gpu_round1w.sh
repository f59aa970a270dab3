#!/bin/bash
for cfg in "q4_0 96 16 256" "q8_0 256 32 512" "q4_K 4096 512 4096" "q8_0 4096 512 4096"; do
  echo "== $cfg"; timeout 120 python scripts/gemm_one.py $cfg 2>&1 | tail -1 | cut -c1-150
done
echo "== sanitizer q8_0 256 32 512"
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python scripts/gemm_one.py q8_0 256 32 512 2>&1 | grep -v "^$" | grep "=========" | head -30
