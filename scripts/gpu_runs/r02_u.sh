#!/bin/bash
# checkpoint: the whole GPU suite, smoke, GEMM stress, bench-pattern parity of the GEMM defaults
mkdir -p gpurun_out
SECONDS=0
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -25 > gpurun_out/pytest_r02_u.log; tail -12 gpurun_out/pytest_r02_u.log; echo "suite: ${SECONDS}s"
timeout 120 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python tests/gpu_tc2_stress.py 10 2>&1 | grep -E "BAD|CLEAN" | cut -c1-200 | tail -5
for sh in "q6_K 4096 512 4096" "q6_K 11008 64 4096" "q4_K 4096 32 4096" "q8_0 4096 200 14336"; do timeout 200 python scripts/gemm_bench_parity.py $sh --serial 2>&1 | tail -3 | cut -c1-260; done
