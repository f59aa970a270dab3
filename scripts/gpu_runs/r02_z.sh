#!/bin/bash
# the very last GPU minutes of round 2: the two test files that reach the mma small-batch kernel through the DEFAULT dispatch, then n = 3..5 timings
mkdir -p gpurun_out
timeout 60 python scripts/gemv_sweep.py --types q5_0,q4_1,q5_1,iq4_nl,iq4_xs,q2_K,q3_K --shapes 4096x14336 --n 1,3,4,5 2>&1 | cut -c1-110 > gpurun_out/r02_z_sweep.log
(timeout 60 python -m pytest tests/test_gpu_next_formats.py -q -m gpu 2>&1 | tail -8;
 timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not small_batch_mma" 2>&1 | tail -8) | tee gpurun_out/r02_z_tests.log
