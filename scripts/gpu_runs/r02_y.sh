#!/bin/bash
# last GPU minutes of round 2: the mma small-batch kernel for all 12 formats (the 7 added last were host-emulated only), then smoke()
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "small_batch_mma" 2>&1 | tail -25 | tee gpurun_out/r02_y_mma.log
timeout 60 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_y_smoke.log
timeout 100 python scripts/gemv_sweep.py --types q5_0,q4_1,q5_1,iq4_nl,iq4_xs,q2_K,q3_K --shapes 4096x14336 --n 2,8 2>&1 | cut -c1-110 | tee gpurun_out/r02_y_sweep.log
