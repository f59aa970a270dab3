"""GPU: the SURVEY §8f-2 formats (Q4_1, Q5_0, Q5_1, Q2_K, Q3_K, IQ4_NL, IQ4_XS) on the superblock and generic mat-vec / MUL_MAT_ID / dequantize kernels.

Status (round 1): the oracle is pinned against the reference and the kernels' decode logic is verified on the host
(tests/test_hostemu_kernel_logic.py), but these kernels were added after the round's GPU budget was spent, so this module has not yet
run on a B200: it is a non-strict xfail (XPASS = the kernels are correct as written), executed in a subprocess so that a fault
in a not-yet-validated kernel cannot poison the CUDA context of the rest of the suite.  The backend (`supports_op`) does not
advertise these formats until this test has passed on hardware."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.xfail(strict=False, reason="kernels for the §8f-2 formats not yet validated on a B200 (host-emulated only)")
def test_next_formats_parity_subprocess():
    p = subprocess.run([sys.executable, str(ROOT / "tests" / "gpu_next_formats_check.py")], capture_output=True, text=True, timeout=150)
    print(p.stdout[-2000:])
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]


@pytest.mark.xfail(strict=False, reason="opt-in tcgen05 path for Q6_K (GGML_B200_TC_Q6K=1): decoder host-verified, kernel path not yet validated on a B200")
def test_q6k_gemm_opt_in_subprocess():
    import os
    env = dict(os.environ, GGML_B200_TC_Q6K="1")
    p = subprocess.run([sys.executable, str(ROOT / "tests" / "gpu_q6k_gemm_check.py")], capture_output=True, text=True, timeout=90, env=env)
    print(p.stdout[-2000:])
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
