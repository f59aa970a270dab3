"""GPU, >= 2 devices: the row-sharded mat-vec with the exchange fused into the kernel (BASELINE.json configs[4]), one
process per GPU under torchrun, every rank checked against the CPU oracle (tests/gpu_multi_check.py).  Skips on a single-GPU box."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_fused_gather_all_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29731",
           str(ROOT / "tests" / "gpu_multi_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(p.stdout[-3000:])
    assert p.returncode == 0, (p.stdout + p.stderr)[-4000:]
    assert p.stdout.count("fused NVLink gather OK") == n
