"""Developer aid (ncu): a few launches of the small-batch mma kernel.  usage: python scripts/mma_one.py TYPE M N K"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g  # noqa: E402

t = {v: k for k, v in g.TYPE_NAMES.items()}[sys.argv[1]]
M, N, K = (int(v) for v in sys.argv[2:5])
rb = g.row_size(t, K)
Ws = [torch.randint(0, 256, (M * rb,), dtype=torch.uint8, device="cuda") for _ in range(4)]
X = torch.rand(N * K, device="cuda") * 2 - 1
for i in range(8):
    Y = g.mul_mat(t, Ws[i % 4], X, M, N, K, flags=g.MM_GEMV | g.MM_GEMV_MMA | g.MM_SRC0_STATIC)
torch.cuda.synchronize()
print("ok", float(Y.abs().max()))
