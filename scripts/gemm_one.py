import sys, numpy as np, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g
t = {v: k for k, v in g.TYPE_NAMES.items()}[sys.argv[1] if len(sys.argv) > 1 else "q4_K"]
M, N, K = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (256, 32, 512)))
rb = g.row_size(t, K)
W = torch.zeros(M * rb, dtype=torch.uint8, device="cuda")
X = torch.rand(N * K, device="cuda")
Y = g.mul_mat(t, W, X, M, N, K, flags=g.MM_GEMM)
torch.cuda.synchronize()
print("ok", float(Y.abs().max()))
