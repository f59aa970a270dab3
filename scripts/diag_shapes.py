"""diagnostic: run the v2 mat-vec on the test shapes one by one, printing before each launch (run under `timeout`)"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g
from oracle import oracle as O
orc = O.Oracle()
types = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 8, 12, 13, 14]
for t in types:
    for (M, K) in [(1000, 4096), (257, 1024), (64, 10752), (24, 256), (4096, 768 if t in (2, 8) else 2048), (3000, 2048), (136, 8192)]:
        plan = g.mul_mat_plan(t, M, 1, K, g.MM_GEMV)
        print("case", t, M, K, "plan", plan, flush=True)
        if plan != g.MM_GEMV:
            continue
        rng = np.random.default_rng(M + K)
        W = O.random_blocks(t, M * K // orc.blck_size(t), rng)
        X = rng.uniform(-1, 1, K).astype(np.float32)
        Wd, Xd = torch.from_numpy(W).cuda(), torch.from_numpy(X).cuda()
        torch.cuda.synchronize(); print("  launching", flush=True)
        Y = g.mul_mat(t, Wd, Xd, M, 1, K, flags=g.MM_GEMV)
        torch.cuda.synchronize(); print("  done", flush=True)
        err = O.nmse(Y.cpu().numpy()[0, 0], orc.mul_mat(t, W, X, M, 1, K))
        print("  nmse", err, flush=True)
