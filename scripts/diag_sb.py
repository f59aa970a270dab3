"""diagnostic: v2 mat-vec in four modes (plain/graph x PDL on/off), each run under an external `timeout`"""
import os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g
mode = sys.argv[1]
t, M, N, K = g.Q4_K, 11008, 1, 4096
rb = g.row_size(t, K)
Ws = [torch.randint(0, 256, (M * rb,), dtype=torch.uint8, device="cuda") for _ in range(4)]
for w in Ws:
    v = w.view(-1, 144); v[:, 0:4] = torch.tensor([0, 0x10, 0, 0x10], dtype=torch.uint8, device="cuda")
X = torch.rand(K, device="cuda") * 2 - 1
Y = torch.empty((1, 1, N, M), device="cuda")
flags = g.MM_GEMV
print(mode, "plan", g.mul_mat_plan(t, M, N, K, flags), flush=True)
g.mul_mat(t, Ws[0], X, M, N, K, flags=flags, out=Y); torch.cuda.synchronize(); print("1 launch ok", float(Y.sum()), flush=True)
for i in range(8):
    g.mul_mat(t, Ws[i % 4], X, M, N, K, flags=flags, out=Y)
torch.cuda.synchronize(); print("8 launches ok", flush=True)
if mode.startswith("graph"):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(4):
            g.mul_mat(t, Ws[i], X, M, N, K, flags=flags, out=Y)
    torch.cuda.synchronize(); print("side stream ok", flush=True)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(4):
            g.mul_mat(t, Ws[i], X, M, N, K, flags=flags, out=Y)
    print("captured", flush=True)
    gr.replay(); torch.cuda.synchronize(); print("replay ok", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print("us per matvec", e0.elapsed_time(e1) * 1000 / 200, flush=True)
