"""prints the %globaltimer timeline of consecutive mat-vec launches inside one CUDA graph (GGML_B200_SB_DEBUG=1)"""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ggml_b200 as g
t, M, N, K = g.Q4_K, 11008, 1, 4096
rb = g.row_size(t, K)
Ws = [torch.randint(0, 256, (M * rb,), dtype=torch.uint8, device="cuda") for _ in range(13)]
X = torch.rand(K, device="cuda"); Y = torch.empty((1, 1, N, M), device="cuda")
st = os.environ.get("STATIC", "1")
fl = g.MM_GEMV | (g.MM_SRC0_STATIC if st in ("1", "2") else 0) | (g.MM_SRC1_STATIC if st == "2" else 0)
Ys = [torch.empty((1, 1, N, M), device="cuda") for _ in range(13)]
for i in range(13):
    g.mul_mat(t, Ws[i], X, M, N, K, flags=fl, out=(Ys[i] if st == "2" else Y))
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for i in range(13):
        g.mul_mat(t, Ws[i], X, M, N, K, flags=fl, out=(Ys[i] if st == "2" else Y))
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 256)()
g.lib().ggml_b200_debug_trace.argtypes = [C.c_void_p]
assert g.lib().ggml_b200_debug_trace(buf) == 0
a = np.array(buf, dtype=np.uint64).reshape(32, 8).astype(np.int64)
a = a[a[:, 0] > 0]
a = a[np.argsort(a[:, 0])][-13:]
t0 = a[0, 0]
print("launch  entry  tma0  prod_wait  cons_wait  stage0  done  all_done  last_cons_release   (us since first entry; NO_PDL=%s STATIC=%s)" % (os.environ.get("GGML_B200_NO_PDL", "0"), os.environ.get("STATIC", "1")))
for i, r in enumerate(a):
    print(i, " ".join(f"{(v - t0) / 1000:8.2f}" if v > 0 else "     n/a" for v in r[:8]))

if len(a) > 4:
    b = a[2:]
    period = np.diff(b[:, 6]).mean() / 1000
    bound = (b[1:, 7] - b[:-1, 6]).mean() / 1000
    quant = (b[:, 4] - b[:, 3]).mean() / 1000
    cons = (b[:, 6] - b[:, 4]).mean() / 1000
    print(f"SUMMARY period {period:.2f} us  boundary {bound:.2f}  quantize(cta0) {quant:.2f}  consume {cons:.2f}   tun " + " ".join(f"{k[13:]}={v}" for k, v in os.environ.items() if k.startswith("GGML_B200_SB_")))
