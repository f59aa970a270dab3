"""Run in a SUBPROCESS (own CUDA context, short timeout): the CTA-pair tcgen05 kernel (mmq_tc2.cu, cta_group::2) against the CPU oracle on
shapes that exercise: both CTAs with and without valid rows, split-K, every BN, ragged N / M tails, K-quant and 2-byte-aligned formats.
With --time also prints device times of BASELINE configs[2] (Q8_0 4096^2 x 512).  Exit code 0 = all checks passed."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ggml_b200 as g  # noqa: E402
from oracle import oracle as O  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main():
    assert os.environ.get("GGML_B200_TC_PAIR", "1") != "0"
    g.lib()
    orc = O.Oracle()
    rng = np.random.default_rng(22)
    cases = [(O.Q8_0, 256, 64, 256), (O.Q8_0, 512, 512, 1024), (O.Q4_K, 4096, 512, 4096), (O.Q4_0, 300, 100, 512), (O.Q5_K, 1000, 257, 2048),
             (O.Q6_K, 640, 130, 2048), (O.Q8_0, 200, 40, 768), (O.Q4_K, 11008, 96, 1024), (O.Q3_K, 520, 70, 2048), (O.Q5_0, 260, 33, 512),
             (O.Q8_0, 4096, 512, 4096)]
    for (t, M, N, K) in cases:
        assert g.mul_mat_plan(t, M, N, K) == g.MM_GEMM, (O.TYPE_NAMES[t], M, N, K)     # (rows must be 16-byte multiples: Q3_K / Q6_K need K % 2048 == 0)
        W = O.random_blocks(t, M * K // orc.blck_size(t), rng)
        X = rng.uniform(-1, 1, N * K).astype(np.float32)
        Wd, Xd = dev(W), dev(X)
        Y = g.mul_mat(t, Wd, Xd, M, N, K).cpu().numpy()[0, 0]
        assert np.isfinite(Y).all(), (O.TYPE_NAMES[t], M, N, K)
        rows = np.sort(rng.choice(M, min(M, 40), replace=False))
        rows[0], rows[-1] = 0, M - 1
        rb = orc.row_size(t, K)
        Wsub = np.concatenate([W[r * rb:(r + 1) * rb] for r in rows])
        want = orc.mul_mat(t, Wsub, X, len(rows), N, K, f64=True)
        err = O.nmse(Y[:, rows], want)
        assert err < 2e-5, (O.TYPE_NAMES[t], M, N, K, err)
        Y2 = g.mul_mat(t, Wd, Xd, M, N, K).cpu().numpy()[0, 0]
        assert np.array_equal(Y, Y2), ("not repeatable", O.TYPE_NAMES[t], M, N, K)
        print(f"ok {O.TYPE_NAMES[t]} {M}x{N}x{K} nmse {err:.2e}", flush=True)
    torch.cuda.synchronize()
    if "--time" in sys.argv:
        t, M, N, K = O.Q8_0, 4096, 512, 4096
        Ws = [dev(O.random_blocks(t, M * K // 32, rng)) for _ in range(8)]
        Xd = dev(rng.uniform(-1, 1, N * K).astype(np.float32))
        Ys = [torch.empty((1, 1, N, M), device="cuda") for _ in range(8)]
        F = g.MM_SRC0_STATIC | g.MM_SRC1_STATIC
        for _ in range(3):
            for i in range(8):
                g.mul_mat(t, Ws[i], Xd, M, N, K, out=Ys[i], flags=F)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            for i in range(8):
                g.mul_mat(t, Ws[i], Xd, M, N, K, out=Ys[i], flags=F)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 160
        print(f"time q8_0 4096x512x4096: {us:.2f} us per mul_mat = {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s (env BN={os.environ.get('GGML_B200_TC2_BN')}, SPLITK={os.environ.get('GGML_B200_TC_SPLITK')}, STAGES={os.environ.get('GGML_B200_TC2_STAGES')})", flush=True)
    print("tc2 pair kernel OK")


if __name__ == "__main__":
    main()
