"""Run in a SUBPROCESS: repeatability / exactness stress of the CTA-pair tcgen05 kernel.  Every case is computed REPS times and every
element of every run is compared (a) with the first run (bitwise: the kernel has no run-to-run freedom) and (b) with the exact product of
the dequantized weights in float64 (computed on the device with torch; tolerance = the fp16-operand rounding).  A stale operand tile
(a missing fence / a barrier race) shows up as a handful of elements far outside that tolerance; their positions are printed.
usage: python tests/gpu_tc2_stress.py [REPS]   (env GGML_B200_TC2_* select the kernel variant).  Exit code 0 = clean."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import ggml_b200 as g  # noqa: E402
from oracle import oracle as O  # noqa: E402


def raw_mismatches():
    """dbg & 32 with GGML_B200_TC2_TRACE=1: number of raw units a dequantizer thread found different from global memory (cumulative)."""
    import ctypes
    if os.environ.get("GGML_B200_TC2_TRACE", "0") == "0":
        return 0
    buf = (ctypes.c_uint64 * (4096 * 8))()
    fn = g.lib().ggml_b200_debug_gemm_trace
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    return int(buf[4096 * 8 - 1]) if fn(buf, 4096) else 0


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10
    orc = O.Oracle()
    rng = np.random.default_rng(5)
    cases = [(O.Q4_K, 4096, 512, 4096), (O.Q6_K, 640, 130, 2048), (O.Q8_0, 4096, 512, 4096), (O.Q8_0, 512, 512, 1024), (O.Q4_0, 300, 100, 512),
             (O.Q5_K, 1000, 257, 2048), (O.Q4_K, 11008, 96, 1024), (O.Q8_0, 4096, 64, 4096), (O.Q4_K, 4096, 128, 4096), (O.Q8_0, 8192, 48, 4096)]
    if "--big" in sys.argv:
        cases = [c for c in cases if c[1] >= 4096]
    bad = 0
    for (t, M, N, K) in cases:
        assert g.mul_mat_plan(t, M, N, K) == g.MM_GEMM
        W = torch.from_numpy(O.random_blocks(t, M * K // orc.blck_size(t), rng)).cuda()
        X = torch.from_numpy(rng.uniform(-1, 1, N * K).astype(np.float32)).cuda()
        Wf = g.dequantize(t, W, M * K).view(M, K).double()
        exact = (X.view(N, K).double() @ Wf.T)                           # [N, M]
        scale = exact.abs().mean().item()
        tol = 6e-3 * scale                                               # fp16 operands: per-element error ~ 1e-3 x typical magnitude
        first = None
        n_diff_runs, n_wrong_runs = 0, 0
        for r in range(reps):
            Y = g.mul_mat(t, W, X, M, N, K)[0, 0]
            err = (Y.double() - exact).abs()
            wrong = err > tol
            if wrong.any():
                n_wrong_runs += 1
                idx = wrong.nonzero()
                cols, rows = idx[:, 0].cpu().numpy(), idx[:, 1].cpu().numpy()
                print(f"  run {r}: {len(rows)} elements off by up to {err.max().item() / scale:.3f} x typical; rows {rows.min()}..{rows.max()} "
                      f"(row % 256 < 128: {int((rows % 256 < 128).sum())}, >= 128: {int((rows % 256 >= 128).sum())}), cols {cols.min()}..{cols.max()}, "
                      f"distinct rows {len(set(rows.tolist()))}, distinct cols {len(set(cols.tolist()))}, rows % 128: {sorted(set((rows % 128).tolist()))[:16]}", flush=True)
            if first is None:
                first = Y.clone()
            elif not torch.equal(first, Y):
                n_diff_runs += 1
                d = (first != Y).nonzero()
                cols, rows = d[:, 0].cpu().numpy(), d[:, 1].cpu().numpy()
                print(f"  run {r} differs from run 0 in {len(rows)} elements: rows {rows.min()}..{rows.max()} (row % 256 >= 128: {int((rows % 256 >= 128).sum())}), "
                      f"cols {cols.min()}..{cols.max()}, max |diff| {((first - Y).abs().max().item()) / scale:.2e} x typical", flush=True)
        raw_bad = raw_mismatches()
        if raw_bad:
            print(f"  raw-ring cross-check: {raw_bad} units differed from global memory so far", flush=True)
        status = "ok" if n_diff_runs == 0 and n_wrong_runs == 0 else "BAD"
        bad += status == "BAD"
        print(f"{status} {O.TYPE_NAMES[t]} {M}x{N}x{K}: {reps} runs, {n_diff_runs} differ from the first, {n_wrong_runs} outside tolerance "
              f"(max err {((Y.double() - exact).abs().max().item()) / scale:.2e} x typical)", flush=True)
    tun = {k: v for k, v in os.environ.items() if k.startswith("GGML_B200_")}
    print("tc2 stress", "CLEAN" if bad == 0 else f"{bad} BAD CASES", tun, flush=True)
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
