#!/bin/bash
# independent-CTA solo variant (no multicast) vs pair mode; dense fp16 path in pair mode: stress + checks + timings
cat > /tmp/dense_stress.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import ggml_b200 as g
from oracle import oracle as O
orc = O.Oracle(); rng = np.random.default_rng(3); bad = 0
for (t, M, N, K) in [(O.IQ2_XXS, 4096, 512, 4096), (O.IQ3_S, 1000, 257, 2048), (O.TQ2_0, 4096, 64, 4096), (O.IQ1_S, 640, 130, 2048)]:
    assert g.mul_mat_plan(t, M, N, K) == g.MM_GEMM
    W = torch.from_numpy(O.random_blocks(t, M * K // orc.blck_size(t), rng)).cuda()
    X = torch.from_numpy(rng.uniform(-1, 1, N * K).astype(np.float32)).cuda()
    exact = X.view(N, K).double() @ g.dequantize(t, W, M * K).view(M, K).double().T
    scale = exact.abs().mean().item(); first = None; nb = 0
    for r in range(8):
        Y = g.mul_mat(t, W, X, M, N, K)[0, 0]
        if ((Y.double() - exact).abs() > 6e-3 * scale).any() or (first is not None and not torch.equal(first, Y)): nb += 1
        first = Y.clone() if first is None else first
    print(("ok" if nb == 0 else "BAD"), O.TYPE_NAMES[t], M, N, K, "bad runs", nb, "max err", ((Y.double() - exact).abs().max().item()) / scale, flush=True); bad += nb > 0
print("dense stress", "CLEAN" if bad == 0 else "BAD")
PY
for cfg in "GGML_B200_TC2_SOLO=2" "GGML_B200_TC2_SOLO=2 GGML_B200_TC2_BN=128" "X=0"; do
  echo "-- $cfg"
  env $cfg timeout 300 python tests/gpu_tc2_stress.py 12 --big 2>&1 | grep -E "BAD|CLEAN" | cut -c1-200
  env $cfg timeout 300 python /tmp/dense_stress.py 2>&1 | tail -5
done
timeout 900 python -m pytest tests/test_gpu_next_formats.py -q -m gpu -x 2>&1 | tail -2
echo "== timings"
for sh in "q8_0 4096 512 4096" "q4_K 4096 512 4096" "q6_K 4096 512 4096" "q8_0 32000 512 4096" "q4_K 11008 512 4096" "q8_0 4096 128 4096" "iq2_xxs 4096 512 14336" "iq3_s 4096 512 4096"; do
  for cfg in "GGML_B200_TC2_SOLO=2" "X=0"; do env $cfg timeout 120 python scripts/gemm_prof.py $sh 2>&1 | tail -1; done
done
GGML_B200_TC2_SOLO=2 GGML_B200_TC2_TRACE=1 timeout 120 python scripts/gemm_prof.py q8_0 4096 512 4096 --trace 2>&1 | tail -19
