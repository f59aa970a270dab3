"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, compiled from /root/reference).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Each fixture holds seeded inputs and the reference's own outputs:
  quant_<type>.npz   x (f32), blocks = ggml_quantize_chunk(type, x, imatrix=NULL; all-ones for the types that require one), deq = to_float(blocks),
                     rnd_blocks (arbitrary valid blocks), rnd_deq = to_float(rnd_blocks)
  act_q8.npz         x, q8_0 = CPU-backend from_float(Q8_0)(x), q8_K = from_float(Q8_K)(x), q8_1 = from_float(Q8_1)(x)
  mulmat_<type>.npz  W blocks, X, Y = MUL_MAT on the reference CPU backend (per case: M, N, K)
  mulmatid_<type>.npz  as above for MUL_MAT_ID
The GPU box has no /root/reference: these files are how its parity tests stay anchored to the reference.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
ref = O.Ref()

MM_CASES = [(16, 1, 256), (16, 8, 256), (48, 3, 512), (64, 2, 2048), (96, 16, 256)]
MM_CASES_32 = [(16, 1, 32), (16, 2, 96)]          # 32-element-block types only (odd block counts)


def synth(n, off):
    i = np.arange(n, dtype=np.float32)
    return (0.1 + 2.0 * np.cos(i + np.float32(off))).astype(np.float32)


B32 = (O.Q4_0, O.Q8_0, O.Q4_1, O.Q5_0, O.Q5_1, O.IQ4_NL)     # 32-element-block formats


def main(types):
    for t in types:
        name = O.TYPE_NAMES[t]
        rng = np.random.default_rng(1000 + t)
        x = np.concatenate([synth(2048, 0.0), rng.uniform(-1, 1, 2048).astype(np.float32)])
        blocks = ref.quantize(t, x, 2, 2048)
        rnd = O.random_blocks(t, 4096 // (32 if t in B32 else 256), rng)
        np.savez_compressed(OUT / f"quant_{name}.npz", x=x, blocks=blocks, deq=ref.dequantize(t, blocks, x.size),
                            rnd_blocks=rnd, rnd_deq=ref.dequantize(t, rnd, 4096))
        d = {}
        cases = MM_CASES + (MM_CASES_32 if t in B32 else [])
        for ci, (M, N, K) in enumerate(cases):
            r = np.random.default_rng(1234 + ci)
            W = ref.quantize(t, r.uniform(-1, 1, M * K).astype(np.float32), M, K)
            X = np.random.default_rng(5678 + ci).uniform(-1, 1, N * K).astype(np.float32)
            Y, _ = ref.mul_mat(t, W, X, M, N, K, threads=1)
            d[f"shape{ci}"] = np.array([M, N, K]); d[f"W{ci}"] = W; d[f"X{ci}"] = X; d[f"Y{ci}"] = Y[0, 0]
        d["ncases"] = np.array(len(cases))
        np.savez_compressed(OUT / f"mulmat_{name}.npz", **d)
        # MUL_MAT_ID: n_expert=4, n_used=2, b not broadcast, n_tok=5, M=32, K=256
        r = np.random.default_rng(4321 + t)
        ne, nu, ntok, M, K = 4, 2, 5, 32, 256
        W = ref.quantize(t, r.uniform(-1, 1, ne * M * K).astype(np.float32), ne * M, K)
        X = r.uniform(-1, 1, ntok * nu * K).astype(np.float32)
        ids = np.stack([r.permutation(ne)[:nu] for _ in range(ntok)]).astype(np.int32)
        Y, _ = ref.mul_mat_id(t, W, X, ids, M, K, ne, nu, nu, ntok, threads=1)
        np.savez_compressed(OUT / f"mulmatid_{name}.npz", W=W, X=X, ids=ids, Y=Y, cfg=np.array([ne, nu, nu, ntok, M, K]))
    rng = np.random.default_rng(77)
    x = np.concatenate([synth(1024, 1.0), rng.uniform(-3, 3, 1024).astype(np.float32), np.zeros(256, np.float32),
                        (np.round(rng.uniform(-127, 127, 768)) / 2).astype(np.float32)])
    np.savez_compressed(OUT / "act_q8.npz", x=x, q8_0=ref.cpu_from_float(O.Q8_0, x), q8_K=ref.cpu_from_float(O.Q8_K, x),
                        q8_1=ref.cpu_from_float(O.Q8_1, x),
                        q8_0_ref=ref.quantize_row_ref(O.Q8_0, x), q4_0_ref=ref.quantize_row_ref(O.Q4_0, x))
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    # python tests/golden/make_golden.py [hot|next|iq|all]   (default all; the fixtures are deterministic)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    main({"hot": O.HOT_TYPES, "next": O.NEXT_TYPES, "iq": O.IQ_TYPES, "all": O.HOT_TYPES + O.NEXT_TYPES + O.IQ_TYPES}[which])
