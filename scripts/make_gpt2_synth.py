"""Writes a synthetic GPT-2 117M model in the legacy `ggml` format read by examples/gpt-2 (SURVEY.md Appendix B;
reference loader examples/gpt-2/main-backend.cpp:112-131, 146-168, 355-430): random N(0, 0.02^2) f16 weights,
single-character ASCII vocabulary so that `-p "a b c"` tokenizes, explicit model/lm_head tensor.
Quantize afterwards with the reference's own tool:  oracle/_ref/gpt-2-quantize f16.bin q4_0.bin 2
Usage: python scripts/make_gpt2_synth.py OUT.bin [--layers 12]"""
import argparse
import struct

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    n_vocab, n_ctx, n_embd, n_head, n_layer, ftype = 50257, 1024, 768, 12, a.layers, 1
    rng = np.random.default_rng(a.seed)
    with open(a.out, "wb") as f:
        f.write(struct.pack("<I", 0x67676d6c))
        f.write(struct.pack("<6i", n_vocab, n_ctx, n_embd, n_head, n_layer, ftype))
        f.write(struct.pack("<i", n_vocab))
        for i in range(n_vocab):
            tok = chr(i).encode() if 32 <= i < 127 else f"<{i}>".encode()
            f.write(struct.pack("<I", len(tok)) + tok)

        def tensor(name, shape, dtype):
            # shape in ggml order (ne[0] first); data laid out with ne[0] contiguous
            n = int(np.prod(shape))
            if dtype == "f16":
                data = (rng.standard_normal(n).astype(np.float32) * 0.02).astype(np.float16)
                tt = 1
            else:
                base = 1.0 if name.endswith("/g") else 0.0
                data = (base + rng.standard_normal(n).astype(np.float32) * 0.02).astype(np.float32)
                tt = 0
            nb = name.encode()
            f.write(struct.pack("<3i", len(shape), len(nb), tt))
            for d in shape:
                f.write(struct.pack("<i", d))
            f.write(nb)
            f.write(data.tobytes())

        tensor("model/ln_f/g", [n_embd], "f32"); tensor("model/ln_f/b", [n_embd], "f32")
        # lm_head must precede wte: the loaders alias lm_head to wte when wte is seen first (main-sched.cpp:504-513)
        tensor("model/lm_head", [n_embd, n_vocab], "f16")
        tensor("model/wte", [n_embd, n_vocab], "f16")
        tensor("model/wpe", [n_embd, n_ctx], "f32")
        for i in range(n_layer):
            p = f"model/h{i}/"
            tensor(p + "ln_1/g", [n_embd], "f32"); tensor(p + "ln_1/b", [n_embd], "f32")
            tensor(p + "ln_2/g", [n_embd], "f32"); tensor(p + "ln_2/b", [n_embd], "f32")
            tensor(p + "attn/c_attn/w", [n_embd, 3 * n_embd], "f16"); tensor(p + "attn/c_attn/b", [3 * n_embd], "f32")
            tensor(p + "attn/c_proj/w", [n_embd, n_embd], "f16"); tensor(p + "attn/c_proj/b", [n_embd], "f32")
            tensor(p + "mlp/c_fc/w", [n_embd, 4 * n_embd], "f16"); tensor(p + "mlp/c_fc/b", [4 * n_embd], "f32")
            tensor(p + "mlp/c_proj/w", [4 * n_embd, n_embd], "f16"); tensor(p + "mlp/c_proj/b", [n_embd], "f32")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
