"""CPU-only: the oracle against the committed golden vectors (produced by the unmodified reference,
tests/golden/make_golden.py).  This is what keeps the oracle pinned on a box without /root/reference."""
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

G = Path(__file__).resolve().parent / "golden"
TYPES = list(O.HOT_TYPES) + list(O.NEXT_TYPES) + list(O.IQ_TYPES)      # hot-path formats, the SURVEY §8f-2 formats, the grid i-quants
IDS = [O.TYPE_NAMES[t] for t in TYPES]


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_dequantize_golden(t, oracle):
    z = np.load(G / f"quant_{O.TYPE_NAMES[t]}.npz")
    assert np.array_equal(oracle.dequantize(t, z["blocks"], z["x"].size).view(np.uint32), z["deq"].view(np.uint32))
    assert np.array_equal(oracle.dequantize(t, z["rnd_blocks"], z["rnd_deq"].size).view(np.uint32), z["rnd_deq"].view(np.uint32))
    # the reference's own round-trip bound (tests/test-quantize-fns.cpp:17-25,150-163)
    rmse = np.sqrt(np.mean((z["deq"][:2048] - z["x"][:2048]) ** 2))
    assert rmse < {O.Q4_0: 0.002 * 1e9, O.Q8_0: 0.002 * 1e9}.get(t, 1e9)   # sanity only: data here is not the reference's length


def test_quantizers_golden(oracle):
    z = np.load(G / "act_q8.npz")
    x = z["x"]
    assert np.array_equal(oracle.quantize(O.Q8_0, x, simd_q8_0=True), z["q8_0"])
    assert np.array_equal(oracle.quantize(O.Q8_1, x), z["q8_1"])
    assert np.array_equal(oracle.quantize(O.Q8_0, x), z["q8_0_ref"])
    assert np.array_equal(oracle.quantize(O.Q4_0, x), z["q4_0_ref"])
    a = oracle.quantize(O.Q8_K, x).reshape(-1, 292).copy(); b = z["q8_K"].reshape(-1, 292).copy()
    zero = np.all(b[:, 4:260] == 0, axis=1)
    a[zero, 260:] = 0; b[zero, 260:] = 0       # bsums of all-zero blocks are uninitialised in the reference
    assert np.array_equal(a, b)


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_mul_mat_golden(t, oracle):
    z = np.load(G / f"mulmat_{O.TYPE_NAMES[t]}.npz")
    for ci in range(int(z["ncases"])):
        M, N, K = (int(v) for v in z[f"shape{ci}"])
        Y = oracle.mul_mat(t, z[f"W{ci}"], z[f"X{ci}"], M, N, K)
        assert O.nmse(Y, z[f"Y{ci}"]) < 1e-12, (ci, M, N, K)


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_mul_mat_id_golden(t, oracle):
    z = np.load(G / f"mulmatid_{O.TYPE_NAMES[t]}.npz")
    ne, nu, nb1, ntok, M, K = (int(v) for v in z["cfg"])
    Y = oracle.mul_mat_id(t, z["W"], z["X"], z["ids"], M, K, ne, nu, nb1, ntok)
    assert O.nmse(Y, z["Y"]) < 1e-12
