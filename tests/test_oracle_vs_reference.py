"""Pins the CPU oracle (oracle/quants_oracle.c) against the UNMODIFIED reference compiled into oracle/_ref.

CPU-only.  Bit-exact for dequantize / quantize; the dot products and mat-muls are compared at float
round-off (the reference does not define a summation order: its scalar, AVX2 and AVX-512 paths differ).
"""
import numpy as np
import pytest

from oracle import oracle as O

TYPES = list(O.HOT_TYPES) + list(O.NEXT_TYPES) + list(O.IQ_TYPES)      # hot-path formats, the SURVEY §8f-2 formats, the grid i-quants
IDS = [O.TYPE_NAMES[t] for t in TYPES]


def synth(n, offset=0.0):
    # the reference's own synthetic data: 0.1 + 2*cos(i + offset)  (tests/test-quantize-fns.cpp:31-35)
    i = np.arange(n, dtype=np.float32)
    return (0.1 + 2.0 * np.cos(i + np.float32(offset))).astype(np.float32)


def test_fp16_conversions_exhaustive(oracle):
    # every half -> float, and float(half) -> half round trips; numpy's float16 is IEEE RNE like F16C
    h = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    f_np = h.view(np.float16).astype(np.float32)
    f_or = np.array([oracle.lib.oq_fp16_to_fp32(int(v)) for v in h], dtype=np.float32)
    nan = np.isnan(f_np)
    assert np.array_equal(f_np[~nan].view(np.uint32), f_or[~nan].view(np.uint32))
    assert np.all(np.isnan(f_or[nan]))
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 100.0, 7e4)])
    x = np.concatenate([x, np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5],
                                    dtype=np.float32)])
    want = x.astype(np.float16).view(np.uint16)
    got = np.array([oracle.lib.oq_fp32_to_fp16(float(v)) for v in x], dtype=np.uint16)
    assert np.array_equal(want, got)


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_dequantize_bit_exact_reference_quantized(t, oracle, ref):
    n = 4096 * 4
    for off in (0.0, 1.0):
        blocks = ref.quantize(t, synth(n, off), 4, 4096)
        a = ref.dequantize(t, blocks, n)
        b = oracle.dequantize(t, blocks, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_dequantize_bit_exact_random_blocks(t, oracle, ref):
    rng = np.random.default_rng(100 + t)
    nb = 2048
    blocks = O.random_blocks(t, nb, rng)
    n = nb * oracle.blck_size(t)
    a = ref.dequantize(t, blocks, n)
    b = oracle.dequantize(t, blocks, n)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("t", [O.Q4_0, O.Q8_0, O.Q8_K], ids=["q4_0", "q8_0", "q8_K"])
def test_quantize_ref_bit_exact(t, oracle, ref):
    rng = np.random.default_rng(5)
    cases = [synth(4096), synth(4096, 1.0), rng.uniform(-1, 1, 8192).astype(np.float32),
             (rng.standard_normal(4096) * 30).astype(np.float32), np.zeros(512, dtype=np.float32),
             np.round(rng.uniform(-8, 8, 2048) * 2).astype(np.float32) / 2]          # many exact .5 ties
    for x in cases:
        # Q8_K has no from_float_ref in type_traits; its only quantizer is the CPU backend's (== quantize_row_q8_K_ref)
        a = ref.cpu_from_float(t, x) if t == O.Q8_K else ref.quantize_row_ref(t, x)
        b = oracle.quantize(t, x)
        if t == O.Q8_K:   # bsums of an all-zero block are left uninitialised by the reference
            a = a.reshape(-1, 292).copy(); b = b.reshape(-1, 292).copy()
            z = np.all(a[:, 4:260] == 0, axis=1)
            a[z, 260:] = 0; b[z, 260:] = 0
        assert np.array_equal(a, b)
        # ggml_quantize_chunk without imatrix is the same function
        if t != O.Q8_K:
            assert np.array_equal(ref.quantize(t, x, 1, x.size), b)


def test_activation_quantizer_matches_cpu_backend(oracle, ref):
    # what ggml_compute_forward_mul_mat actually calls (type_traits_cpu[vec_dot_type].from_float, ggml-cpu.c:7490-7509)
    rng = np.random.default_rng(11)
    for x in (synth(4096), rng.uniform(-1, 1, 4096).astype(np.float32), (rng.standard_normal(8192) * 5).astype(np.float32),
              np.round(rng.uniform(-127, 127, 4096)).astype(np.float32) / 2):
        assert np.array_equal(ref.cpu_from_float(O.Q8_0, x), oracle.quantize(O.Q8_0, x, simd_q8_0=True))
        assert np.array_equal(ref.cpu_from_float(O.Q8_1, x), oracle.quantize(O.Q8_1, x))
        a = ref.cpu_from_float(O.Q8_K, x); b = oracle.quantize(O.Q8_K, x)
        assert np.array_equal(a, b)


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_vec_dot_matches_cpu_backend(t, oracle, ref):
    rng = np.random.default_rng(20 + t)
    K = 4096
    vdt = oracle.vec_dot_type(t)
    assert vdt == ref.lib.probe_cpu_vec_dot_type(t)
    for trial in range(8):
        w = ref.quantize(t, rng.uniform(-1, 1, K).astype(np.float32), 1, K) if trial % 2 == 0 \
            else O.random_blocks(t, K // oracle.blck_size(t), rng)
        x = rng.uniform(-1, 1, K).astype(np.float32)
        yq = ref.cpu_from_float(vdt, x)
        a = ref.cpu_vec_dot(t, K, w, yq)
        b = oracle.vec_dot(t, K, w, yq)
        scale = float(np.linalg.norm(oracle.dequantize(t, w, K)) * np.linalg.norm(x)) + 1e-30
        assert abs(a - b) <= 2e-6 * scale, (a, b)
        # and the test the reference applies to its own vec_dot: |dot - float dot| / n < 0.02 (test-quantize-fns.cpp:172-183)
        fdot = float(np.dot(oracle.dequantize(t, w, K).astype(np.float64), x.astype(np.float64)))
        assert abs(b - fdot) / K < 0.02


@pytest.mark.parametrize("t", TYPES, ids=IDS)
@pytest.mark.parametrize("shape", [(16, 1, 256), (16, 9, 256), (33, 3, 1024), (512, 32, 256)])
def test_mul_mat_matches_cpu_backend(t, shape, oracle, ref):
    M, N, K = shape
    rng = np.random.default_rng(1234)
    W = ref.quantize(t, rng.uniform(-1, 1, M * K).astype(np.float32), M, K)
    X = rng.uniform(-1, 1, N * K).astype(np.float32)
    Yr, _ = ref.mul_mat(t, W, X, M, N, K, threads=2)
    Yo = oracle.mul_mat(t, W, X, M, N, K)
    assert O.nmse(Yo, Yr[0, 0]) < 1e-12
    # the oracle is also within the reference's own tolerance of the ideal (unquantized-activation) answer
    Yi = oracle.mul_mat(t, W, X, M, N, K, f64=True)
    assert O.nmse(Yo, Yi) < 5e-4


@pytest.mark.parametrize("t", [O.Q4_0, O.Q4_K, O.Q6_K], ids=["q4_0", "q4_K", "q6_K"])
@pytest.mark.parametrize("cfg", [(4, 1, 0, 1), (4, 2, 0, 32), (8, 4, 1, 32), (8, 2, 1, 1)])
def test_mul_mat_id_matches_cpu_backend(t, cfg, oracle, ref):
    # shapes of test_mul_mat_id (tests/test-backend-ops.cpp:4089-4119): m=512, k=256
    n_expert, n_used, bcast, n_tok = cfg
    M, K = 512, 256
    nb1 = 1 if bcast else n_used
    rng = np.random.default_rng(99)
    W = ref.quantize(t, rng.uniform(-1, 1, n_expert * M * K).astype(np.float32), n_expert * M, K)
    X = rng.uniform(-1, 1, n_tok * nb1 * K).astype(np.float32)
    ids = np.stack([rng.permutation(n_expert)[:n_used] for _ in range(n_tok)]).astype(np.int32)
    Yr, _ = ref.mul_mat_id(t, W, X, ids, M, K, n_expert, n_used, nb1, n_tok, threads=2)
    Yo = oracle.mul_mat_id(t, W, X, ids, M, K, n_expert, n_used, nb1, n_tok)
    assert O.nmse(Yo, Yr) < 1e-12


def test_gguf_file_round_trip_cpu(oracle, ref, tmp_path):
    """the on-disk format next to the path, with the reference's unmodified writer and loader: quantized weights written to a
    GGUF file, read back, MUL_MAT on the CPU backend == the oracle (the same probe drives the B200 backend in tests/test_gpu_backend_plugin.py)"""
    rng = np.random.default_rng(5)
    for t, M, N, K in [(O.Q4_K, 64, 2, 512), (O.Q8_0, 33, 1, 256), (O.IQ3_S, 16, 3, 256)]:
        W = O.random_blocks(t, M * K // oracle.blck_size(t), rng)
        X = rng.uniform(-1, 1, N * K).astype(np.float32)
        Y, _ = ref.gguf_mul_mat(t, W, X, M, N, K, "CPU", tmp_path / "w.gguf")
        assert O.nmse(Y, oracle.mul_mat(t, W, X, M, N, K)) < 1e-12, O.TYPE_NAMES[t]
