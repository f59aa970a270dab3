"""oracle/oracle.py — ctypes access to the CPU oracle and to the compiled reference.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py.  The product package (ggml_b200) must never
import this module (tests/test_product_isolation.py enforces it).

Two objects:
  * `Oracle` — liboracle_quants.so, this repo's plain-C restatement (oracle/quants_oracle.c).
  * `Ref`    — oracle/_ref/libggml_probe.so, a thin shim over the UNMODIFIED reference's public API
               (oracle/ref_probe.cpp) compiled from /root/reference by oracle/Makefile.  The
               `native` CPU-backend variant is used when the host CPU has every ISA extension it
               was compiled for (oracle/_ref/native/REQUIRED_FLAGS), else the x86-64-v3 build.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"

# enum ggml_type ids (reference include/ggml.h:351-390)
F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1 = 0, 1, 2, 3, 6, 7, 8, 9
Q2_K, Q3_K, Q4_K, Q5_K, Q6_K, Q8_K = 10, 11, 12, 13, 14, 15
IQ4_NL, IQ4_XS = 20, 23
IQ2_XXS, IQ3_XXS, IQ1_S = 16, 18, 19
IQ2_XS, IQ3_S, IQ2_S, IQ1_M, TQ1_0, TQ2_0 = 17, 21, 22, 29, 34, 35
TYPE_NAMES = {F32: "f32", F16: "f16", Q4_0: "q4_0", Q4_1: "q4_1", Q5_0: "q5_0", Q5_1: "q5_1", Q8_0: "q8_0", Q8_1: "q8_1",
              Q2_K: "q2_K", Q3_K: "q3_K", Q4_K: "q4_K", Q5_K: "q5_K", Q6_K: "q6_K", Q8_K: "q8_K", IQ4_NL: "iq4_nl", IQ4_XS: "iq4_xs", IQ2_XXS: "iq2_xxs", IQ3_XXS: "iq3_xxs", IQ1_S: "iq1_s",
              IQ2_XS: "iq2_xs", IQ3_S: "iq3_s", IQ2_S: "iq2_s", IQ1_M: "iq1_m", TQ1_0: "tq1_0", TQ2_0: "tq2_0"}
HOT_TYPES = (Q4_0, Q8_0, Q4_K, Q5_K, Q6_K)
# the next weight formats; oracle pinned this round, CUDA kernels follow
NEXT_TYPES = (Q4_1, Q5_0, Q5_1, Q2_K, Q3_K, IQ4_NL, IQ4_XS)
# the grid-codebook i-quants (round 2): generic mat-vec / MUL_MAT_ID / dequantize kernels
IQ_TYPES = (IQ2_XXS, IQ3_XXS, IQ1_S, IQ2_XS, IQ2_S, IQ3_S, IQ1_M, TQ1_0, TQ2_0)


def build(ref: bool = True) -> None:
    """(Re)build liboracle_quants.so and, where /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-j8", "-C", str(HERE), "oracle"] + (["ref"] if ref else []), check=True)
    # the reference's programs linked against the plug-in (gpt-2-backend-b200, -dump variants, gpt2-compare): needed by
    # tests/test_gpu_gpt2.py and bench.py's gpt-2 leg on the GPU box, so they are part of every build where the reference exists
    if ref and (HERE.parent / "ggml_b200" / "libggml-b200.so").exists():
        subprocess.run(["make", "-s", "-j8", "-C", str(HERE), "b200bins"], check=True)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    def __init__(self):
        so = HERE / "liboracle_quants.so"
        if not so.exists():
            build(ref=False)
        L = self.lib = C.CDLL(str(so))
        L.oq_fp16_to_fp32.restype = C.c_float
        L.oq_fp16_to_fp32.argtypes = [C.c_uint16]
        L.oq_fp32_to_fp16.restype = C.c_uint16
        L.oq_fp32_to_fp16.argtypes = [C.c_float]
        L.oq_row_size.restype = C.c_size_t
        L.oq_row_size.argtypes = [C.c_int, C.c_int64]
        L.oq_blck_size.restype = C.c_int64
        L.oq_blck_size.argtypes = [C.c_int]
        L.oq_type_size.restype = C.c_size_t
        L.oq_type_size.argtypes = [C.c_int]
        L.oq_dequantize_row.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.oq_quantize_row_ref.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.oq_quantize_row_q8_0_simd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oq_quantize_row_q8_0_simd.restype = None
        L.oq_quantize_row_q8_1_simd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oq_quantize_row_q8_1_simd.restype = None
        L.oq_vec_dot.restype = C.c_float
        L.oq_vec_dot.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.oq_vec_dot_type.argtypes = [C.c_int]
        L.oq_mul_mat.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]
        L.oq_mul_mat_f64.argtypes = L.oq_mul_mat.argtypes
        L.oq_mul_mat_id.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p] + [C.c_int64] * 6

    def row_size(self, t, k):
        return int(self.lib.oq_row_size(t, k))

    def blck_size(self, t):
        return int(self.lib.oq_blck_size(t))

    def dequantize(self, t, blocks: np.ndarray, n: int) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        out = np.empty(n, dtype=np.float32)
        rc = self.lib.oq_dequantize_row(t, _p(blocks), _p(out), n)
        assert rc == 0, f"oracle cannot dequantize type {t}"
        return out

    def quantize(self, t, x, simd_q8_0: bool = False) -> np.ndarray:
        x = _f32(x).ravel()
        out = np.zeros(self.row_size(t, x.size), dtype=np.uint8)
        if t == Q8_1:                       # only the CPU backend's (SIMD-flavoured) activation quantizer is restated
            self.lib.oq_quantize_row_q8_1_simd(_p(x), _p(out), x.size)
        elif simd_q8_0:
            assert t == Q8_0
            self.lib.oq_quantize_row_q8_0_simd(_p(x), _p(out), x.size)
        else:
            assert self.lib.oq_quantize_row_ref(t, _p(x), _p(out), x.size) == 0
        return out

    def vec_dot(self, t, k, wrow: np.ndarray, yq: np.ndarray) -> float:
        return float(self.lib.oq_vec_dot(t, k, _p(np.ascontiguousarray(wrow)), _p(np.ascontiguousarray(yq))))

    def vec_dot_type(self, t):
        return int(self.lib.oq_vec_dot_type(t))

    def mul_mat(self, t, W: np.ndarray, X, M, N, K, f64: bool = False) -> np.ndarray:
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        assert W.size == self.row_size(t, K) * M and X.size == N * K
        Y = np.empty((N, M), dtype=np.float32)
        fn = self.lib.oq_mul_mat_f64 if f64 else self.lib.oq_mul_mat
        assert fn(t, _p(W), _p(X), _p(Y), M, N, K) == 0
        return Y

    def mul_mat_id(self, t, W, X, ids, M, K, n_expert, n_used, nb1, n_tok) -> np.ndarray:
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        assert ids.shape == (n_tok, ids.shape[1]) and ids.shape[1] >= n_used
        Y = np.empty((n_tok, n_used, M), dtype=np.float32)
        rc = self.lib.oq_mul_mat_id(t, _p(W), _p(X), _p(ids), ids.shape[1], _p(Y), M, K, n_expert, n_used, nb1, n_tok)
        assert rc == 0, rc
        return Y


def _host_has_native_isa() -> bool:
    req = REF_DIR / "native" / "REQUIRED_FLAGS"
    so = REF_DIR / "native" / "libggml-cpu.so"
    if not (req.exists() and so.exists()):
        return False
    try:
        flags = set(re.search(r"^flags\s*:\s*(.*)$", Path("/proc/cpuinfo").read_text(), re.M).group(1).split())
    except Exception:
        return False
    need = re.findall(r"__([A-Z0-9_]+?)__ 1", req.read_text())
    ren = {"AVX512VNNI": "avx512_vnni", "AVX512VBMI2": "avx512_vbmi2", "AVX512BF16": "avx512_bf16",
           "AVX512FP16": "avx512_fp16", "AVX512BITALG": "avx512_bitalg", "AVX512VPOPCNTDQ": "avx512_vpopcntdq",
           "AVXVNNI": "avx_vnni", "AVX512VP2INTERSECT": "avx512_vp2intersect"}
    for n in need:
        if ren.get(n, n.lower()) not in flags:
            return False
    return True


def ref_available() -> bool:
    return (REF_DIR / "libggml_probe.so").exists()


def ref_env(native: bool | None = None) -> dict:
    """Environment for running oracle/_ref binaries (test-backend-ops, gpt-2-*)."""
    if native is None:
        native = _host_has_native_isa()
    env = dict(os.environ)
    paths = ([str(REF_DIR / "native")] if native else []) + [str(REF_DIR)]
    env["LD_LIBRARY_PATH"] = ":".join(paths + [env.get("LD_LIBRARY_PATH", "")]).rstrip(":")
    return env


class Ref:
    """The unmodified reference, through oracle/_ref/libggml_probe.so."""

    _loaded = None

    def __init__(self, native: bool | None = None):
        if not ref_available():
            build(ref=True)
        if native is None:
            native = _host_has_native_isa()
        if Ref._loaded is None:
            C.CDLL(str(REF_DIR / "libggml-base.so"), mode=C.RTLD_GLOBAL)
            C.CDLL(str(REF_DIR / ("native" if native else ".") / "libggml-cpu.so"), mode=C.RTLD_GLOBAL)
            C.CDLL(str(REF_DIR / "libggml.so"), mode=C.RTLD_GLOBAL)
            Ref._loaded = (C.CDLL(str(REF_DIR / "libggml_probe.so")), native)
        L, self.native = Ref._loaded
        self.lib = L
        L.probe_load_backend.argtypes = [C.c_char_p]
        L.probe_dev_name.restype = C.c_char_p
        L.probe_dev_name.argtypes = [C.c_int]
        L.probe_dev_desc.restype = C.c_char_p
        L.probe_dev_desc.argtypes = [C.c_int]
        L.probe_row_size.restype = C.c_size_t
        L.probe_row_size.argtypes = [C.c_int, C.c_int64]
        L.probe_quantize.restype = C.c_size_t
        L.probe_quantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.probe_quantize_row_ref.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.probe_quantize_row_ref.restype = None
        L.probe_dequantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.probe_dequantize.restype = None
        L.probe_cpu_from_float.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.probe_cpu_from_float.restype = None
        L.probe_cpu_vec_dot.restype = C.c_float
        L.probe_cpu_vec_dot.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.probe_mul_mat.restype = C.c_double
        L.probe_mul_mat.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int64] * 7 + [C.c_int] * 5
        L.probe_mul_mat_sweep.restype = C.c_double
        L.probe_mul_mat_sweep.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int64] * 3 + [C.c_int] * 5
        L.probe_mul_mat_split.restype = C.c_double
        L.probe_mul_mat_split.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int64] * 3 + [C.c_int]
        L.probe_gguf_mul_mat.restype = C.c_double
        L.probe_gguf_mul_mat.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int64] * 3
        L.probe_mul_mat_id.restype = C.c_double
        L.probe_mul_mat_id.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int64] * 6 + [C.c_int] * 2

    # --- formats
    def row_size(self, t, k):
        return int(self.lib.probe_row_size(t, k))

    def quantize(self, t, x, nrows, k) -> np.ndarray:
        """ggml_quantize_chunk(type, x, …, imatrix=NULL)"""
        x = _f32(x)
        out = np.zeros(self.row_size(t, k) * nrows, dtype=np.uint8)
        n = self.lib.probe_quantize(t, _p(x), _p(out), nrows, k)
        assert n == out.size
        return out

    def quantize_row_ref(self, t, x) -> np.ndarray:
        x = _f32(x).ravel()
        out = np.zeros(self.row_size(t, x.size), dtype=np.uint8)
        self.lib.probe_quantize_row_ref(t, _p(x), _p(out), x.size)
        return out

    def dequantize(self, t, blocks, n) -> np.ndarray:
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        out = np.empty(n, dtype=np.float32)
        self.lib.probe_dequantize(t, _p(blocks), _p(out), n)
        return out

    def cpu_from_float(self, t, x) -> np.ndarray:
        x = _f32(x).ravel()
        out = np.zeros(self.row_size(t, x.size), dtype=np.uint8)
        self.lib.probe_cpu_from_float(t, _p(x), _p(out), x.size)
        return out

    def cpu_vec_dot(self, t, k, wrow, yq) -> float:
        return float(self.lib.probe_cpu_vec_dot(t, k, _p(np.ascontiguousarray(wrow)), _p(np.ascontiguousarray(yq))))

    # --- backends
    def load_backend(self, path) -> bool:
        return bool(self.lib.probe_load_backend(str(path).encode()))

    def devices(self):
        return [self.lib.probe_dev_name(i).decode() for i in range(self.lib.probe_dev_count())]

    def hw_threads(self):
        return int(self.lib.probe_hw_threads())

    def mul_mat(self, t, W, X, M, N, K, dev="CPU", batch=(1, 1, 1, 1), threads=0, repeat=1, iters=1, warmup=0,
                e2e=False):
        """returns (Y[ne3b, ne2b, N, M], seconds per mul_mat)"""
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        ne2a, ne3a, ne2b, ne3b = batch
        assert W.size == self.row_size(t, K) * M * ne2a * ne3a, (W.size, self.row_size(t, K) * M * ne2a * ne3a)
        assert X.size == K * N * ne2b * ne3b
        Y = np.empty((ne3b, ne2b, N, M), dtype=np.float32)
        s = self.lib.probe_mul_mat(dev.encode(), t, _p(W), _p(X), _p(Y), M, N, K, ne2a, ne3a, ne2b, ne3b,
                                   threads, repeat, iters, warmup, int(e2e))
        if s < 0:
            raise RuntimeError(f"probe_mul_mat({dev}) failed: {s}")
        return Y, float(s)

    def mul_mat_sweep(self, t, W, X, M, N, K, nw, dev="CPU", threads=0, iters=1, warmup=0, e2e=False):
        """nw distinct weight tensors (same bytes, separate memory), nw MUL_MAT nodes per graph -> (Y[N, M], seconds per mul_mat)"""
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        assert W.size == self.row_size(t, K) * M and X.size == K * N
        Y = np.empty((N, M), dtype=np.float32)
        s = self.lib.probe_mul_mat_sweep(dev.encode(), t, _p(W), _p(X), _p(Y), M, N, K, nw, threads, iters, warmup, int(e2e))
        if s < 0:
            raise RuntimeError(f"probe_mul_mat_sweep({dev}) failed: {s}")
        return Y, float(s)

    def mul_mat_split(self, t, W, X, M, N, K, dev, main_device=0, tensor_split=None, iters=1):
        """MUL_MAT with W in the backend's split buffer type (rows sharded over its devices) -> (Y[N, M], seconds per graph_compute)"""
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        assert W.size == self.row_size(t, K) * M and X.size == K * N
        Y = np.empty((N, M), dtype=np.float32)
        ts = None if tensor_split is None else _f32(list(tensor_split) + [0.0] * (16 - len(tensor_split)))
        s = self.lib.probe_mul_mat_split(dev.encode(), main_device, None if ts is None else _p(ts), t, _p(W), _p(X), _p(Y), M, N, K, iters)
        if s < 0:
            raise RuntimeError(f"probe_mul_mat_split({dev}) failed: {s}")
        return Y, float(s)

    def gguf_mul_mat(self, t, W, X, M, N, K, dev, path):
        """W -> GGUF file (reference writer) -> reference loader -> device buffer via the pinned host buffer type + async upload -> MUL_MAT.
        Returns (Y[N, M], pinned: bool)"""
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        Y = np.empty((N, M), dtype=np.float32)
        rc = self.lib.probe_gguf_mul_mat(dev.encode(), str(path).encode(), t, _p(W), _p(X), _p(Y), M, N, K)
        if rc < 0:
            raise RuntimeError(f"probe_gguf_mul_mat({dev}) failed: {rc}")
        return Y, rc == 0.0

    def mul_mat_id(self, t, W, X, ids, M, K, n_expert, n_used, nb1, n_tok, dev="CPU", threads=0, iters=1):
        W = np.ascontiguousarray(W, dtype=np.uint8)
        X = _f32(X)
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        assert ids.shape == (n_tok, n_used)
        Y = np.empty((n_tok, n_used, M), dtype=np.float32)
        s = self.lib.probe_mul_mat_id(dev.encode(), t, _p(W), _p(X), _p(ids), _p(Y), M, K, n_expert, n_used, nb1, n_tok,
                                      threads, iters)
        if s < 0:
            raise RuntimeError(f"probe_mul_mat_id({dev}) failed: {s}")
        return Y, float(s)


def random_blocks(t: int, nblocks: int, rng: np.random.Generator, scale: float = 0.05) -> np.ndarray:
    """Arbitrary-but-valid packed blocks: uniformly random code bytes (every nibble / 6-bit scale /
    high-bit pattern occurs) with finite fp16 scales of magnitude ~`scale`."""
    ts = {Q4_0: 18, Q8_0: 34, Q4_K: 144, Q5_K: 176, Q6_K: 210, Q4_1: 20, Q5_0: 22, Q5_1: 24, Q2_K: 84, Q3_K: 110, IQ4_NL: 18, IQ4_XS: 136,
          IQ2_XXS: 66, IQ3_XXS: 98, IQ1_S: 50, IQ2_XS: 74, IQ2_S: 82, IQ3_S: 110, IQ1_M: 56, TQ1_0: 54, TQ2_0: 66}[t]
    b = rng.integers(0, 256, size=(nblocks, ts), dtype=np.uint8)

    def put_half(col, vals):
        h = vals.astype(np.float16).view(np.uint16)
        b[:, col] = (h & 0xFF).astype(np.uint8)
        b[:, col + 1] = (h >> 8).astype(np.uint8)

    if t in (Q4_0, Q8_0):
        put_half(0, rng.uniform(-scale, scale, nblocks))
    elif t in (Q4_K, Q5_K):
        put_half(0, rng.uniform(0, scale / 32, nblocks))
        put_half(2, rng.uniform(0, scale / 32, nblocks))
    elif t == Q6_K:
        put_half(208, rng.uniform(-scale / 64, scale / 64, nblocks))
    elif t in (Q5_0, IQ4_NL):
        put_half(0, rng.uniform(-scale / (8 if t == IQ4_NL else 1), scale / (8 if t == IQ4_NL else 1), nblocks))
    elif t in (Q4_1, Q5_1):
        put_half(0, rng.uniform(0, scale, nblocks))
        put_half(2, rng.uniform(-scale, scale, nblocks))
    elif t == Q2_K:
        put_half(80, rng.uniform(0, scale / 8, nblocks))
        put_half(82, rng.uniform(0, scale / 8, nblocks))
    elif t == Q3_K:
        put_half(108, rng.uniform(-scale / 32, scale / 32, nblocks))
    elif t == IQ4_XS:
        put_half(0, rng.uniform(-scale / 256, scale / 256, nblocks))
    elif t in (IQ2_XXS, IQ3_XXS, IQ1_S, IQ2_XS, IQ2_S, IQ3_S):
        put_half(0, rng.uniform(-scale / 16, scale / 16, nblocks))
    elif t == IQ1_M:                                       # the f16 super-scale lives in the top nibbles of the four u16 scale words
        h = rng.uniform(-scale / 16, scale / 16, nblocks).astype(np.float16).view(np.uint16).astype(np.uint32)
        for i in range(4):
            b[:, 49 + 2 * i] = (b[:, 49 + 2 * i] & 0x0F) | (((h >> (4 * i)) & 0xF) << 4).astype(np.uint8)
    elif t in (TQ1_0, TQ2_0):
        put_half(52 if t == TQ1_0 else 64, rng.uniform(-scale, scale, nblocks))
        if t == TQ1_0:                                     # valid base-3 bytes: 5 trits -> ceil(v * 256 / 243), 4 trits (qh) likewise
            v5 = rng.integers(0, 243, size=(nblocks, 48)); b[:, 0:48] = ((v5 * 256 + 242) // 243).astype(np.uint8)
            v4 = rng.integers(0, 81, size=(nblocks, 4)) * 3; b[:, 48:52] = ((v4 * 256 + 242) // 243).astype(np.uint8)
        else:                                              # 2-bit codes 0..2 only
            c = rng.integers(0, 3, size=(nblocks, 64, 4))
            b[:, 0:64] = (c[..., 0] | (c[..., 1] << 2) | (c[..., 2] << 4) | (c[..., 3] << 6)).astype(np.uint8)
    return b.reshape(-1)


def nmse(a: np.ndarray, b: np.ndarray) -> float:
    """normalized mean squared error = mse(a, b) / mse(a, 0), exactly as tests/test-backend-ops.cpp:174-188
    (a = the backend under test, b = the CPU reference)"""
    a = np.asarray(a, dtype=np.float32).ravel()
    b = np.asarray(b, dtype=np.float32).ravel()
    diff = (a - b).astype(np.float32)
    num = float(np.sum((diff * diff).astype(np.float64)))
    den = float(np.sum((a * a).astype(np.float64)))
    return num / den if den > 0 else num
