// ops.cu — the small device ops either side of the quantized mat-mul in the examples/gpt-2 graph
// (SURVEY.md §8f-1): GET_ROWS, ADD/MUL/SUB/DIV with broadcast, NORM / RMS_NORM, SCALE, DIAG_MASK_INF, SOFT_MAX,
// unary GELU/SILU/RELU/…, CPY/CONT/DUP (strided, f32/f16 and f32 -> Q8_0/Q4_0), and the float (f32/f16 x f32)
// batched strided MUL_MAT used for KQ and KQV.  They exist so that `gpt-2-backend` runs entirely on the device
// (it has no scheduler to fall back to the CPU).  Semantics follow the CPU backend (src/ggml-cpu/ggml-cpu.c):
//   get_rows :8560-8760   add/mul bcast :4660-5560   norm :8915-8975   rms_norm :8990-9050   scale :8300-8345
//   diag_mask :9745-9805  soft_max :9810-9925        gelu :6520-6570 (+ fp16 table, ggml-cpu.c:1355)   dup/cpy :3220-4300
// They replace the reference's getrows.cu, binbcast.cu, norm.cu, scale.cu, diagmask.cu, softmax.cu, unary.cu, cpy.cu, mmv.cu.
#include "b200_internal.h"
#include "b200_quants.cuh"
#include "b200_dequant.cuh"

#include <cfloat>
#include <cstdlib>

namespace b200 {

// Programmatic dependent launch for every small op: each kernel is launched with the programmatic-serialization attribute (launch_pdl),
// so it becomes resident while its predecessor still runs (the launch latency of a ~2 us kernel chain is hidden), waits for the
// predecessor's completion before touching any global memory (griddepcontrol.wait: completion stays transitive along the stream),
// and only then lets ITS successor launch -- a one-kernel lookahead.  A quantized mat-vec launched after one of these small ops starts
// its weight prefetch while the small op runs and waits for this kernel's results before reading them.
__device__ __forceinline__ void pdl_trigger() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    static const bool use_pdl = !(getenv("GGML_B200_NO_PDL") && atoi(getenv("GGML_B200_NO_PDL")) != 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct tdesc {      // device-side copy of ggml_b200_tensor
    uint8_t * data; int32_t type; int64_t ne[4]; size_t nb[4];
};
static inline tdesc T(const ggml_b200_tensor * t) {
    tdesc d; d.data = (uint8_t *)t->data; d.type = t->type;
    for (int i = 0; i < 4; ++i) { d.ne[i] = t->ne[i]; d.nb[i] = t->nb[i]; }
    return d;
}
static inline int64_t nelem(const tdesc & t) { return t.ne[0] * t.ne[1] * t.ne[2] * t.ne[3]; }
static inline int64_t nrows(const tdesc & t) { return t.ne[1] * t.ne[2] * t.ne[3]; }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide reduce (blockDim multiple of 32, <= 1024); result broadcast to all threads
template <bool MAX> __device__ __forceinline__ float block_reduce(float v, float * sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = MAX ? warp_max_f(v) : warp_sum_f(v);
    if (nw == 1) return v;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = lane < nw ? sh[lane] : (MAX ? -INFINITY : 0.0f);
    r = MAX ? warp_max_f(r) : warp_sum_f(r);
    return r;
}

// ------------------------------------------------------------------ dequantize one element (bit-exact, as dequant.cu)
template <int T> __device__ __forceinline__ float elem_via_dequant4(const uint8_t * row, int64_t i) {
    float o[4];
    dequant4<T>(row, i & ~(int64_t)3, o);
    return o[i & 3];
}
__device__ __forceinline__ float load_elem(const uint8_t * row, int type, int64_t i) {
    switch (type) {
        case T_F32: return ((const float *)row)[i];
        case T_F16: return __half2float(((const __half *)row)[i]);
        case T_Q4_0: {
            const uint8_t * b = row + (i / 32) * 18; const int j = (int)(i % 32);
            const int q = j < 16 ? (b[2 + j] & 0x0F) : (b[2 + j - 16] >> 4);
            return __fmul_rn((float)(q - 8), h2f(load_u16(b)));
        }
        case T_Q8_0: {
            const uint8_t * b = row + (i / 32) * 34;
            return __fmul_rn((float)(int8_t)b[2 + (i % 32)], h2f(load_u16(b)));
        }
        case T_Q4_K: case T_Q5_K: {
            const int BY = type == T_Q4_K ? 144 : 176;
            const uint8_t * b = row + (i / 256) * BY;
            const int w = (int)(i % 256), c = w / 64, l = w % 32, hi = (w % 64) / 32, j = 2 * c + hi;
            const uint8_t * s = b + 4;
            int sc, mn;
            if (j < 4) { sc = s[j] & 63; mn = s[j + 4] & 63; }
            else       { sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4); mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
            const uint8_t qb = b[(type == T_Q5_K ? 48 : 16) + 32 * c + l];
            int v = hi ? (qb >> 4) : (qb & 0x0F);
            if (type == T_Q5_K) v += ((b[16 + l] >> j) & 1) << 4;
            return __fsub_rn(__fmul_rn(__fmul_rn(h2f(load_u16(b)), (float)sc), (float)v), __fmul_rn(h2f(load_u16(b + 2)), (float)mn));
        }
        case T_Q6_K: {
            const uint8_t * b = row + (i / 256) * 210;
            const int w = (int)(i % 256), h = w / 128, pos = (w % 128) / 32, l = w % 32;
            const uint8_t ql = b[64 * h + (pos & 1) * 32 + l], qh = b[128 + 32 * h + l];
            const int lo = pos >= 2 ? (ql >> 4) : (ql & 0x0F);
            const int v = (int)(int8_t)(lo | (((qh >> (2 * pos)) & 3) << 4)) - 32;
            const int sc = (int)(int8_t)b[192 + 8 * h + l / 16 + 2 * pos];
            return __fmul_rn(__fmul_rn(h2f(load_u16(b + 208)), (float)sc), (float)v);
        }
        // every other block format: through the element decoders of b200_dequant.cuh (four consecutive weights, pick one)
        case T_Q4_1:   return elem_via_dequant4<T_Q4_1>(row, i);
        case T_Q5_0:   return elem_via_dequant4<T_Q5_0>(row, i);
        case T_Q5_1:   return elem_via_dequant4<T_Q5_1>(row, i);
        case T_Q2_K:   return elem_via_dequant4<T_Q2_K>(row, i);
        case T_Q3_K:   return elem_via_dequant4<T_Q3_K>(row, i);
        case T_IQ4_NL: return elem_via_dequant4<T_IQ4_NL>(row, i);
        case T_IQ4_XS: return elem_via_dequant4<T_IQ4_XS>(row, i);
        case T_IQ2_XXS: return elem_via_dequant4<T_IQ2_XXS>(row, i);
        case T_IQ3_XXS: return elem_via_dequant4<T_IQ3_XXS>(row, i);
        case T_IQ1_S:  return elem_via_dequant4<T_IQ1_S>(row, i);
        case T_IQ2_XS: return elem_via_dequant4<T_IQ2_XS>(row, i);
        case T_IQ2_S: return elem_via_dequant4<T_IQ2_S>(row, i);
        case T_IQ3_S: return elem_via_dequant4<T_IQ3_S>(row, i);
        case T_IQ1_M: return elem_via_dequant4<T_IQ1_M>(row, i);
        case T_TQ1_0: return elem_via_dequant4<T_TQ1_0>(row, i);
        case T_TQ2_0: return elem_via_dequant4<T_TQ2_0>(row, i);
        default: return __int_as_float(0x7fc00000);          // unreachable (ggml_b200_op_get_rows rejects unknown types): NaN, never a silent 0
    }
}

// ------------------------------------------------------------------ GET_ROWS
__global__ void get_rows_kernel(tdesc src, tdesc ids, tdesc dst) {
    pdl_trigger();
    // one CTA per destination row (i10, i11, i12)
    const int64_t r = blockIdx.x;
    const int64_t i10 = r % ids.ne[0], i11 = (r / ids.ne[0]) % ids.ne[1], i12 = r / (ids.ne[0] * ids.ne[1]);
    const int32_t i01 = *(const int32_t *)(ids.data + i10 * ids.nb[0] + i11 * ids.nb[1] + i12 * ids.nb[2]);
    const uint8_t * srow = src.data + (int64_t)i01 * src.nb[1] + i11 * src.nb[2] + i12 * src.nb[3];
    float * drow = (float *)(dst.data + i10 * dst.nb[1] + i11 * dst.nb[2] + i12 * dst.nb[3]);
    for (int64_t i = threadIdx.x; i < src.ne[0]; i += blockDim.x) drow[i] = load_elem(srow, src.type, i);
}

// ------------------------------------------------------------------ binary ops with broadcast (f32)
template <int OP> __device__ __forceinline__ float bin_op(float a, float b) {
    if (OP == 0) return a + b;
    if (OP == 1) return a * b;
    if (OP == 2) return a - b;
    return a / b;
}
template <int OP> __global__ void bin_bcast_kernel(tdesc a, tdesc b, tdesc d, int64_t n) {
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t i0 = i % d.ne[0], i1 = (i / d.ne[0]) % d.ne[1], i2 = (i / (d.ne[0] * d.ne[1])) % d.ne[2], i3 = i / (d.ne[0] * d.ne[1] * d.ne[2]);
    const float x = *(const float *)(a.data + i0 * a.nb[0] + i1 * a.nb[1] + i2 * a.nb[2] + i3 * a.nb[3]);
    const float y = *(const float *)(b.data + (i0 % b.ne[0]) * b.nb[0] + (i1 % b.ne[1]) * b.nb[1] + (i2 % b.ne[2]) * b.nb[2] + (i3 % b.ne[3]) * b.nb[3]);
    *(float *)(d.data + i0 * d.nb[0] + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]) = bin_op<OP>(x, y);
}

// ------------------------------------------------------------------ NORM / RMS_NORM (rows contiguous along dim 0)
template <bool RMS> __global__ void norm_kernel(tdesc s, tdesc d, float eps) {
    pdl_trigger();
    __shared__ float sh[32];
    const int64_t r = blockIdx.x;
    const int64_t i1 = r % s.ne[1], i2 = (r / s.ne[1]) % s.ne[2], i3 = r / (s.ne[1] * s.ne[2]);
    const float * x = (const float *)(s.data + i1 * s.nb[1] + i2 * s.nb[2] + i3 * s.nb[3]);
    float * y = (float *)(d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3]);
    const int64_t n = s.ne[0];
    float mean = 0.0f;
    if (!RMS) {
        float sum = 0.0f;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) sum += x[i];
        mean = block_reduce<false>(sum, sh) / (float)n;
    }
    float sq = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float v = x[i] - mean; sq += v * v; }
    const float var = block_reduce<false>(sq, sh) / (float)n;
    const float scale = 1.0f / sqrtf(var + eps);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) y[i] = (x[i] - mean) * scale;
}

// NORM -> MUL(gain) -> ADD(bias) in one pass; every intermediate tensor of the three ggml nodes is still written
template <bool RMS> __global__ void norm_affine_kernel(tdesc s, tdesc d1, const float * gain, tdesc d2, const float * bias, tdesc d3, float eps) {
    pdl_trigger();
    __shared__ float sh[32];
    const int64_t r = blockIdx.x;
    const int64_t i1 = r % s.ne[1], i2 = (r / s.ne[1]) % s.ne[2], i3 = r / (s.ne[1] * s.ne[2]);
    const float * x = (const float *)(s.data + i1 * s.nb[1] + i2 * s.nb[2] + i3 * s.nb[3]);
    float * y1 = (float *)(d1.data + i1 * d1.nb[1] + i2 * d1.nb[2] + i3 * d1.nb[3]);
    float * y2 = (float *)(d2.data + i1 * d2.nb[1] + i2 * d2.nb[2] + i3 * d2.nb[3]);
    float * y3 = (float *)(d3.data + i1 * d3.nb[1] + i2 * d3.nb[2] + i3 * d3.nb[3]);
    const int64_t n = s.ne[0];
    float mean = 0.0f;
    if (!RMS) {
        float sum = 0.0f;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) sum += x[i];
        mean = block_reduce<false>(sum, sh) / (float)n;
    }
    float sq = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float v = x[i] - mean; sq += v * v; }
    const float var = block_reduce<false>(sq, sh) / (float)n;
    const float scale = 1.0f / sqrtf(var + eps);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = (x[i] - mean) * scale, b = __fmul_rn(a, gain[i]), c = __fadd_rn(b, bias[i]);   // separately rounded, like the three ggml ops
        y1[i] = a; y2[i] = b; y3[i] = c;
    }
}

// ------------------------------------------------------------------ SCALE, DIAG_MASK_INF, unary
__global__ void scale_kernel(const float * x, float * y, float s, int64_t n) {
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i] * s;
}
__global__ void diag_mask_inf_kernel(const float * x, float * y, int64_t ne0, int64_t ne1, int n_past, int64_t n) {
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = i % ne0, r = (i / ne0) % ne1;
    y[i] = c > n_past + r ? -INFINITY : x[i];
}
enum { U_GELU = 0, U_SILU = 1, U_RELU = 2, U_TANH = 3, U_NEG = 4, U_ABS = 5, U_GELU_QUICK = 6, U_SIGMOID = 7, U_EXP = 8, U_SQR = 9, U_SQRT = 10 };
__global__ void unary_kernel(int uop, const float * x, float * y, int64_t n) {
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    float r;
    switch (uop) {
        case U_GELU: r = gelu_ggml(v); break;
        case U_SILU: r = v / (1.0f + expf(-v)); break;
        case U_RELU: r = fmaxf(v, 0.0f); break;
        case U_TANH: r = tanhf(v); break;
        case U_NEG: r = -v; break;
        case U_ABS: r = fabsf(v); break;
        case U_GELU_QUICK: r = v * (1.0f / (1.0f + expf(-1.702f * v))); break;
        case U_SIGMOID: r = 1.0f / (1.0f + expf(-v)); break;
        case U_EXP: r = expf(v); break;
        case U_SQR: r = v * v; break;
        default: r = sqrtf(v); break;
    }
    y[i] = r;
}

// ------------------------------------------------------------------ SOFT_MAX (rows contiguous): softmax(x*scale + mask*slope)
// diag_n_past >= 0 folds a preceding GGML_OP_DIAG_MASK_INF (and, through `scale`, a preceding GGML_OP_SCALE) into the row pass:
// element i0 of row i1 is -inf where i0 > diag_n_past + i1 (ggml_compute_forward_diag_mask_f32, src/ggml-cpu/ggml-cpu.c)
__global__ void soft_max_kernel(const float * x, const uint8_t * mask, int mask_type, float * y, int64_t ne0, int64_t ne1, int64_t ne2,
                                float scale, float max_bias, float m0, float m1, uint32_t n_head_log2, int diag_n_past) {
    pdl_trigger();
    __shared__ float sh[32];
    const int64_t r = blockIdx.x;                 // row = i1 + ne1 * (i2 + ne2 * i3)
    const int64_t i1 = r % ne1;
    const float * xr = x + r * ne0;
    float * yr = y + r * ne0;
    float slope = 1.0f;
    if (max_bias > 0.0f) {
        const uint32_t h = (uint32_t)((r / ne1) % ne2);
        slope = h < n_head_log2 ? powf(m0, (float)(h + 1)) : powf(m1, (float)(2 * (h - n_head_log2) + 1));
    }
    const uint8_t * mr = mask ? mask + (size_t)i1 * ne0 * (mask_type == T_F16 ? 2 : 4) : nullptr;
    auto val = [&](int64_t i) {
        float v = xr[i] * scale;
        if (diag_n_past >= 0 && i > (int64_t)diag_n_past + i1) v = -INFINITY;
        if (mr) v += slope * (mask_type == T_F16 ? __half2float(((const __half *)mr)[i]) : ((const float *)mr)[i]);
        return v;
    };
    float mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) mx = fmaxf(mx, val(i));
    mx = block_reduce<true>(mx, sh);
    float sum = 0.0f;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) { const float e = expf(val(i) - mx); yr[i] = e; sum += e; }
    sum = block_reduce<false>(sum, sh);
    const float inv = 1.0f / sum;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) yr[i] *= inv;
}

// ------------------------------------------------------------------ CPY / CONT / DUP (same element count, any strides)
__device__ __forceinline__ size_t offset_of(const tdesc & t, int64_t i) {    // i = linear element index in t's logical order
    const int64_t i0 = i % t.ne[0], i1 = (i / t.ne[0]) % t.ne[1], i2 = (i / (t.ne[0] * t.ne[1])) % t.ne[2], i3 = i / (t.ne[0] * t.ne[1] * t.ne[2]);
    return i0 * t.nb[0] + i1 * t.nb[1] + i2 * t.nb[2] + i3 * t.nb[3];
}
// blockIdx.y selects one of two independent copies of the same size (the K and V cache updates of a layer: one launch)
__global__ void cpy_kernel(tdesc s, tdesc d, tdesc s2, tdesc d2, int64_t n) {
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (blockIdx.y) { s = s2; d = d2; }
    const uint8_t * sp = s.data + offset_of(s, i);
    uint8_t * dp = d.data + offset_of(d, i);
    const float v = s.type == T_F32 ? *(const float *)sp : __half2float(*(const __half *)sp);
    if (d.type == T_F32) *(float *)dp = v;
    else if (s.type == T_F16) *(__half *)dp = *(const __half *)sp;
    else *(__half *)dp = __float2half_rn(v);
}
// f32 (strided, dim-0 contiguous) -> Q8_0 / Q4_0 rows (dst rows contiguous blocks); one thread per 32-block
template <int QT> __global__ void cpy_f32_q_kernel(tdesc s, tdesc d, int64_t nblocks) {
    pdl_trigger();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int64_t e = b * 32;                                   // linear element index of the block start
    const float * x = (const float *)(s.data + offset_of(s, e));
    const int64_t bpr = d.ne[0] / 32;                           // dst blocks per row
    const int64_t drow = b / bpr, dblk = b % bpr;
    const int64_t i1 = drow % d.ne[1], i2 = (drow / d.ne[1]) % d.ne[2], i3 = drow / (d.ne[1] * d.ne[2]);
    uint8_t * o = d.data + i1 * d.nb[1] + i2 * d.nb[2] + i3 * d.nb[3] + dblk * (QT == T_Q8_0 ? 34 : 18);
    float amax = 0.0f, vmax = 0.0f;
    for (int i = 0; i < 32; ++i) if (amax < fabsf(x[i])) { amax = fabsf(x[i]); vmax = x[i]; }
    if (QT == T_Q8_0) {
        const float dd = __fdiv_rn(amax, 127.0f), id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
        *(__half *)o = __float2half_rn(dd);
        for (int i = 0; i < 32; ++i) o[2 + i] = (uint8_t)(int8_t)roundf(__fmul_rn(x[i], id));
    } else {
        const float dd = __fdiv_rn(vmax, -8.0f), id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
        *(__half *)o = __float2half_rn(dd);
        for (int i = 0; i < 16; ++i) {
            const int lo = min(15, (int)(int8_t)(int)__fadd_rn(__fmul_rn(x[i], id), 8.5f));
            const int hi = min(15, (int)(int8_t)(int)__fadd_rn(__fmul_rn(x[i + 16], id), 8.5f));
            o[2 + i] = (uint8_t)((lo & 0xFF) | (hi << 4));
        }
    }
}

// ------------------------------------------------------------------ float MUL_MAT (f32 / f16 weights x f32), batched + strided
// one warp per output element; the CPU rounds src1 to f16 when src0 is f16 (vec_dot_type, ggml-cpu.c:262-268)
__global__ void __launch_bounds__(128) mul_mat_f_kernel(tdesc a, tdesc b, tdesc d, int64_t nout) {
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t o = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (o >= nout) return;
    const int64_t m = o % d.ne[0], n = (o / d.ne[0]) % d.ne[1], i12 = (o / (d.ne[0] * d.ne[1])) % d.ne[2], i13 = o / (d.ne[0] * d.ne[1] * d.ne[2]);
    const int64_t i02 = i12 / (b.ne[2] / a.ne[2]), i03 = i13 / (b.ne[3] / a.ne[3]);
    const uint8_t * ar = a.data + m * a.nb[1] + i02 * a.nb[2] + i03 * a.nb[3];
    const uint8_t * br = b.data + n * b.nb[1] + i12 * b.nb[2] + i13 * b.nb[3];
    float acc = 0.0f;
    if (a.type == T_F32) {
        for (int64_t k = lane; k < a.ne[0]; k += 32) acc += *(const float *)(ar + k * a.nb[0]) * *(const float *)(br + k * b.nb[0]);
    } else {
        for (int64_t k = lane; k < a.ne[0]; k += 32)
            acc += __half2float(*(const __half *)(ar + k * a.nb[0])) * __half2float(__float2half_rn(*(const float *)(br + k * b.nb[0])));
    }
    acc = warp_sum_f(acc);
    if (lane == 0) *(float *)(d.data + m * d.nb[0] + n * d.nb[1] + i12 * d.nb[2] + i13 * d.nb[3]) = acc;
}

// ------------------------------------------------------------------ FLASH_ATTN_EXT (f32 Q; f16 / f32 / block-quantized K and V; f16 mask)
// What ggml_compute_forward_flash_attn_ext_f16 computes (src/ggml-cpu/ggml-cpu.c:10805-10990): per (query row, head, batch) an online-softmax
// pass over the KV positions, dst[d, head, q, b] = sum_kv softmax(scale * K.q [softcap] + slope * mask) V.  One CTA per query row and head;
// its four warps take the KV positions round-robin, each lane owns the elements lane, lane + 32, ... of the head dimension (<= 256), K and V
// rows are decoded on the fly (any advertised block format: quantized KV caches), the four partial (max, sum, accumulator) triples are
// merged through shared memory.  Accumulation is f32 (the CPU keeps an f16 accumulator for f16 V; the reference's own gate is NMSE 5e-4).
// Replaces src/ggml-cuda/fattn*.cu for correctness; this is the bandwidth-shaped decode form, not a tensor-core prefill kernel.
struct fa_params {
    tdesc q, k, v, mask, dst;
    float scale, max_bias, softcap, m0, m1;
    uint32_t n_head_log2;
};
__global__ void __launch_bounds__(128) flash_attn_ext_kernel(fa_params p) {
    pdl_trigger();
    __shared__ float s_m[4], s_s[4], s_acc[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t iq1 = blockIdx.x, iq2 = blockIdx.y, iq3 = blockIdx.z;
    const int D = (int)p.q.ne[0];
    const int64_t nkv = p.k.ne[1];
    const int64_t ik2 = iq2 / (p.q.ne[2] / p.k.ne[2]), ik3 = iq3 / (p.q.ne[3] / p.k.ne[3]);
    const int64_t iv2 = iq2 / (p.q.ne[2] / p.v.ne[2]), iv3 = iq3 / (p.q.ne[3] / p.v.ne[3]);
    const float * pq = (const float *)(p.q.data + iq1 * p.q.nb[1] + iq2 * p.q.nb[2] + iq3 * p.q.nb[3]);
    float qv[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int d = lane + 32 * i; qv[i] = d < D ? pq[d] : 0.0f; acc[i] = 0.0f; }
    float slope = 1.0f;
    if (p.max_bias > 0.0f) {
        const uint32_t h = (uint32_t)iq2;
        slope = h < p.n_head_log2 ? powf(p.m0, (float)(h + 1)) : powf(p.m1, (float)(2 * (h - p.n_head_log2) + 1));
    }
    const __half * mp = p.mask.data ? (const __half *)(p.mask.data + iq1 * p.mask.nb[1]) : nullptr;
    float M = -INFINITY, S = 0.0f;
    for (int64_t ic = warp; ic < nkv; ic += 4) {
        const float mv = mp ? slope * __half2float(mp[ic]) : 0.0f;
        if (mv == -INFINITY) continue;                                          // warp-uniform
        const uint8_t * krow = p.k.data + ic * p.k.nb[1] + ik2 * p.k.nb[2] + ik3 * p.k.nb[3];
        float part = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int d = lane + 32 * i; if (d < D) part += qv[i] * load_elem(krow, p.k.type, d); }
        float sc = warp_sum_f(part) * p.scale;
        if (p.softcap != 0.0f) sc = p.softcap * tanhf(sc);
        sc += mv;
        float vs = 1.0f;
        if (sc > M) {
            const float ms = expf(M - sc);                                      // 0 on the first position (M = -inf)
            M = sc;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= ms;
            S = S * ms + 1.0f;
        } else {
            vs = expf(sc - M);
            S += vs;
        }
        const uint8_t * vrow = p.v.data + ic * p.v.nb[1] + iv2 * p.v.nb[2] + iv3 * p.v.nb[3];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int d = lane + 32 * i; if (d < D) acc[i] += vs * load_elem(vrow, p.v.type, d); }
    }
    if (lane == 0) { s_m[warp] = M; s_s[warp] = S; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int d = lane + 32 * i; if (d < D) s_acc[warp][d] = acc[i]; }
    __syncthreads();
    if (warp == 0) {
        const float Mx = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
        float f[4], St = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { f[w] = s_m[w] == -INFINITY ? 0.0f : expf(s_m[w] - Mx); St += s_s[w] * f[w]; }
        const float inv = 1.0f / St;
        float * out = (float *)(p.dst.data + iq2 * p.dst.nb[1] + iq1 * p.dst.nb[2] + iq3 * p.dst.nb[3]);
        for (int d = lane; d < D; d += 32)
            out[d] = (s_acc[0][d] * f[0] + s_acc[1][d] * f[1] + s_acc[2][d] * f[2] + s_acc[3][d] * f[3]) * inv;
    }
}

static inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

} // namespace b200

using namespace b200;

#define REQUIRE(cond, msg) do { if (!(cond)) { set_error("%s: %s", __func__, msg); return GGML_B200_EUNSUPPORTED; } } while (0)

extern "C" {

int ggml_b200_op_get_rows(const ggml_b200_tensor * src0, const ggml_b200_tensor * ids, const ggml_b200_tensor * dst, void * stream) {
    const tdesc s = T(src0), i = T(ids), d = T(dst);
    REQUIRE(d.type == T_F32 && i.type == 26 /* GGML_TYPE_I32 */, "dst must be f32, ids i32");
    REQUIRE(s.type == T_F32 || s.type == T_F16 || type_bytes(s.type) != 0, "unsupported row type");
    REQUIRE(d.nb[0] == 4, "dst rows must be contiguous");
    const int64_t rows = i.ne[0] * i.ne[1] * i.ne[2];
    if (rows == 0 || s.ne[0] == 0) return GGML_B200_OK;
    B200_CUDA_TRY(launch_pdl(get_rows_kernel, dim3((unsigned)rows), dim3(256), (cudaStream_t)stream, s, i, d));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_bin_bcast(int32_t op, const ggml_b200_tensor * src0, const ggml_b200_tensor * src1, const ggml_b200_tensor * dst, void * stream) {
    const tdesc a = T(src0), b = T(src1), d = T(dst);
    REQUIRE(a.type == T_F32 && b.type == T_F32 && d.type == T_F32, "f32 only");
    const int64_t n = nelem(d);
    if (n == 0) return GGML_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = blocks_for(n, 256);
    switch (op) {
        case 0: B200_CUDA_TRY(launch_pdl(bin_bcast_kernel<0>, dim3(g), dim3(256), st, a, b, d, n)); break;
        case 1: B200_CUDA_TRY(launch_pdl(bin_bcast_kernel<1>, dim3(g), dim3(256), st, a, b, d, n)); break;
        case 2: B200_CUDA_TRY(launch_pdl(bin_bcast_kernel<2>, dim3(g), dim3(256), st, a, b, d, n)); break;
        case 3: B200_CUDA_TRY(launch_pdl(bin_bcast_kernel<3>, dim3(g), dim3(256), st, a, b, d, n)); break;
        default: set_error("bin_bcast: bad op %d", op); return GGML_B200_EINVAL;
    }
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_norm(int32_t rms, const ggml_b200_tensor * src, const ggml_b200_tensor * dst, float eps, void * stream) {
    const tdesc s = T(src), d = T(dst);
    REQUIRE(s.type == T_F32 && d.type == T_F32 && s.nb[0] == 4 && d.nb[0] == 4, "f32 rows contiguous along dim 0");
    const int64_t rows = nrows(s);
    if (rows == 0 || s.ne[0] == 0) return GGML_B200_OK;
    const int threads = s.ne[0] >= 1024 ? 256 : s.ne[0] >= 256 ? 128 : 32;
    if (rms) B200_CUDA_TRY(launch_pdl(norm_kernel<true>, dim3((unsigned)rows), dim3(threads), (cudaStream_t)stream, s, d, eps));
    else     B200_CUDA_TRY(launch_pdl(norm_kernel<false>, dim3((unsigned)rows), dim3(threads), (cudaStream_t)stream, s, d, eps));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_norm_affine(int32_t rms, const ggml_b200_tensor * src, const ggml_b200_tensor * dst_norm, const float * gain, const ggml_b200_tensor * dst_mul,
                             const float * bias, const ggml_b200_tensor * dst_add, float eps, void * stream) {
    const tdesc s = T(src), d1 = T(dst_norm), d2 = T(dst_mul), d3 = T(dst_add);
    REQUIRE(s.type == T_F32 && d1.type == T_F32 && d2.type == T_F32 && d3.type == T_F32, "f32 only");
    REQUIRE(s.nb[0] == 4 && d1.nb[0] == 4 && d2.nb[0] == 4 && d3.nb[0] == 4 && gain && bias, "rows contiguous along dim 0");
    const int64_t rows = nrows(s);
    if (rows == 0 || s.ne[0] == 0) return GGML_B200_OK;
    const int threads = s.ne[0] >= 1024 ? 256 : s.ne[0] >= 256 ? 128 : 32;
    if (rms) B200_CUDA_TRY(launch_pdl(norm_affine_kernel<true>, dim3((unsigned)rows), dim3(threads), (cudaStream_t)stream, s, d1, gain, d2, bias, d3, eps));
    else     B200_CUDA_TRY(launch_pdl(norm_affine_kernel<false>, dim3((unsigned)rows), dim3(threads), (cudaStream_t)stream, s, d1, gain, d2, bias, d3, eps));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_scale(const float * src, float * dst, float s, int64_t n, void * stream) {
    if (n <= 0) return GGML_B200_OK;
    B200_CUDA_TRY(launch_pdl(scale_kernel, dim3(blocks_for(n, 256)), dim3(256), (cudaStream_t)stream, src, dst, s, n));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_diag_mask_inf(const float * src, float * dst, int64_t ne0, int64_t ne1, int64_t n, int32_t n_past, void * stream) {
    if (n <= 0) return GGML_B200_OK;
    B200_CUDA_TRY(launch_pdl(diag_mask_inf_kernel, dim3(blocks_for(n, 256)), dim3(256), (cudaStream_t)stream, src, dst, ne0, ne1, n_past, n));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_unary(int32_t uop, const float * src, float * dst, int64_t n, void * stream) {
    if (n <= 0) return GGML_B200_OK;
    if (uop < 0 || uop > U_SQRT) { set_error("unary: bad op %d", uop); return GGML_B200_EINVAL; }
    B200_CUDA_TRY(launch_pdl(unary_kernel, dim3(blocks_for(n, 256)), dim3(256), (cudaStream_t)stream, uop, src, dst, n));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_soft_max(const float * src, const void * mask, int32_t mask_type, float * dst, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
                          float scale, float max_bias, void * stream) {
    return ggml_b200_op_soft_max_diag(src, mask, mask_type, dst, ne0, ne1, ne2, ne3, scale, max_bias, -1, stream);
}

int ggml_b200_op_soft_max_diag(const float * src, const void * mask, int32_t mask_type, float * dst, int64_t ne0, int64_t ne1, int64_t ne2, int64_t ne3,
                               float scale, float max_bias, int32_t diag_n_past, void * stream) {
    const int64_t rows = ne1 * ne2 * ne3;
    if (rows == 0 || ne0 == 0) return GGML_B200_OK;
    const uint32_t n_head = (uint32_t)ne2;
    uint32_t n_head_log2 = 1; while (n_head_log2 * 2 <= n_head) n_head_log2 *= 2;
    const float m0 = powf(2.0f, -(max_bias) / n_head_log2), m1 = powf(2.0f, -(max_bias / 2.0f) / n_head_log2);
    const int threads = ne0 >= 1024 ? 256 : ne0 >= 128 ? 128 : 32;
    B200_CUDA_TRY(launch_pdl(soft_max_kernel, dim3((unsigned)rows), dim3(threads), (cudaStream_t)stream, src, (const uint8_t *)mask, mask_type, dst, ne0, ne1, ne2, scale, max_bias, m0, m1, n_head_log2, (int)diag_n_past));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_cpy(const ggml_b200_tensor * src, const ggml_b200_tensor * dst, void * stream) {
    const tdesc s = T(src), d = T(dst);
    const int64_t n = nelem(s);
    REQUIRE(n == nelem(d), "element counts differ");
    if (n == 0) return GGML_B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if ((s.type == T_F32 || s.type == T_F16) && (d.type == T_F32 || d.type == T_F16)) {
        B200_CUDA_TRY(launch_pdl(cpy_kernel, dim3(blocks_for(n, 256)), dim3(256), st, s, d, s, d, n));
    } else if (s.type == T_F32 && (d.type == T_Q8_0 || d.type == T_Q4_0)) {
        REQUIRE(s.nb[0] == 4 && s.ne[0] % 32 == 0 && d.ne[0] % 32 == 0, "f32 -> q needs dim-0 contiguous rows of whole blocks");
        const int64_t nb = n / 32;
        if (d.type == T_Q8_0) B200_CUDA_TRY(launch_pdl(cpy_f32_q_kernel<T_Q8_0>, dim3(blocks_for(nb, 128)), dim3(128), st, s, d, nb));
        else                  B200_CUDA_TRY(launch_pdl(cpy_f32_q_kernel<T_Q4_0>, dim3(blocks_for(nb, 128)), dim3(128), st, s, d, nb));
    } else {
        set_error("cpy: unsupported type pair %d -> %d", s.type, d.type);
        return GGML_B200_EUNSUPPORTED;
    }
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_cpy2(const ggml_b200_tensor * src_a, const ggml_b200_tensor * dst_a, const ggml_b200_tensor * src_b, const ggml_b200_tensor * dst_b, void * stream) {
    const tdesc s = T(src_a), d = T(dst_a), s2 = T(src_b), d2 = T(dst_b);
    const int64_t n = nelem(s);
    REQUIRE(n == nelem(d) && n == nelem(s2) && n == nelem(d2), "element counts differ");
    auto fl = [](const tdesc & t) { return t.type == T_F32 || t.type == T_F16; };
    if (!(fl(s) && fl(d) && fl(s2) && fl(d2))) { set_error("cpy2: float tensors only"); return GGML_B200_EUNSUPPORTED; }
    if (n == 0) return GGML_B200_OK;
    B200_CUDA_TRY(launch_pdl(cpy_kernel, dim3(blocks_for(n, 256), 2), dim3(256), (cudaStream_t)stream, s, d, s2, d2, n));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_flash_attn_ext(const ggml_b200_tensor * q, const ggml_b200_tensor * k, const ggml_b200_tensor * v, const ggml_b200_tensor * mask,
                                const ggml_b200_tensor * dst, float scale, float max_bias, float logit_softcap, void * stream) {
    fa_params p;
    p.q = T(q); p.k = T(k); p.v = T(v); p.dst = T(dst);
    if (mask) p.mask = T(mask); else { p.mask = tdesc{}; p.mask.data = nullptr; }
    auto decodable = [](int t) { return t == T_F32 || t == T_F16 || type_bytes(t) != 0; };
    REQUIRE(p.q.type == T_F32 && p.dst.type == T_F32 && p.q.nb[0] == 4 && p.dst.nb[0] == 4, "q and dst must be f32 rows");
    REQUIRE(decodable(p.k.type) && decodable(p.v.type), "unsupported K / V type");
    REQUIRE(p.q.ne[0] >= 1 && p.q.ne[0] <= 256 && p.k.ne[0] == p.q.ne[0] && p.v.ne[0] == p.q.ne[0] && p.v.ne[1] == p.k.ne[1], "head size must be <= 256 and agree");
    REQUIRE(p.k.ne[2] > 0 && p.q.ne[2] % p.k.ne[2] == 0 && p.q.ne[3] % p.k.ne[3] == 0 && p.q.ne[2] % p.v.ne[2] == 0 && p.q.ne[3] % p.v.ne[3] == 0, "heads do not broadcast");
    REQUIRE(!mask || (p.mask.type == T_F16 && p.mask.nb[0] == 2 && p.mask.ne[0] >= p.k.ne[1] && p.mask.ne[1] >= p.q.ne[1]), "mask must be f16 [n_kv, >= n_q]");
    if (p.q.ne[1] == 0 || p.q.ne[2] == 0 || p.q.ne[3] == 0) return GGML_B200_OK;
    REQUIRE(p.q.ne[2] <= 65535 && p.q.ne[3] <= 65535, "too many heads / batches for one grid");
    p.scale = scale; p.max_bias = max_bias; p.softcap = logit_softcap;
    if (logit_softcap != 0.0f) p.scale /= logit_softcap;
    const uint32_t n_head = (uint32_t)p.q.ne[2];
    uint32_t n_head_log2 = 1; while (n_head_log2 * 2 <= n_head) n_head_log2 *= 2;
    p.n_head_log2 = n_head_log2;
    p.m0 = powf(2.0f, -(max_bias) / n_head_log2); p.m1 = powf(2.0f, -(max_bias / 2.0f) / n_head_log2);
    B200_CUDA_TRY(launch_pdl(flash_attn_ext_kernel, dim3((unsigned)p.q.ne[1], (unsigned)p.q.ne[2], (unsigned)p.q.ne[3]), dim3(128), (cudaStream_t)stream, p));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int ggml_b200_op_mul_mat_f(const ggml_b200_tensor * src0, const ggml_b200_tensor * src1, const ggml_b200_tensor * dst, void * stream) {
    const tdesc a = T(src0), b = T(src1), d = T(dst);
    REQUIRE((a.type == T_F32 || a.type == T_F16) && b.type == T_F32 && d.type == T_F32, "f32/f16 x f32 -> f32");
    REQUIRE(a.ne[0] == b.ne[0] && d.ne[0] == a.ne[1] && d.ne[1] == b.ne[1] && d.ne[2] == b.ne[2] && d.ne[3] == b.ne[3], "shape mismatch");
    REQUIRE(b.ne[2] % a.ne[2] == 0 && b.ne[3] % a.ne[3] == 0, "batch dims do not broadcast");
    const int64_t nout = nelem(d);
    if (nout == 0) return GGML_B200_OK;
    B200_CUDA_TRY(launch_pdl(mul_mat_f_kernel, dim3(blocks_for(nout, 4)), dim3(128), (cudaStream_t)stream, a, b, d, nout));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

} // extern "C"
