// b200_dequant.cuh — element decoders of the packed block formats: four consecutive weights -> f32, bit-identical to the CPU
// reference dequantize_row_* (src/ggml-quants.c; multiplies and subtracts separately rounded, the reference's ggml-base is built
// without FMA).  Pure per-thread code: also compiled for the host by tests/hostemu (CPU-only logic tests).
#pragma once
#include "b200_quants.cuh"
#include "b200_iq.cuh"

namespace b200 {

__device__ __forceinline__ void k4_scale_min_bytes(const uint8_t * s, int j, int & sc, int & mn) {
    if (j < 4) { sc = s[j] & 63; mn = s[j + 4] & 63; }
    else       { sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4); mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}

// elements [e, e + 4) of the flat tensor `src` (e % 4 == 0; the four never straddle a 16-element group)
template <int T> __device__ __forceinline__ void dequant4(const uint8_t * __restrict__ src, int64_t e, float (&o)[4]) {
    if constexpr (T == T_Q4_0) {
        const uint8_t * b = src + (e / 32) * 18;
        const int j = (int)(e % 32);
        const float d = h2f(load_u16(b));
        const uint8_t * q = b + 2 + (j & 15);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fmul_rn((float)((j < 16 ? (q[i] & 0x0F) : (q[i] >> 4)) - 8), d);
    } else if constexpr (T == T_Q8_0) {
        const uint8_t * b = src + (e / 32) * 34;
        const int j = (int)(e % 32);
        const float d = h2f(load_u16(b));
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fmul_rn((float)(int8_t)b[2 + j + i], d);
    } else if constexpr (T == T_Q4_K || T == T_Q5_K) {
        constexpr int BYTES = fmt<T>::BYTES;
        const uint8_t * b = src + (e / 256) * BYTES;
        const int w = (int)(e % 256), c = w / 64, l = w % 32, hi = (w % 64) / 32;
        const float d = h2f(load_u16(b)), dmin = h2f(load_u16(b + 2));
        int sc, mn;
        k4_scale_min_bytes(b + 4, 2 * c + hi, sc, mn);
        const float dd = __fmul_rn(d, (float)sc), mm = __fmul_rn(dmin, (float)mn);
        const uint8_t * q = b + (T == T_Q5_K ? 48 : 16) + 32 * c + l;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int v = hi ? (q[i] >> 4) : (q[i] & 0x0F);
            if constexpr (T == T_Q5_K) v += ((b[16 + l + i] >> (2 * c + hi)) & 1) << 4;
            o[i] = __fsub_rn(__fmul_rn(dd, (float)v), mm);
        }
    } else if constexpr (T == T_Q6_K) {
        const uint8_t * b = src + (e / 256) * 210;
        const int w = (int)(e % 256), h = w / 128, pos = (w % 128) / 32, l = w % 32;
        const float d = h2f(load_u16(b + 208));
        const int sc = (int)(int8_t)b[192 + 8 * h + l / 16 + 2 * pos];
        const float ds = __fmul_rn(d, (float)sc);
        const uint8_t * ql = b + 64 * h + (pos & 1) * 32 + l, * qh = b + 128 + 32 * h + l;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int lo = pos >= 2 ? (ql[i] >> 4) : (ql[i] & 0x0F);
            const int v = (int)(int8_t)(lo | (((qh[i] >> (2 * pos)) & 3) << 4)) - 32;
            o[i] = __fmul_rn(ds, (float)v);
        }
    } else if constexpr (T == T_IQ2_XXS || T == T_IQ3_XXS || T == T_IQ1_S) {
        iq_dequant4<T>(src, e, o);
    } else if constexpr (T == T_IQ2_XS || T == T_IQ2_S || T == T_IQ3_S || T == T_IQ1_M || T == T_TQ1_0 || T == T_TQ2_0) {
        iq_dequant4_b<T>(src, e, o);
    } else if constexpr (T == T_IQ4_NL) {
        const uint8_t * b = src + (e / 32) * 18;             // dequantize_row_iq4_nl, src/ggml-quants.c:2436-2452
        const int j = (int)(e % 32);
        const float d = h2f(load_u16(b));
        const uint8_t * q = b + 2 + (j & 15);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fmul_rn(d, (float)iq4nl_value(j < 16 ? (q[i] & 0x0F) : (q[i] >> 4)));
    } else if constexpr (T == T_IQ4_XS) {
        const uint8_t * b = src + (e / 256) * 136;           // dequantize_row_iq4_xs: d @0, scales_h @2, scales_l @4, qs @8
        const int w = (int)(e % 256), ib = w >> 5, j = w & 31;
        const uint32_t w0 = (uint32_t)load_u16(b) | ((uint32_t)load_u16(b + 2) << 16), w1 = (uint32_t)load_u16(b + 4) | ((uint32_t)load_u16(b + 6) << 16);
        const float dl = __fmul_rn(h2f(w0 & 0xFFFF), (float)iq4xs_scale(w0, w1, ib));
        const uint8_t * q = b + 8 + 16 * ib + (j & 15);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fmul_rn(dl, (float)iq4nl_value(j < 16 ? (q[i] & 0x0F) : (q[i] >> 4)));
    } else if constexpr (T == T_Q4_1) {
        const uint8_t * b = src + (e / 32) * 20;             // d @0, m @2, qs[16] @4 (dequantize_row_q4_1, src/ggml-quants.c:275-294)
        const int j = (int)(e % 32);
        const float d = h2f(load_u16(b)), m = h2f(load_u16(b + 2));
        const uint8_t * q = b + 4 + (j & 15);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fadd_rn(__fmul_rn((float)(j < 16 ? (q[i] & 0x0F) : (q[i] >> 4)), d), m);
    } else if constexpr (T == T_Q5_0 || T == T_Q5_1) {
        // Q5_0: d @0, qh @2, qs @6 (:296-320)   Q5_1: d @0, m @2, qh @4, qs @8 (:322-348); fifth bit of element j is bit j of qh
        constexpr bool ONE = (T == T_Q5_1);
        const uint8_t * b = src + (e / 32) * (ONE ? 24 : 22);
        const int j = (int)(e % 32);
        const float d = h2f(load_u16(b)), m = ONE ? h2f(load_u16(b + 2)) : 0.0f;
        const uint8_t * hb = b + (ONE ? 4 : 2);
        const uint32_t qh = (uint32_t)hb[0] | ((uint32_t)hb[1] << 8) | ((uint32_t)hb[2] << 16) | ((uint32_t)hb[3] << 24);
        const uint8_t * q = b + (ONE ? 8 : 6) + (j & 15);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int code = (j < 16 ? (q[i] & 0x0F) : (q[i] >> 4)) | (int)(((qh >> (j + i)) & 1u) << 4);
            if constexpr (ONE) o[i] = __fadd_rn(__fmul_rn((float)code, d), m);
            else               o[i] = __fmul_rn((float)(code - 16), d);
        }
    } else if constexpr (T == T_Q2_K || T == T_Q3_K) {
        // element w = 128 h + 32 jj + l of a superblock: 2-bit code in bits 2jj.. of qs[32 h + l], 16-element group 8 h + 2 jj + l / 16
        constexpr bool THREE = (T == T_Q3_K);
        const uint8_t * b = src + (e / 256) * (THREE ? 110 : 84);
        const int w = (int)(e % 256), h = w >> 7, jj = (w >> 5) & 3, l = w & 31, g16 = 8 * h + 2 * jj + (l >> 4);
        if constexpr (!THREE) {                               // scales[16] @0, qs @16, d @80, dmin @82 (dequantize_row_q2_K :712-745)
            const float d = h2f(load_u16(b + 80)), dmin = h2f(load_u16(b + 82));
            const int sc = b[g16];
            const float dl = __fmul_rn(d, (float)(sc & 0x0F)), ml = __fmul_rn(dmin, (float)(sc >> 4));
            const uint8_t * q = b + 16 + 32 * h + l;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __fsub_rn(__fmul_rn(dl, (float)((q[i] >> (2 * jj)) & 3)), ml);
        } else {                                              // hmask[32] @0, qs @32, scales[12] @96, d @108 (dequantize_row_q3_K :1056-1104)
            const uint8_t * sp = b + 96;
            const int lo = g16 < 8 ? (sp[g16] & 0x0F) : (sp[g16 - 8] >> 4);
            const int hi = (sp[8 + (g16 & 3)] >> (2 * (g16 >> 2))) & 3;
            const float dl = __fmul_rn(h2f(load_u16(b + 108)), (float)((lo | (hi << 4)) - 32));
            const uint8_t * q = b + 32 + 32 * h + l, * hm = b + l;
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __fmul_rn(dl, (float)(((q[i] >> (2 * jj)) & 3) - (((hm[i] >> (4 * h + jj)) & 1) ? 0 : 4)));
        }
    }
}

} // namespace b200
