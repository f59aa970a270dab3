// b200_iq.cuh — the grid-codebook i-quants (IQ2_XXS, IQ3_XXS, IQ1_S): codebooks, 64-weight unit dot products against Q8_K activations and
// element decoders.  Replaces the reference's vec_dot_iq2_xxs_q8_1 / iq3_xxs / iq1_s (src/ggml-cuda/vecdotq.cuh:792, 914, 992) and
// dequantize_block_iq2_xxs / iq3_xxs / iq1_s (src/ggml-cuda/convert.cu); computes what the CPU backend's ggml_vec_dot_iq*_q8_K
// (src/ggml-cpu/ggml-cpu-quants.c, generic branches) and dequantize_row_iq* (src/ggml-quants.c:2197, 2284, 2359) compute.
// The codebooks are file-format data: generated/iq_grids.h is extracted from the reference's src/ggml-common.h when the libraries are
// built (scripts/extract_iq_grids.py), not kept in this repository.
// Block layouts (src/ggml-common.h): IQ2_XXS 66 B = d f16, 8 x { 4 grid indices u8, u32: 4 x 7 sign bits | 4-bit scale << 28 };
// IQ3_XXS 98 B = d, 64 grid indices (8 per 32 weights), 8 x u32 signs | scale; IQ1_S 50 B = d, qs[32], qh[8] u16 (3 high index bits x 4,
// 3-bit scale << 12, delta sign << 15).  Every sub-block of 32 weights = 4 groups of 8 weights from one (IQ2_XXS, IQ1_S) or two (IQ3_XXS)
// codebook entries.
#pragma once
#include "b200_quants.cuh"

#ifdef B200_HOST_EMU
#define IQ_GRID_QUAL static
#else
#define IQ_GRID_QUAL static __device__
#endif
#include "generated/iq_grids.h"

namespace b200 {

// the 8 sign bits of a group from its 7 stored bits: bit 7 = parity (ksigns_iq2xs[i] == i | parity(i) << 7, checked at extraction time)
__device__ __forceinline__ uint32_t iq_signs8(uint32_t s7) { return s7 | (((0x6996u >> ((s7 ^ (s7 >> 4)) & 0xFu)) & 1u) << 7); }
// four magnitudes (the bytes of g, all non-zero in these codebooks) negated where the corresponding low bit of s is set
__device__ __forceinline__ uint32_t iq_apply_signs4(uint32_t g, uint32_t s) {
    const uint32_t b = ((s & 0xFu) * 0x00204081u) & 0x01010101u;     // bit i -> bit 0 of byte i
    const uint32_t m = (b << 8) - b;                                 // 0xFF in the bytes to negate
    return (g ^ m) + b;                                              // two's complement per byte; no carries because no byte is 0
}

template <> __device__ __forceinline__ float unit_dot<T_IQ2_XXS>(const uint8_t * row, int u, const unit_act & A) {
    const uint8_t * sb = row + 66 * (u >> 2);
    uint32_t hd[1], w[4];
    load_words_a2<1>(sb, hd);
    load_words_a2<4>(sb + 2 + 16 * (u & 3), w);                      // sub-blocks 2c, 2c+1: { indices, signs | scale } each
    int tot = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t idx = w[2 * k], aux = w[2 * k + 1];
        int s = 0;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const uint64_t g = iq2xxs_grid[(idx >> (8 * l)) & 0xFF];
            const uint32_t sg = iq_signs8((aux >> (7 * l)) & 127);
            s = dp4a_s((int)iq_apply_signs4((uint32_t)g, sg),              A.q[8 * k + 2 * l],     s);
            s = dp4a_s((int)iq_apply_signs4((uint32_t)(g >> 32), sg >> 4), A.q[8 * k + 2 * l + 1], s);
        }
        tot += s * (int)(2 * (aux >> 28) + 1);
    }
    return (h2f(hd[0] & 0xFFFF) * A.d[0]) * 0.125f * (float)tot;
}

template <> __device__ __forceinline__ float unit_dot<T_IQ3_XXS>(const uint8_t * row, int u, const unit_act & A) {
    const uint8_t * sb = row + 98 * (u >> 2);
    const int c = u & 3;
    uint32_t hd[1], q[4], a[2];
    load_words_a2<1>(sb, hd);
    load_words_a2<4>(sb + 2 + 16 * c, q);                            // 8 grid indices per sub-block
    load_words_a2<2>(sb + 66 + 8 * c, a);
    int tot = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t aux = a[k];
        int s = 0;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const uint32_t pair = q[2 * k + (l >> 1)] >> (16 * (l & 1));    // indices 2l, 2l+1 of the sub-block
            const uint32_t sg = iq_signs8((aux >> (7 * l)) & 127);
            s = dp4a_s((int)iq_apply_signs4(iq3xxs_grid[pair & 0xFF], sg),             A.q[8 * k + 2 * l],     s);
            s = dp4a_s((int)iq_apply_signs4(iq3xxs_grid[(pair >> 8) & 0xFF], sg >> 4), A.q[8 * k + 2 * l + 1], s);
        }
        tot += s * (int)(2 * (aux >> 28) + 1);
    }
    return (h2f(hd[0] & 0xFFFF) * A.d[0]) * 0.25f * (float)tot;
}

template <> __device__ __forceinline__ float unit_dot<T_IQ1_S>(const uint8_t * row, int u, const unit_act & A) {
    const uint8_t * sb = row + 50 * (u >> 2);
    const int c = u & 3;
    uint32_t hd[1], qs[2], qhw[1];
    load_words_a2<1>(sb, hd);
    load_words_a2<2>(sb + 2 + 8 * c, qs);
    load_words_a2<1>(sb + 34 + 4 * c, qhw);
    int sumi = 0, sumi1 = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t qh = (qhw[0] >> (16 * k)) & 0xFFFF;
        const int ls = (int)(2 * ((qh >> 12) & 7) + 1), delta = (qh & 0x8000) ? -1 : 1;
        int s = 0;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const uint64_t g = iq1s_grid[((qs[k] >> (8 * l)) & 0xFF) | (((qh >> (3 * l)) & 7) << 8)];      // eight int8 in { -1, 0, 1 }
            s = dp4a_s((int)(uint32_t)g,         A.q[8 * k + 2 * l],     s);
            s = dp4a_s((int)(uint32_t)(g >> 32), A.q[8 * k + 2 * l + 1], s);
        }
        sumi += ls * s;
        sumi1 += ls * delta * (A.bs[2 * k] + A.bs[2 * k + 1]);
    }
    return h2f(hd[0] & 0xFFFF) * A.d[0] * ((float)sumi + 0.125f * (float)sumi1);
}

// elements [e, e + 4) of the flat tensor (e % 4 == 0: inside one group of 8 weights), bit-identical to dequantize_row_iq*
template <int T> __device__ __forceinline__ void iq_dequant4(const uint8_t * __restrict__ src, int64_t e, float (&o)[4]) {
    const int w = (int)(e % 256), ib = w >> 5, l = (w >> 3) & 3, j0 = w & 7;
    if constexpr (T == T_IQ2_XXS) {
        const uint8_t * b = src + (e / 256) * 66, * q = b + 2 + 8 * ib;
        const uint32_t aux = (uint32_t)load_u16(q + 4) | ((uint32_t)load_u16(q + 6) << 16);
        const float db = __fmul_rn(__fmul_rn(h2f(load_u16(b)), 0.5f + (float)(aux >> 28)), 0.25f);
        const uint32_t g = (uint32_t)(iq2xxs_grid[q[l]] >> (8 * j0)), sg = iq_signs8((aux >> (7 * l)) & 127) >> j0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float v = __fmul_rn(db, (float)((g >> (8 * i)) & 0xFF)); o[i] = ((sg >> i) & 1) ? -v : v; }
    } else if constexpr (T == T_IQ3_XXS) {
        const uint8_t * b = src + (e / 256) * 98, * q = b + 2 + 8 * ib, * ap = b + 66 + 4 * ib;
        const uint32_t aux = (uint32_t)load_u16(ap) | ((uint32_t)load_u16(ap + 2) << 16);
        const float db = __fmul_rn(__fmul_rn(h2f(load_u16(b)), 0.5f + (float)(aux >> 28)), 0.5f);
        const uint32_t g = iq3xxs_grid[q[2 * l + (j0 >> 2)]], sg = iq_signs8((aux >> (7 * l)) & 127) >> j0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float v = __fmul_rn(db, (float)((g >> (8 * i)) & 0xFF)); o[i] = ((sg >> i) & 1) ? -v : v; }
    } else {   // IQ1_S
        const uint8_t * b = src + (e / 256) * 50;
        const uint32_t qh = load_u16(b + 34 + 2 * ib);
        const float dl = __fmul_rn(h2f(load_u16(b)), (float)(2 * ((qh >> 12) & 7) + 1));
        const float delta = (qh & 0x8000) ? -0.125f : 0.125f;
        const uint32_t g = (uint32_t)(iq1s_grid[b[2 + 4 * ib + l] | (((qh >> (3 * l)) & 7) << 8)] >> (8 * j0));
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fmul_rn(dl, __fadd_rn((float)(int8_t)((g >> (8 * i)) & 0xFF), delta));
    }
}

// ----------------------------------------------------------------------------- the remaining i-quants and the ternary formats
// IQ2_XS 74 B = d, qs u16[32] (9-bit grid index | 7 sign bits << 9), scales[8] (two 4-bit scales per 32 weights);
// IQ2_S 82 B = d, qs[32] grid low bytes, signs[32], qh[8] (2 high index bits x 4), scales[8];
// IQ3_S 110 B = d, qs[64], qh[8] (1 high index bit x 8), signs[32], scales[4] (one 4-bit scale per 32 weights);
// IQ1_M 56 B = qs[32], qh[16] (3 high index bits + delta sign per 8 weights), scales[8] (3-bit scale per 16 weights; the f16 super-scale is
// spread over the top nibbles of the four u16); TQ1_0 54 B = qs[48] (5 trits per byte), qh[4] (4 trits per byte), d; TQ2_0 66 B = qs[64]
// (2 bits per weight), d.  (src/ggml-common.h:226-240, 338-396; dequantize_row_*: src/ggml-quants.c:2218-2420, 2061-2120)
// One decoder per format serves the dot products and the element decoders: the eight SIGNED integer codes of group l (0..3) of
// sub-block ib (0..7), packed as two int8x4 words, plus the group's integer and float scales.
__device__ __forceinline__ uint32_t iq_rd16(const uint8_t * p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
__device__ __forceinline__ int tq1_trit(uint32_t byte, int n) {                 // trit n of a base-3 packed byte: ((byte * 3^n mod 256) * 3) >> 8
    const uint32_t pw = n == 0 ? 1u : n == 1 ? 3u : n == 2 ? 9u : n == 3 ? 27u : 81u;
    return (int)((((byte * pw) & 0xFFu) * 3u) >> 8);
}
// element e (0..255) of a TQ1_0 block -> code in { 0, 1, 2 }
__device__ __forceinline__ int tq1_code(const uint8_t * sb, int e) {
    if (e < 160) return tq1_trit(sb[e & 31], e >> 5);                            // bytes 0..31: element n * 32 + m
    if (e < 240) { const int f = e - 160; return tq1_trit(sb[32 + (f & 15)], f >> 4); }   // bytes 32..47: element 160 + n * 16 + m
    const int f = e - 240;                                                      // qh: element 240 + n * 4 + j
    return tq1_trit(sb[48 + (f & 3)], f >> 2);
}
template <int T> __device__ __forceinline__ void iq_group8(const uint8_t * sb, int ib, int l, uint32_t & lo, uint32_t & hi) {
    if constexpr (T == T_IQ2_XS) {
        const uint32_t q = iq_rd16(sb + 2 + 8 * ib + 2 * l);
        const uint64_t g = iq2xs_grid[q & 511];
        const uint32_t sg = iq_signs8(q >> 9);
        lo = iq_apply_signs4((uint32_t)g, sg); hi = iq_apply_signs4((uint32_t)(g >> 32), sg >> 4);
    } else if constexpr (T == T_IQ2_S) {
        const uint32_t idx = (uint32_t)sb[2 + 4 * ib + l] | ((((uint32_t)sb[66 + ib]) << (8 - 2 * l)) & 0x300u);
        const uint64_t g = iq2s_grid[idx];
        const uint32_t sg = sb[34 + 4 * ib + l];
        lo = iq_apply_signs4((uint32_t)g, sg); hi = iq_apply_signs4((uint32_t)(g >> 32), sg >> 4);
    } else if constexpr (T == T_IQ3_S) {
        const uint32_t qh = sb[66 + ib];
        const uint32_t i1 = (uint32_t)sb[2 + 8 * ib + 2 * l] | ((qh << (8 - 2 * l)) & 256u), i2 = (uint32_t)sb[2 + 8 * ib + 2 * l + 1] | ((qh << (7 - 2 * l)) & 256u);
        const uint32_t sg = sb[74 + 4 * ib + l];
        lo = iq_apply_signs4(iq3s_grid[i1], sg); hi = iq_apply_signs4(iq3s_grid[i2], sg >> 4);
    } else if constexpr (T == T_IQ1_M) {
        const uint32_t qh = sb[32 + 2 * ib + (l >> 1)];
        const uint64_t g = iq1s_grid[(uint32_t)sb[4 * ib + l] | ((qh << (8 - 4 * (l & 1))) & 0x700u)];
        lo = (uint32_t)g; hi = (uint32_t)(g >> 32);
    } else if constexpr (T == T_TQ1_0) {
        lo = 0; hi = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lo |= (uint32_t)((tq1_code(sb, 32 * ib + 8 * l + j) - 1) & 0xFF) << (8 * j);
            hi |= (uint32_t)((tq1_code(sb, 32 * ib + 8 * l + 4 + j) - 1) & 0xFF) << (8 * j);
        }
    } else {   // TQ2_0: element 128 a + 32 n + m  <->  bits 2n.. of byte 32 a + m
        lo = 0; hi = 0;
        const int e0 = 32 * ib + 8 * l, a = e0 >> 7, n = (e0 >> 5) & 3, m0 = e0 & 31;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            lo |= (uint32_t)(((int)((sb[32 * a + m0 + j] >> (2 * n)) & 3) - 1) & 0xFF) << (8 * j);
            hi |= (uint32_t)(((int)((sb[32 * a + m0 + 4 + j] >> (2 * n)) & 3) - 1) & 0xFF) << (8 * j);
        }
    }
}
// integer scale of group (ib, l): the 2 s + 1 of the 16-value half (IQ2_XS, IQ2_S, IQ1_M) or of the sub-block (IQ3_S); 1 for the ternary formats
template <int T> __device__ __forceinline__ int iq_ls(const uint8_t * sb, int ib, int l) {
    if constexpr (T == T_IQ2_XS) return 2 * (int)((sb[66 + ib] >> (4 * (l >> 1))) & 0xF) + 1;
    else if constexpr (T == T_IQ2_S) return 2 * (int)((sb[74 + ib] >> (4 * (l >> 1))) & 0xF) + 1;
    else if constexpr (T == T_IQ3_S) return 2 * (int)((sb[106 + (ib >> 1)] >> (4 * (ib & 1))) & 0xF) + 1;
    else if constexpr (T == T_IQ1_M) return 2 * (int)((iq_rd16(sb + 48 + 2 * (ib >> 1)) >> (6 * (ib & 1) + 3 * (l >> 1))) & 7) + 1;
    else return 1;
}
// the block's f16 super-scale (bits)
template <int T> __device__ __forceinline__ uint32_t iq_dbits(const uint8_t * sb) {
    if constexpr (T == T_IQ1_M) {
        const uint32_t s0 = iq_rd16(sb + 48), s1 = iq_rd16(sb + 50), s2 = iq_rd16(sb + 52), s3 = iq_rd16(sb + 54);
        return (s0 >> 12) | ((s1 >> 8) & 0x00F0u) | ((s2 >> 4) & 0x0F00u) | (s3 & 0xF000u);
    } else if constexpr (T == T_TQ1_0) return iq_rd16(sb + 52);
    else if constexpr (T == T_TQ2_0) return iq_rd16(sb + 64);
    else return iq_rd16(sb);
}
// IQ1_M: -1 where the delta of group (ib, l) is negative
__device__ __forceinline__ int iq1m_delta_sign(const uint8_t * sb, int ib, int l) { return (sb[32 + 2 * ib + (l >> 1)] & (0x08u << (4 * (l & 1)))) ? -1 : 1; }

template <int T> __device__ __forceinline__ float iq_unit_dot(const uint8_t * row, int u, const unit_act & A) {
    const uint8_t * sb = row + fmt<T>::BYTES * (u >> 2);
    const int c = u & 3;
    int tot = 0, tot2 = 0;
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) {
        const int ib = 2 * c + (g8 >> 2), l = g8 & 3;
        uint32_t lo, hi;
        iq_group8<T>(sb, ib, l, lo, hi);
        const int ls = iq_ls<T>(sb, ib, l);
        int s = dp4a_s((int)lo, A.q[2 * g8], 0);
        s = dp4a_s((int)hi, A.q[2 * g8 + 1], s);
        tot += ls * s;
        if constexpr (T == T_IQ1_M) {
            const int s2 = dp4a_s(0x01010101, A.q[2 * g8], dp4a_s(0x01010101, A.q[2 * g8 + 1], 0));
            tot2 += ls * iq1m_delta_sign(sb, ib, l) * s2;
        }
    }
    const float d = h2f(iq_dbits<T>(sb));
    if constexpr (T == T_IQ2_XS || T == T_IQ2_S) return (d * A.d[0]) * 0.125f * (float)tot;
    else if constexpr (T == T_IQ1_M) return d * A.d[0] * ((float)tot + 0.125f * (float)tot2);
    else if constexpr (T == T_TQ1_0 || T == T_TQ2_0) return (float)tot * (d * A.d[0]);
    else return (d * A.d[0]) * (float)tot;                               // IQ3_S
}
template <> __device__ __forceinline__ float unit_dot<T_IQ2_XS>(const uint8_t * row, int u, const unit_act & A) { return iq_unit_dot<T_IQ2_XS>(row, u, A); }
template <> __device__ __forceinline__ float unit_dot<T_IQ2_S>(const uint8_t * row, int u, const unit_act & A)  { return iq_unit_dot<T_IQ2_S>(row, u, A); }
template <> __device__ __forceinline__ float unit_dot<T_IQ3_S>(const uint8_t * row, int u, const unit_act & A)  { return iq_unit_dot<T_IQ3_S>(row, u, A); }
template <> __device__ __forceinline__ float unit_dot<T_IQ1_M>(const uint8_t * row, int u, const unit_act & A)  { return iq_unit_dot<T_IQ1_M>(row, u, A); }
template <> __device__ __forceinline__ float unit_dot<T_TQ1_0>(const uint8_t * row, int u, const unit_act & A)  { return iq_unit_dot<T_TQ1_0>(row, u, A); }
template <> __device__ __forceinline__ float unit_dot<T_TQ2_0>(const uint8_t * row, int u, const unit_act & A)  { return iq_unit_dot<T_TQ2_0>(row, u, A); }

// elements [e, e + 4): bit-identical to dequantize_row_iq2_xs / iq2_s / iq3_s / iq1_m / tq1_0 / tq2_0
template <int T> __device__ __forceinline__ void iq_dequant4_b(const uint8_t * __restrict__ src, int64_t e, float (&o)[4]) {
    const uint8_t * sb = src + (e / 256) * fmt<T>::BYTES;
    const int w = (int)(e % 256), ib = w >> 5, l = (w >> 3) & 3, j0 = w & 7;
    uint32_t lo, hi;
    iq_group8<T>(sb, ib, l, lo, hi);
    const uint32_t codes = j0 ? hi : lo;
    const float d = h2f(iq_dbits<T>(sb));
    float sc, delta = 0.0f;
    if constexpr (T == T_IQ2_XS || T == T_IQ2_S) sc = __fmul_rn(__fmul_rn(d, 0.5f + (float)((iq_ls<T>(sb, ib, l) - 1) >> 1)), 0.25f);
    else if constexpr (T == T_IQ3_S || T == T_IQ1_M) sc = __fmul_rn(d, (float)iq_ls<T>(sb, ib, l));
    else sc = d;
    if constexpr (T == T_IQ1_M) delta = iq1m_delta_sign(sb, ib, l) < 0 ? -0.125f : 0.125f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float c = (float)(int8_t)((codes >> (8 * i)) & 0xFF);
        if constexpr (T == T_IQ1_M) o[i] = __fmul_rn(sc, __fadd_rn(c, delta));
        else if constexpr (T == T_TQ1_0 || T == T_TQ2_0) o[i] = __fmul_rn(c, sc);
        else o[i] = __fmul_rn(sc, c);
    }
}

} // namespace b200
