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

} // namespace b200
