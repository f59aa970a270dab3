/* oracle/quants_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement (written for this repo, not copied) of the reference's algorithm
 * for the hot path GGML_OP_MUL_MAT / GGML_OP_MUL_MAT_ID over block-quantized weights.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the
 * product path (ggml_b200/csrc) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference compiled from /root/reference into oracle/_ref (bit-exact
 * for dequantize / quantize, <= 1e-5 relative for the dot products whose float summation
 * order is unspecified by the reference), and tests/golden/ holds vectors produced by that
 * reference (tests/golden/make_golden.py) for boxes where oracle/_ref cannot be rebuilt.
 *
 * The SURVEY §8f-2 "next" formats (Q4_1, Q5_0, Q5_1, Q2_K, Q3_K, and the Q8_1 activation format the CPU backend pairs with
 * Q4_1/Q5_1) are restated and pinned the same way (same tests, same golden fixtures); their CUDA kernels follow in a later round.
 *
 * Algorithms restated (reference @ 9a4acb37, paths relative to /root/reference):
 *   block layouts ............ src/ggml-common.h:161-166 (q4_0) 203-208 (q8_0) 279-290 (q4_K)
 *                              296-308 (q5_K) 314-320 (q6_K) 323-328 (q8_K)
 *   dequantize_row_q4_0 ...... src/ggml-quants.c:255-273     dequantize_row_q8_0 .. :349-363
 *   dequantize_row_q4_K ...... src/ggml-quants.c:1280-1302   get_scale_min_k4 ..... :631-638
 *   dequantize_row_q5_K ...... src/ggml-quants.c:1482-1507   dequantize_row_q6_K .. :1690-1719
 *   quantize_row_q4_0_ref .... src/ggml-quants.c:31-66       quantize_row_q8_0_ref  :194-217
 *   quantize_row_q8_K_ref .... src/ggml-quants.c:2479-2516   nearest_int .......... :372-377
 *   quantize_row_q8_0 (AVX2) . src/ggml-cpu/ggml-cpu-quants.c:778-835
 *   vec_dot q4_0.q8_0 ........ src/ggml-cpu/ggml-cpu-quants.c:2294-2311 (scalar tail)
 *   vec_dot q8_0.q8_0 ........ src/ggml-cpu/ggml-cpu-quants.c:3335 ff.
 *   vec_dot q4_K.q8_K ........ src/ggml-cpu/ggml-cpu-quants.c:6137-6193 (scalar)
 *   vec_dot q5_K / q6_K ...... src/ggml-cpu/ggml-cpu-quants.c:6196 ff. / 6833 ff.
 *   mul_mat driver ........... src/ggml-cpu/ggml-cpu.c:7428-7605 (quantize src1 rows to vec_dot_type,
 *                              then one vec_dot per (row of src0, column of src1))
 *   mul_mat_id driver ........ src/ggml-cpu/ggml-cpu.c:7609-7784
 *
 * Compiled with -ffp-contract=off: the reference's ggml-base is built without FMA, so
 * `d*q - m` is a rounded multiply followed by a rounded subtract.
 */
#include "quants_oracle.h"
#include "_ref/iq_grids.h"      /* the i-quant codebooks: file-format DATA extracted from the reference's src/ggml-common.h at build time
                                   (scripts/extract_iq_grids.py, oracle/Makefile); not kept in this repository */

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- fp16 */

float oq_fp16_to_fp32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp  = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: renormalise */
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

/* IEEE round-to-nearest-even, the behaviour of F16C / _cvtss_sh(…, 0) used by the reference build */
uint16_t oq_fp32_to_fp16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (x > 0x7F800000u ? 0x200u : 0u));
    if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);          /* rounds to inf */
    if (x < 0x33000001u)  return sign;                                  /* rounds to zero (<= 2^-25) */
    int e = (int)(x >> 23) - 127;
    uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t he;
    if (e < -14) { shift = 13 + (-14 - e); he = 0; }                    /* subnormal half */
    else         { shift = 13;             he = (uint32_t)(e + 15); }
    uint32_t hm = m >> shift;
    const uint32_t rem  = m & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (hm & 1u))) hm++;
    uint32_t out;
    if (he == 0) out = hm;                      /* may carry into exponent 1: correct by construction */
    else         out = ((he - 1) << 10) + hm;   /* hm includes the implicit bit (0x400) */
    return (uint16_t)(sign | out);
}

/* ---------------------------------------------------------------- sizes */

int64_t oq_blck_size(int type) {
    switch (type) {
        case OQ_F32: case OQ_F16: return 1;
        case OQ_Q4_0: case OQ_Q8_0: case OQ_Q4_1: case OQ_Q5_0: case OQ_Q5_1: case OQ_Q8_1: case OQ_IQ4_NL: return 32;
        case OQ_Q4_K: case OQ_Q5_K: case OQ_Q6_K: case OQ_Q8_K: case OQ_Q2_K: case OQ_Q3_K: case OQ_IQ4_XS: return 256;
        case OQ_IQ2_XXS: case OQ_IQ3_XXS: case OQ_IQ1_S: case OQ_IQ2_XS: case OQ_IQ2_S: case OQ_IQ3_S: case OQ_IQ1_M: case OQ_TQ1_0: case OQ_TQ2_0: return 256;
        default: return 0;
    }
}
size_t oq_type_size(int type) {
    switch (type) {
        case OQ_F32: return 4;   case OQ_F16: return 2;
        case OQ_Q4_0: return 18; case OQ_Q8_0: return 34;
        case OQ_Q4_1: return 20; case OQ_Q5_0: return 22; case OQ_Q5_1: return 24; case OQ_Q8_1: return 36;
        case OQ_Q2_K: return 84; case OQ_Q3_K: return 110; case OQ_IQ4_NL: return 18; case OQ_IQ4_XS: return 136;
        case OQ_Q4_K: return 144; case OQ_Q5_K: return 176; case OQ_Q6_K: return 210; case OQ_Q8_K: return 292;
        case OQ_IQ2_XXS: return 66; case OQ_IQ3_XXS: return 98; case OQ_IQ1_S: return 50;
        case OQ_IQ2_XS: return 74; case OQ_IQ2_S: return 82; case OQ_IQ3_S: return 110; case OQ_IQ1_M: return 56; case OQ_TQ1_0: return 54; case OQ_TQ2_0: return 66;
        default: return 0;
    }
}
size_t oq_row_size(int type, int64_t k) { return (size_t)(k / oq_blck_size(type)) * oq_type_size(type); }

static inline uint16_t rd16(const uint8_t * p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline void wr16(uint8_t * p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

/* the 12-byte packing of eight 6-bit (scale, min) pairs used by Q4_K and Q5_K */
static void k4_scale_min(int j, const uint8_t * s, int * sc, int * mn) {
    if (j < 4) { *sc = s[j] & 63;  *mn = s[j + 4] & 63; }
    else       { *sc = (s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4);
                 *mn = (s[j + 4] >> 4)   | ((s[j]     >> 6) << 4); }
}

/* ---------------------------------------------------------------- dequantize */

static void deq_q4_0(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, b += 18, y += 32) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int j = 0; j < 16; ++j) {
            y[j]      = (float)((b[2 + j] & 0x0F) - 8) * d;
            y[j + 16] = (float)((b[2 + j] >> 4)   - 8) * d;
        }
    }
}
static void deq_q8_0(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, b += 34, y += 32) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int j = 0; j < 32; ++j) y[j] = (float)(int8_t)b[2 + j] * d;
    }
}
static void deq_q4_K(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 144) {
        const float d = oq_fp16_to_fp32(rd16(b)), dmin = oq_fp16_to_fp32(rd16(b + 2));
        const uint8_t * sc = b + 4, * q = b + 16;
        for (int c = 0; c < 4; ++c, q += 32) {
            int s0, m0, s1, m1;
            k4_scale_min(2 * c, sc, &s0, &m0); k4_scale_min(2 * c + 1, sc, &s1, &m1);
            const float d0 = d * (float)s0, mm0 = dmin * (float)m0;
            const float d1 = d * (float)s1, mm1 = dmin * (float)m1;
            for (int l = 0; l < 32; ++l) *y++ = d0 * (float)(q[l] & 0x0F) - mm0;
            for (int l = 0; l < 32; ++l) *y++ = d1 * (float)(q[l] >> 4)   - mm1;
        }
    }
}
static void deq_q5_K(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 176) {
        const float d = oq_fp16_to_fp32(rd16(b)), dmin = oq_fp16_to_fp32(rd16(b + 2));
        const uint8_t * sc = b + 4, * qh = b + 16, * q = b + 48;
        for (int c = 0; c < 4; ++c, q += 32) {
            int s0, m0, s1, m1;
            k4_scale_min(2 * c, sc, &s0, &m0); k4_scale_min(2 * c + 1, sc, &s1, &m1);
            const float d0 = d * (float)s0, mm0 = dmin * (float)m0;
            const float d1 = d * (float)s1, mm1 = dmin * (float)m1;
            for (int l = 0; l < 32; ++l) *y++ = d0 * (float)((q[l] & 0x0F) + (((qh[l] >> (2 * c))     & 1) << 4)) - mm0;
            for (int l = 0; l < 32; ++l) *y++ = d1 * (float)((q[l] >> 4)   + (((qh[l] >> (2 * c + 1)) & 1) << 4)) - mm1;
        }
    }
}
static void deq_q6_K(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 210) {
        const float d = oq_fp16_to_fp32(rd16(b + 208));
        for (int h = 0; h < 2; ++h) {
            const uint8_t * ql = b + 64 * h, * qh = b + 128 + 32 * h;
            const int8_t  * sc = (const int8_t *)(b + 192 + 8 * h);
            for (int l = 0; l < 32; ++l) {
                const int g = l / 16;
                const int q1 = (int)((ql[l]      & 0x0F) | (((qh[l] >> 0) & 3) << 4)) - 32;
                const int q2 = (int)((ql[l + 32] & 0x0F) | (((qh[l] >> 2) & 3) << 4)) - 32;
                const int q3 = (int)((ql[l]      >> 4)   | (((qh[l] >> 4) & 3) << 4)) - 32;
                const int q4 = (int)((ql[l + 32] >> 4)   | (((qh[l] >> 6) & 3) << 4)) - 32;
                y[128 * h + l]      = d * (float)sc[g]     * (float)q1;
                y[128 * h + l + 32] = d * (float)sc[g + 2] * (float)q2;
                y[128 * h + l + 64] = d * (float)sc[g + 4] * (float)q3;
                y[128 * h + l + 96] = d * (float)sc[g + 6] * (float)q4;
            }
        }
        y += 256;
    }
}
/* ---- SURVEY §8f-2 "next" formats (oracle first; the CUDA side follows in a later round) ----
 * Q4_1 (src/ggml-common.h:168-180, dequantize_row_q4_1 src/ggml-quants.c:275-294): d @0, m @2, 16 nibble bytes: code * d + m
 * Q5_0 (:182-188, :296-320): d @0, 32 fifth bits @2 (bit j -> element j, bit j+16 -> element j+16), nibbles @6: (code - 16) * d
 * Q5_1 (:190-203, :322-348): d @0, m @2, fifth bits @4, nibbles @8: code * d + m */
static void deq_q4_1(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, b += 20, y += 32) {
        const float d = oq_fp16_to_fp32(rd16(b)), m = oq_fp16_to_fp32(rd16(b + 2));
        for (int j = 0; j < 16; ++j) {
            y[j]      = (float)(b[4 + j] & 0x0F) * d + m;
            y[j + 16] = (float)(b[4 + j] >> 4)   * d + m;
        }
    }
}
static inline uint32_t rd32(const uint8_t * p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline int q5_code(const uint8_t * qs, uint32_t qh, int e) {      /* element e (0..31) of a Q5 block: nibble | fifth bit << 4 */
    const int nib = e < 16 ? (qs[e] & 0x0F) : (qs[e - 16] >> 4);
    return nib | (int)(((qh >> e) & 1u) << 4);
}
static void deq_q5_0(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, b += 22, y += 32) {
        const float d = oq_fp16_to_fp32(rd16(b));
        const uint32_t qh = rd32(b + 2);
        for (int e = 0; e < 32; ++e) y[e] = (float)(q5_code(b + 6, qh, e) - 16) * d;
    }
}
static void deq_q5_1(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, b += 24, y += 32) {
        const float d = oq_fp16_to_fp32(rd16(b)), m = oq_fp16_to_fp32(rd16(b + 2));
        const uint32_t qh = rd32(b + 4);
        for (int e = 0; e < 32; ++e) y[e] = (float)q5_code(b + 8, qh, e) * d + m;
    }
}
/* Q2_K (src/ggml-common.h:247-262, dequantize_row_q2_K src/ggml-quants.c:712-745): 84 bytes = scales[16] (low nibble scale, high
 * nibble min of each 16-element group) @0, qs[64] @16 (2-bit codes), d @80, dmin @82.  Element e = 128 h + 32 j + l (h half,
 * j = 0..3 bit pair, l = 0..31) has its code in bits 2j..2j+1 of qs[32 h + l] and belongs to group 8 h + 2 j + l / 16. */
static inline int q2_code(const uint8_t * qs, int e) { return (qs[32 * (e >> 7) + (e & 31)] >> (2 * ((e >> 5) & 3))) & 3; }
static inline int k_group16(int e) { return 8 * (e >> 7) + 2 * ((e >> 5) & 3) + ((e & 31) >> 4); }
static void deq_q2_K(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 84, y += 256) {
        const float d = oq_fp16_to_fp32(rd16(b + 80)), dmin = oq_fp16_to_fp32(rd16(b + 82));
        for (int e = 0; e < 256; ++e) {
            const uint8_t sc = b[k_group16(e)];
            const float dl = d * (float)(sc & 0x0F), ml = dmin * (float)(sc >> 4);
            y[e] = dl * (float)q2_code(b + 16, e) - ml;
        }
    }
}
/* Q3_K (src/ggml-common.h:264-276, dequantize_row_q3_K src/ggml-quants.c:1056-1104): 110 bytes = hmask[32] @0 (bit 4 h + j of
 * hmask[l] CLEAR means "subtract 4"), qs[64] @32 (low 2 bits, same element order as Q2_K), scales[12] @96 (sixteen 6-bit
 * scales, value - 32), d @108. */
static inline int q3_scale(const uint8_t * s, int g) {                   /* 6-bit scale of 16-element group g, biased by 32 */
    const int lo = g < 8 ? (s[g] & 0x0F) : (s[g - 8] >> 4);
    const int hi = (s[8 + (g & 3)] >> (2 * (g >> 2))) & 3;
    return (lo | (hi << 4)) - 32;
}
static inline int q3_code(const uint8_t * hm, const uint8_t * qs, int e) {
    const int bit = 4 * (e >> 7) + ((e >> 5) & 3);
    return q2_code(qs, e) - (((hm[e & 31] >> bit) & 1) ? 0 : 4);
}
static void deq_q3_K(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 110, y += 256) {
        const float d_all = oq_fp16_to_fp32(rd16(b + 108));
        for (int e = 0; e < 256; ++e) {
            const float dl = d_all * (float)q3_scale(b + 96, k_group16(e));
            y[e] = dl * (float)q3_code(b, b + 32, e);
        }
    }
}
/* IQ4_NL (src/ggml-common.h:398-403, dequantize_row_iq4_nl src/ggml-quants.c:2436-2452): the Q4_0 layout (d @0, 16 nibble bytes)
 * with the nibble indexing a fixed non-linear 16-entry int8 codebook (kvalues_iq4nl, src/ggml-quants.c:2434): value = d * codebook[nibble] */
static const int8_t iq4nl_codebook[16] = { -127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113 };
static void deq_iq4_nl(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, b += 18, y += 32) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int j = 0; j < 16; ++j) {
            y[j]      = d * (float)iq4nl_codebook[b[2 + j] & 0x0F];
            y[j + 16] = d * (float)iq4nl_codebook[b[2 + j] >> 4];
        }
    }
}
/* IQ4_XS (src/ggml-common.h:406-411, dequantize_row_iq4_xs src/ggml-quants.c): 136 bytes = d @0, scales_h @2 (2 high bits of each of
 * the eight 6-bit sub-block scales), scales_l[4] @4 (their low nibbles), qs[128] @8; sub-block ib (32 values, Q4_0 nibble order):
 * value = d * (scale_ib - 32) * codebook[nibble] */
static inline int iq4xs_scale(const uint8_t * b, int ib) {
    const int lo = (b[4 + ib / 2] >> (4 * (ib % 2))) & 0x0F, hi = (rd16(b + 2) >> (2 * ib)) & 3;
    return (lo | (hi << 4)) - 32;
}
static void deq_iq4_xs(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 136) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int ib = 0; ib < 8; ++ib, y += 32) {
            const float dl = d * (float)iq4xs_scale(b, ib);
            const uint8_t * q = b + 8 + 16 * ib;
            for (int j = 0; j < 16; ++j) {
                y[j]      = dl * (float)iq4nl_codebook[q[j] & 0x0F];
                y[j + 16] = dl * (float)iq4nl_codebook[q[j] >> 4];
            }
        }
    }
}
/* ---- i-quants (grid codebooks).  Layouts (src/ggml-common.h): IQ2_XXS 66 B = d, qs[32] u16; IQ3_XXS 98 B = d, qs[64] grid indices,
 * 32 B scales_and_signs; IQ1_S 50 B = d, qs[32], qh[8] u16.  dequantize_row_iq2_xxs / iq3_xxs / iq1_s: src/ggml-quants.c:2197, 2284, 2359. */
static inline float iq_sign(int signs, int j) { return (signs & kmask_iq2xs[j]) ? -1.0f : 1.0f; }
static void deq_iq2_xxs(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 66) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int ib = 0; ib < 8; ++ib) {
            const uint8_t * q = b + 2 + 8 * ib;                 /* 4 grid indices, then 4 x 7 sign bits + 4-bit scale */
            const uint32_t aux = rd32(q + 4);
            const float db = d * (0.5f + (float)(aux >> 28)) * 0.25f;
            for (int l = 0; l < 4; ++l) {
                const uint8_t * grid = (const uint8_t *)(iq2xxs_grid + q[l]);
                const int signs = ksigns_iq2xs[(aux >> (7 * l)) & 127];
                for (int j = 0; j < 8; ++j) *y++ = db * (float)grid[j] * iq_sign(signs, j);
            }
        }
    }
}
static void deq_iq3_xxs(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 98) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int ib = 0; ib < 8; ++ib) {
            const uint8_t * q = b + 2 + 8 * ib;
            const uint32_t aux = rd32(b + 66 + 4 * ib);
            const float db = d * (0.5f + (float)(aux >> 28)) * 0.5f;
            for (int l = 0; l < 4; ++l) {
                const int signs = ksigns_iq2xs[(aux >> (7 * l)) & 127];
                const uint8_t * g1 = (const uint8_t *)(iq3xxs_grid + q[2 * l]), * g2 = (const uint8_t *)(iq3xxs_grid + q[2 * l + 1]);
                for (int j = 0; j < 4; ++j) y[j] = db * (float)g1[j] * iq_sign(signs, j);
                for (int j = 0; j < 4; ++j) y[j + 4] = db * (float)g2[j] * iq_sign(signs, j + 4);
                y += 8;
            }
        }
    }
}
#define OQ_IQ1S_DELTA 0.125f
static void deq_iq1_s(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 50) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int ib = 0; ib < 8; ++ib) {
            const uint32_t qh = rd16(b + 34 + 2 * ib);
            const float dl = d * (float)(2 * ((qh >> 12) & 7) + 1);
            const float delta = (qh & 0x8000) ? -OQ_IQ1S_DELTA : OQ_IQ1S_DELTA;
            for (int l = 0; l < 4; ++l) {
                const int8_t * grid = (const int8_t *)(iq1s_grid + (b[2 + 4 * ib + l] | (((qh >> (3 * l)) & 7) << 8)));
                for (int j = 0; j < 8; ++j) *y++ = dl * ((float)grid[j] + delta);
            }
        }
    }
}

/* ---- IQ2_XS / IQ2_S / IQ3_S / IQ1_M and the ternary TQ1_0 / TQ2_0 (layouts: src/ggml-common.h:226-240, 338-396).  One helper per format yields the
 * eight signed integer codes of group l (0..3) of sub-block ib (0..7); dequantize_row_* (src/ggml-quants.c:2061-2120, 2218-2420) and the
 * generic ggml_vec_dot_*_q8_K branches are stated on top of it. */
static void iq2_xs_codes(const uint8_t * b, int ib, int l, int * c) {
    const uint32_t q = rd16(b + 2 + 8 * ib + 2 * l);
    const uint8_t * grid = (const uint8_t *)(iq2xs_grid + (q & 511));
    const int signs = ksigns_iq2xs[q >> 9];
    for (int j = 0; j < 8; ++j) c[j] = (int)grid[j] * ((signs & kmask_iq2xs[j]) ? -1 : 1);
}
static void iq2_s_codes(const uint8_t * b, int ib, int l, int * c) {
    const uint8_t * grid = (const uint8_t *)(iq2s_grid + (b[2 + 4 * ib + l] | ((b[66 + ib] << (8 - 2 * l)) & 0x300)));
    const int signs = b[34 + 4 * ib + l];
    for (int j = 0; j < 8; ++j) c[j] = (int)grid[j] * ((signs & kmask_iq2xs[j]) ? -1 : 1);
}
static void iq3_s_codes(const uint8_t * b, int ib, int l, int * c) {
    const uint32_t qh = b[66 + ib];
    const uint8_t * g1 = (const uint8_t *)(iq3s_grid + (b[2 + 8 * ib + 2 * l] | ((qh << (8 - 2 * l)) & 256)));
    const uint8_t * g2 = (const uint8_t *)(iq3s_grid + (b[2 + 8 * ib + 2 * l + 1] | ((qh << (7 - 2 * l)) & 256)));
    const int signs = b[74 + 4 * ib + l];
    for (int j = 0; j < 4; ++j) { c[j] = (int)g1[j] * ((signs & kmask_iq2xs[j]) ? -1 : 1); c[j + 4] = (int)g2[j] * ((signs & kmask_iq2xs[j + 4]) ? -1 : 1); }
}
static void iq1_m_codes(const uint8_t * b, int ib, int l, int * c) {
    const uint32_t qh = b[32 + 2 * ib + (l >> 1)];
    const int8_t * grid = (const int8_t *)(iq1s_grid + (b[4 * ib + l] | ((qh << (8 - 4 * (l & 1))) & 0x700)));
    for (int j = 0; j < 8; ++j) c[j] = grid[j];
}
static inline int iq1_m_delta(const uint8_t * b, int ib, int l) { return (b[32 + 2 * ib + (l >> 1)] & (0x08 << (4 * (l & 1)))) ? -1 : 1; }
static inline int iq1_m_ls(const uint8_t * b, int ib, int l) { return 2 * ((rd16(b + 48 + 2 * (ib >> 1)) >> (6 * (ib & 1) + 3 * (l >> 1))) & 7) + 1; }
static inline float iq1_m_d(const uint8_t * b) {
    const uint32_t s0 = rd16(b + 48), s1 = rd16(b + 50), s2 = rd16(b + 52), s3 = rd16(b + 54);
    return oq_fp16_to_fp32((uint16_t)((s0 >> 12) | ((s1 >> 8) & 0x00F0) | ((s2 >> 4) & 0x0F00) | (s3 & 0xF000)));
}
static inline int tq1_code(const uint8_t * b, int e) {                   /* element e of a TQ1_0 block, in { 0, 1, 2 } */
    static const uint8_t pow3[5] = { 1, 3, 9, 27, 81 };
    uint8_t byte; int n;
    if (e < 160) { byte = b[e & 31]; n = e >> 5; }
    else if (e < 240) { byte = b[32 + ((e - 160) & 15)]; n = (e - 160) >> 4; }
    else { byte = b[48 + ((e - 240) & 3)]; n = (e - 240) >> 2; }
    const uint8_t q = (uint8_t)(byte * pow3[n]);
    return (int)(((uint16_t)q * 3) >> 8);
}
static inline int tq2_code(const uint8_t * b, int e) { return (b[32 * (e >> 7) + (e & 31)] >> (2 * ((e >> 5) & 3))) & 3; }

static void deq_iq2_xs_s(int type, const uint8_t * b, float * y, int64_t k) {
    const int bytes = type == OQ_IQ2_XS ? 74 : 82, so = type == OQ_IQ2_XS ? 66 : 74;
    for (int64_t i = 0; i < k / 256; ++i, b += bytes) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int ib = 0; ib < 8; ++ib) {
            const float db[2] = { d * (0.5f + (float)(b[so + ib] & 0xF)) * 0.25f, d * (0.5f + (float)(b[so + ib] >> 4)) * 0.25f };
            for (int l = 0; l < 4; ++l) {
                int c[8];
                if (type == OQ_IQ2_XS) iq2_xs_codes(b, ib, l, c); else iq2_s_codes(b, ib, l, c);
                for (int j = 0; j < 8; ++j) *y++ = db[l / 2] * (float)c[j];          /* = db * grid * (+-1) exactly */
            }
        }
    }
}
static void deq_iq3_s(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 110) {
        const float d = oq_fp16_to_fp32(rd16(b));
        for (int ib = 0; ib < 8; ++ib) {
            const float db = d * (float)(1 + 2 * ((b[106 + (ib >> 1)] >> (4 * (ib & 1))) & 0xF));
            for (int l = 0; l < 4; ++l) { int c[8]; iq3_s_codes(b, ib, l, c); for (int j = 0; j < 8; ++j) *y++ = db * (float)c[j]; }
        }
    }
}
static void deq_iq1_m(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 56) {
        const float d = iq1_m_d(b);
        for (int ib = 0; ib < 8; ++ib)
            for (int l = 0; l < 4; ++l) {
                int c[8]; iq1_m_codes(b, ib, l, c);
                const float dl = d * (float)iq1_m_ls(b, ib, l), delta = iq1_m_delta(b, ib, l) < 0 ? -OQ_IQ1S_DELTA : OQ_IQ1S_DELTA;
                for (int j = 0; j < 8; ++j) *y++ = dl * ((float)c[j] + delta);
            }
    }
}
static void deq_tq(int type, const uint8_t * b, float * y, int64_t k) {
    const int bytes = type == OQ_TQ1_0 ? 54 : 66;
    for (int64_t i = 0; i < k / 256; ++i, b += bytes) {
        const float d = oq_fp16_to_fp32(rd16(b + bytes - 2));
        for (int e = 0; e < 256; ++e) *y++ = (float)((type == OQ_TQ1_0 ? tq1_code(b, e) : tq2_code(b, e)) - 1) * d;
    }
}

static void deq_q8_K(const uint8_t * b, float * y, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, b += 292, y += 256) {
        float d; memcpy(&d, b, 4);
        for (int j = 0; j < 256; ++j) y[j] = d * (float)(int8_t)b[4 + j];
    }
}

int oq_dequantize_row(int type, const void * src, float * dst, int64_t k) {
    const uint8_t * b = (const uint8_t *)src;
    switch (type) {
        case OQ_F32:  memcpy(dst, src, (size_t)k * 4); return 0;
        case OQ_F16:  for (int64_t i = 0; i < k; ++i) dst[i] = oq_fp16_to_fp32(rd16(b + 2 * i)); return 0;
        case OQ_Q4_0: deq_q4_0(b, dst, k); return 0;
        case OQ_Q8_0: deq_q8_0(b, dst, k); return 0;
        case OQ_Q4_K: deq_q4_K(b, dst, k); return 0;
        case OQ_Q5_K: deq_q5_K(b, dst, k); return 0;
        case OQ_Q6_K: deq_q6_K(b, dst, k); return 0;
        case OQ_Q8_K: deq_q8_K(b, dst, k); return 0;
        case OQ_Q4_1: deq_q4_1(b, dst, k); return 0;
        case OQ_Q5_0: deq_q5_0(b, dst, k); return 0;
        case OQ_Q5_1: deq_q5_1(b, dst, k); return 0;
        case OQ_Q2_K: deq_q2_K(b, dst, k); return 0;
        case OQ_Q3_K: deq_q3_K(b, dst, k); return 0;
        case OQ_IQ4_NL: deq_iq4_nl(b, dst, k); return 0;
        case OQ_IQ4_XS: deq_iq4_xs(b, dst, k); return 0;
        case OQ_IQ2_XXS: deq_iq2_xxs(b, dst, k); return 0;
        case OQ_IQ3_XXS: deq_iq3_xxs(b, dst, k); return 0;
        case OQ_IQ1_S: deq_iq1_s(b, dst, k); return 0;
        case OQ_IQ2_XS: case OQ_IQ2_S: deq_iq2_xs_s(type, b, dst, k); return 0;
        case OQ_IQ3_S: deq_iq3_s(b, dst, k); return 0;
        case OQ_IQ1_M: deq_iq1_m(b, dst, k); return 0;
        case OQ_TQ1_0: case OQ_TQ2_0: deq_tq(type, b, dst, k); return 0;
        default: return -1;
    }
}

/* ---------------------------------------------------------------- quantize */

/* round-to-nearest-even through the 1.5*2^23 magic constant (valid for |v| <= 4194303) */
static inline int rne_int(float v) {
    float t = v + 12582912.0f;
    int32_t i; memcpy(&i, &t, 4);
    return (i & 0x007FFFFF) - 0x00400000;
}

static void quant_q4_0(const float * x, uint8_t * b, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, x += 32, b += 18) {
        float amax = 0.0f, vmax = 0.0f;
        for (int j = 0; j < 32; ++j) if (fabsf(x[j]) > amax) { amax = fabsf(x[j]); vmax = x[j]; }
        const float d  = vmax / -8.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        wr16(b, oq_fp32_to_fp16(d));
        for (int j = 0; j < 16; ++j) {
            int lo = (int)(int8_t)(x[j] * id + 8.5f);       /* truncation toward zero, as the C cast does */
            int hi = (int)(int8_t)(x[j + 16] * id + 8.5f);
            if (lo > 15) lo = 15;
            if (hi > 15) hi = 15;
            b[2 + j] = (uint8_t)((lo & 0xFF) | (hi << 4));
        }
    }
}
static void quant_q8_0(const float * x, uint8_t * b, int64_t k) {
    for (int64_t i = 0; i < k / 32; ++i, x += 32, b += 34) {
        float amax = 0.0f;
        for (int j = 0; j < 32; ++j) if (fabsf(x[j]) > amax) amax = fabsf(x[j]);
        const float d  = amax / 127.0f;
        const float id = d != 0.0f ? 1.0f / d : 0.0f;
        wr16(b, oq_fp32_to_fp16(d));
        for (int j = 0; j < 32; ++j) b[2 + j] = (uint8_t)(int8_t)roundf(x[j] * id);   /* ties away from zero */
    }
}
void oq_quantize_row_q8_0_simd(const float * x, void * dst, int64_t k) {
    uint8_t * b = (uint8_t *)dst;
    for (int64_t i = 0; i < k / 32; ++i, x += 32, b += 34) {
        float amax = 0.0f;
        for (int j = 0; j < 32; ++j) if (fabsf(x[j]) > amax) amax = fabsf(x[j]);
        wr16(b, oq_fp32_to_fp16(amax / 127.0f));
        const float id = amax != 0.0f ? 127.0f / amax : 0.0f;
        for (int j = 0; j < 32; ++j) b[2 + j] = (uint8_t)(int8_t)nearbyintf(x[j] * id); /* ties to even (vroundps) */
    }
}
/* Q8_1 as the CPU backend produces it on x86 (quantize_row_q8_1, AVX2 branch, src/ggml-cpu/ggml-cpu-quants.c:1076-1130):
 * 36 bytes = d @0, s @2, 32 int8; d = fp16(amax / 127), codes = rne(x * (127 / amax)), s = fp16(d_unrounded * sum of codes) */
void oq_quantize_row_q8_1_simd(const float * x, void * dst, int64_t k) {
    uint8_t * b = (uint8_t *)dst;
    for (int64_t i = 0; i < k / 32; ++i, x += 32, b += 36) {
        float amax = 0.0f;
        for (int j = 0; j < 32; ++j) if (fabsf(x[j]) > amax) amax = fabsf(x[j]);
        const float d = amax / 127.0f;
        const float id = amax != 0.0f ? 127.0f / amax : 0.0f;
        int sum = 0;
        for (int j = 0; j < 32; ++j) { const int q = (int)nearbyintf(x[j] * id); b[4 + j] = (uint8_t)(int8_t)q; sum += q; }
        wr16(b, oq_fp32_to_fp16(d));
        wr16(b + 2, oq_fp32_to_fp16(d * (float)sum));
    }
}
static void quant_q8_K(const float * x, uint8_t * b, int64_t k) {
    for (int64_t i = 0; i < k / 256; ++i, x += 256, b += 292) {
        float amax = 0.0f, vmax = 0.0f;
        for (int j = 0; j < 256; ++j) if (fabsf(x[j]) > amax) { amax = fabsf(x[j]); vmax = x[j]; }
        if (amax == 0.0f) { memset(b, 0, 292); continue; }   /* bsums are left untouched by the reference; zero here */
        const float iscale = -127.0f / vmax;
        int8_t * q = (int8_t *)(b + 4);
        for (int j = 0; j < 256; ++j) { int v = rne_int(iscale * x[j]); q[j] = (int8_t)(v > 127 ? 127 : v); }
        for (int g = 0; g < 16; ++g) {
            int s = 0;
            for (int j = 0; j < 16; ++j) s += q[16 * g + j];
            wr16(b + 260 + 2 * g, (uint16_t)(int16_t)s);
        }
        const float d = 1.0f / iscale;
        memcpy(b, &d, 4);
    }
}

int oq_quantize_row_ref(int type, const float * src, void * dst, int64_t k) {
    switch (type) {
        case OQ_Q4_0: quant_q4_0(src, (uint8_t *)dst, k); return 0;
        case OQ_Q8_0: quant_q8_0(src, (uint8_t *)dst, k); return 0;
        case OQ_Q8_K: quant_q8_K(src, (uint8_t *)dst, k); return 0;
        default: return -1;
    }
}

/* ---------------------------------------------------------------- dot products */

int oq_vec_dot_type(int type) {
    switch (type) {
        case OQ_Q4_0: case OQ_Q8_0: case OQ_Q5_0: case OQ_IQ4_NL: return OQ_Q8_0;
        case OQ_Q4_1: case OQ_Q5_1: return OQ_Q8_1;
        case OQ_Q4_K: case OQ_Q5_K: case OQ_Q6_K: case OQ_Q2_K: case OQ_Q3_K: case OQ_IQ4_XS: return OQ_Q8_K;
        case OQ_IQ2_XXS: case OQ_IQ3_XXS: case OQ_IQ1_S: case OQ_IQ2_XS: case OQ_IQ2_S: case OQ_IQ3_S: case OQ_IQ1_M: case OQ_TQ1_0: case OQ_TQ2_0: return OQ_Q8_K;
        default: return -1;
    }
}

static float dot_q4_0_q8_0(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 32; ++i, w += 18, y += 34) {
        int s = 0;
        for (int j = 0; j < 16; ++j) {
            s += ((int)(w[2 + j] & 0x0F) - 8) * (int)(int8_t)y[2 + j];
            s += ((int)(w[2 + j] >> 4)   - 8) * (int)(int8_t)y[2 + j + 16];
        }
        acc += (float)s * oq_fp16_to_fp32(rd16(w)) * oq_fp16_to_fp32(rd16(y));
    }
    return acc;
}
static float dot_q8_0_q8_0(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 32; ++i, w += 34, y += 34) {
        int s = 0;
        for (int j = 0; j < 32; ++j) s += (int)(int8_t)w[2 + j] * (int)(int8_t)y[2 + j];
        acc += (float)s * (oq_fp16_to_fp32(rd16(w)) * oq_fp16_to_fp32(rd16(y)));
    }
    return acc;
}
/* Q4_K and Q5_K share everything but the code extraction: sum_j sc_j*(q.y)_j scaled by d*yd, minus dmin*yd*sum_j m_j*bsum_j */
static float dot_q45_K_q8_K(int five, int64_t k, const uint8_t * w, const uint8_t * y) {
    const size_t wb = five ? 176 : 144;
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += wb, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        const float d = oq_fp16_to_fp32(rd16(w)) * yd, dmin = oq_fp16_to_fp32(rd16(w + 2)) * yd;
        const uint8_t * sc = w + 4, * qh = w + 16, * q = w + (five ? 48 : 16);
        int32_t sum_sc = 0, sum_mn = 0;
        for (int c = 0; c < 4; ++c) {
            int s0, m0, s1, m1;
            k4_scale_min(2 * c, sc, &s0, &m0); k4_scale_min(2 * c + 1, sc, &s1, &m1);
            int p0 = 0, p1 = 0, b0 = 0, b1 = 0;
            for (int l = 0; l < 32; ++l) {
                int lo = q[32 * c + l] & 0x0F, hi = q[32 * c + l] >> 4;
                if (five) { lo += ((qh[l] >> (2 * c)) & 1) << 4; hi += ((qh[l] >> (2 * c + 1)) & 1) << 4; }
                p0 += lo * q8[64 * c + l];      b0 += q8[64 * c + l];
                p1 += hi * q8[64 * c + 32 + l]; b1 += q8[64 * c + 32 + l];
            }
            sum_sc += s0 * p0 + s1 * p1;
            sum_mn += m0 * b0 + m1 * b1;
        }
        acc += d * (float)sum_sc - dmin * (float)sum_mn;
    }
    return acc;
}
static float dot_q6_K_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 210, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        const float d = oq_fp16_to_fp32(rd16(w + 208)) * yd;
        int32_t total = 0;
        for (int h = 0; h < 2; ++h) {
            const uint8_t * ql = w + 64 * h, * qh = w + 128 + 32 * h;
            const int8_t  * sc = (const int8_t *)(w + 192 + 8 * h);
            const int8_t  * yy = q8 + 128 * h;
            int part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int l = 0; l < 32; ++l) {
                const int g = l / 16;
                const int q1 = (int)((ql[l]      & 0x0F) | (((qh[l] >> 0) & 3) << 4)) - 32;
                const int q2 = (int)((ql[l + 32] & 0x0F) | (((qh[l] >> 2) & 3) << 4)) - 32;
                const int q3 = (int)((ql[l]      >> 4)   | (((qh[l] >> 4) & 3) << 4)) - 32;
                const int q4 = (int)((ql[l + 32] >> 4)   | (((qh[l] >> 6) & 3) << 4)) - 32;
                part[g]     += q1 * yy[l];
                part[g + 2] += q2 * yy[l + 32];
                part[g + 4] += q3 * yy[l + 64];
                part[g + 6] += q4 * yy[l + 96];
            }
            for (int g = 0; g < 8; ++g) total += (int)sc[g] * part[g];
        }
        acc += d * (float)total;
    }
    return acc;
}

/* ---- "next" formats.  ggml_vec_dot_q4_1_q8_1 / q5_0_q8_0 / q5_1_q8_1 (src/ggml-cpu/ggml-cpu-quants.c:2313, :2606, :2961):
 * per block (d_w * d_y) * sum(code * q) [+ m_w * s_y for the formats with a minimum]; Q8_1 block = d @0, s @2, qs @4 */
static float dot_q4_1_q8_1(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f, mins = 0.0f;
    for (int64_t i = 0; i < k / 32; ++i, w += 20, y += 36) {
        int s = 0;
        for (int j = 0; j < 16; ++j) s += (int)(w[4 + j] & 0x0F) * (int)(int8_t)y[4 + j] + (int)(w[4 + j] >> 4) * (int)(int8_t)y[4 + j + 16];
        acc  += (oq_fp16_to_fp32(rd16(w)) * oq_fp16_to_fp32(rd16(y))) * (float)s;
        mins += oq_fp16_to_fp32(rd16(w + 2)) * oq_fp16_to_fp32(rd16(y + 2));
    }
    return acc + mins;
}
static float dot_q5_0_q8_0(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 32; ++i, w += 22, y += 34) {
        const uint32_t qh = rd32(w + 2);
        int s = 0;
        for (int e = 0; e < 32; ++e) s += (q5_code(w + 6, qh, e) - 16) * (int)(int8_t)y[2 + e];
        acc += (oq_fp16_to_fp32(rd16(w)) * oq_fp16_to_fp32(rd16(y))) * (float)s;
    }
    return acc;
}
static float dot_q5_1_q8_1(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f, mins = 0.0f;
    for (int64_t i = 0; i < k / 32; ++i, w += 24, y += 36) {
        const uint32_t qh = rd32(w + 4);
        int s = 0;
        for (int e = 0; e < 32; ++e) s += q5_code(w + 8, qh, e) * (int)(int8_t)y[4 + e];
        acc  += (oq_fp16_to_fp32(rd16(w)) * oq_fp16_to_fp32(rd16(y))) * (float)s;
        mins += oq_fp16_to_fp32(rd16(w + 2)) * oq_fp16_to_fp32(rd16(y + 2));
    }
    return acc + mins;
}
/* ggml_vec_dot_iq4_nl_q8_0 (src/ggml-cpu/ggml-cpu-quants.c): per block (d_y * d_w) * sum(codebook[nibble] * q) */
static float dot_iq4_nl_q8_0(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 32; ++i, w += 18, y += 34) {
        int s = 0;
        for (int j = 0; j < 16; ++j) s += (int)iq4nl_codebook[w[2 + j] & 0x0F] * (int)(int8_t)y[2 + j] + (int)iq4nl_codebook[w[2 + j] >> 4] * (int)(int8_t)y[2 + j + 16];
        acc += (oq_fp16_to_fp32(rd16(y)) * oq_fp16_to_fp32(rd16(w))) * (float)s;
    }
    return acc;
}
/* ggml_vec_dot_iq4_xs_q8_K: per superblock (d_w * d_y) * sum_ib (scale_ib - 32) * sum(codebook[nibble] * q) */
static float dot_iq4_xs_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 136, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int tot = 0;
        for (int ib = 0; ib < 8; ++ib) {
            const uint8_t * q = w + 8 + 16 * ib;
            int s = 0;
            for (int j = 0; j < 16; ++j) s += (int)iq4nl_codebook[q[j] & 0x0F] * (int)q8[32 * ib + j] + (int)iq4nl_codebook[q[j] >> 4] * (int)q8[32 * ib + j + 16];
            tot += iq4xs_scale(w, ib) * s;
        }
        acc += (oq_fp16_to_fp32(rd16(w)) * yd) * (float)tot;
    }
    return acc;
}
/* ggml_vec_dot_q2_K_q8_K (:4190) / q3_K_q8_K (:4768): per superblock  d_w * d_y * sum_g scale_g * (codes . q)_g  and, for Q2_K,
 * - dmin_w * d_y * sum_g min_g * bsum_g.  Q8_K block = f32 d @0, 256 int8 @4, sixteen int16 bsums @260 */
static float dot_q2_K_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 84, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int part[16] = { 0 }, isum = 0, msum = 0;
        for (int e = 0; e < 256; ++e) part[k_group16(e)] += q2_code(w + 16, e) * (int)q8[e];
        for (int g = 0; g < 16; ++g) {
            isum += (int)(w[g] & 0x0F) * part[g];
            msum += (int)(w[g] >> 4) * (int)(int16_t)rd16(y + 260 + 2 * g);
        }
        const float dall = yd * oq_fp16_to_fp32(rd16(w + 80)), dmin = yd * oq_fp16_to_fp32(rd16(w + 82));
        acc += dall * (float)isum - dmin * (float)msum;
    }
    return acc;
}
static float dot_q3_K_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float acc = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 110, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int part[16] = { 0 }, isum = 0;
        for (int e = 0; e < 256; ++e) part[k_group16(e)] += q3_code(w, w + 32, e) * (int)q8[e];
        for (int g = 0; g < 16; ++g) isum += q3_scale(w + 96, g) * part[g];
        acc += (oq_fp16_to_fp32(rd16(w + 108)) * yd) * (float)isum;
    }
    return acc;
}

/* ggml_vec_dot_iq2_xxs_q8_K / iq3_xxs / iq1_s (src/ggml-cpu/ggml-cpu-quants.c, generic branches): integer sums per 32-value sub-block
 * weighted by the odd sub-block scale 2 s + 1, one f32 multiply-add per superblock, a constant factor at the end */
static float dot_iq2_xxs_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float sumf = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 66, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int bsum = 0;
        for (int ib = 0; ib < 8; ++ib) {
            const uint8_t * q = w + 2 + 8 * ib;
            const uint32_t aux = rd32(q + 4);
            int sumi = 0;
            for (int l = 0; l < 4; ++l) {
                const uint8_t * grid = (const uint8_t *)(iq2xxs_grid + q[l]);
                const int signs = ksigns_iq2xs[(aux >> (7 * l)) & 127];
                for (int j = 0; j < 8; ++j) sumi += (int)grid[j] * (int)q8[32 * ib + 8 * l + j] * ((signs & kmask_iq2xs[j]) ? -1 : 1);
            }
            bsum += sumi * (int)(2 * (aux >> 28) + 1);
        }
        sumf += (oq_fp16_to_fp32(rd16(w)) * yd) * (float)bsum;
    }
    return 0.125f * sumf;
}
static float dot_iq3_xxs_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float sumf = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 98, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int bsum = 0;
        for (int ib = 0; ib < 8; ++ib) {
            const uint8_t * q = w + 2 + 8 * ib;
            const uint32_t aux = rd32(w + 66 + 4 * ib);
            int sumi = 0;
            for (int l = 0; l < 4; ++l) {
                const uint8_t * g1 = (const uint8_t *)(iq3xxs_grid + q[2 * l]), * g2 = (const uint8_t *)(iq3xxs_grid + q[2 * l + 1]);
                const int signs = ksigns_iq2xs[(aux >> (7 * l)) & 127];
                for (int j = 0; j < 4; ++j) {
                    sumi += (int)g1[j] * (int)q8[32 * ib + 8 * l + j]     * ((signs & kmask_iq2xs[j])     ? -1 : 1);
                    sumi += (int)g2[j] * (int)q8[32 * ib + 8 * l + j + 4] * ((signs & kmask_iq2xs[j + 4]) ? -1 : 1);
                }
            }
            bsum += sumi * (int)(2 * (aux >> 28) + 1);
        }
        sumf += (oq_fp16_to_fp32(rd16(w)) * yd) * (float)bsum;
    }
    return 0.25f * sumf;
}
static float dot_iq1_s_q8_K(int64_t k, const uint8_t * w, const uint8_t * y) {
    float sumf = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += 50, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int sumi = 0, sumi1 = 0;
        for (int ib = 0; ib < 8; ++ib) {
            const uint32_t qh = rd16(w + 34 + 2 * ib);
            const int ls = (int)(2 * ((qh >> 12) & 7) + 1), delta = (qh & 0x8000) ? -1 : 1;
            int lsum = 0;
            for (int l = 0; l < 4; ++l) {
                const int8_t * grid = (const int8_t *)(iq1s_grid + (w[2 + 4 * ib + l] | (((qh >> (3 * l)) & 7) << 8)));
                for (int j = 0; j < 8; ++j) lsum += (int)q8[32 * ib + 8 * l + j] * (int)grid[j];
            }
            sumi += ls * lsum;
            sumi1 += ls * delta * ((int)(int16_t)rd16(y + 260 + 4 * ib) + (int)(int16_t)rd16(y + 260 + 4 * ib + 2));
        }
        sumf += oq_fp16_to_fp32(rd16(w)) * yd * ((float)sumi + OQ_IQ1S_DELTA * (float)sumi1);
    }
    return sumf;
}

/* ggml_vec_dot_iq2_xs / iq2_s / iq3_s / iq1_m / tq1_0 / tq2_0 _q8_K, generic branches (src/ggml-cpu/ggml-cpu-quants.c) */
static float dot_iq_more_q8_K(int type, int64_t k, const uint8_t * w, const uint8_t * y) {
    const int bytes = (int)oq_type_size(type);
    float sumf = 0.0f;
    for (int64_t i = 0; i < k / 256; ++i, w += bytes, y += 292) {
        float yd; memcpy(&yd, y, 4);
        const int8_t * q8 = (const int8_t *)(y + 4);
        int bsum = 0, bsum2 = 0;
        for (int ib = 0; ib < 8; ++ib)
            for (int l = 0; l < 4; ++l) {
                int c[8], s = 0, s2 = 0, ls = 1;
                switch (type) {
                    case OQ_IQ2_XS: iq2_xs_codes(w, ib, l, c); ls = 2 * ((w[66 + ib] >> (4 * (l >> 1))) & 0xF) + 1; break;
                    case OQ_IQ2_S:  iq2_s_codes(w, ib, l, c);  ls = 2 * ((w[74 + ib] >> (4 * (l >> 1))) & 0xF) + 1; break;
                    case OQ_IQ3_S:  iq3_s_codes(w, ib, l, c);  ls = 2 * ((w[106 + (ib >> 1)] >> (4 * (ib & 1))) & 0xF) + 1; break;
                    case OQ_IQ1_M:  iq1_m_codes(w, ib, l, c);  ls = iq1_m_ls(w, ib, l); break;
                    case OQ_TQ1_0:  for (int j = 0; j < 8; ++j) c[j] = tq1_code(w, 32 * ib + 8 * l + j) - 1; break;
                    default:        for (int j = 0; j < 8; ++j) c[j] = tq2_code(w, 32 * ib + 8 * l + j) - 1; break;
                }
                for (int j = 0; j < 8; ++j) { s += c[j] * (int)q8[32 * ib + 8 * l + j]; s2 += (int)q8[32 * ib + 8 * l + j]; }
                bsum += ls * s;
                if (type == OQ_IQ1_M) bsum2 += ls * iq1_m_delta(w, ib, l) * s2;
            }
        switch (type) {
            case OQ_IQ1_M: sumf += iq1_m_d(w) * yd * ((float)bsum + OQ_IQ1S_DELTA * (float)bsum2); break;
            case OQ_TQ1_0: case OQ_TQ2_0: sumf += (float)bsum * (oq_fp16_to_fp32(rd16(w + bytes - 2)) * yd); break;
            default: sumf += (oq_fp16_to_fp32(rd16(w)) * yd) * (float)bsum; break;
        }
    }
    return (type == OQ_IQ2_XS || type == OQ_IQ2_S) ? 0.125f * sumf : sumf;
}

float oq_vec_dot(int type, int64_t k, const void * wrow, const void * yq) {
    const uint8_t * w = (const uint8_t *)wrow, * y = (const uint8_t *)yq;
    switch (type) {
        case OQ_Q4_0: return dot_q4_0_q8_0(k, w, y);
        case OQ_Q8_0: return dot_q8_0_q8_0(k, w, y);
        case OQ_Q4_K: return dot_q45_K_q8_K(0, k, w, y);
        case OQ_Q5_K: return dot_q45_K_q8_K(1, k, w, y);
        case OQ_Q6_K: return dot_q6_K_q8_K(k, w, y);
        case OQ_Q4_1: return dot_q4_1_q8_1(k, w, y);
        case OQ_Q5_0: return dot_q5_0_q8_0(k, w, y);
        case OQ_Q5_1: return dot_q5_1_q8_1(k, w, y);
        case OQ_Q2_K: return dot_q2_K_q8_K(k, w, y);
        case OQ_Q3_K: return dot_q3_K_q8_K(k, w, y);
        case OQ_IQ4_NL: return dot_iq4_nl_q8_0(k, w, y);
        case OQ_IQ4_XS: return dot_iq4_xs_q8_K(k, w, y);
        case OQ_IQ2_XXS: return dot_iq2_xxs_q8_K(k, w, y);
        case OQ_IQ3_XXS: return dot_iq3_xxs_q8_K(k, w, y);
        case OQ_IQ1_S: return dot_iq1_s_q8_K(k, w, y);
        case OQ_IQ2_XS: case OQ_IQ2_S: case OQ_IQ3_S: case OQ_IQ1_M: case OQ_TQ1_0: case OQ_TQ2_0: return dot_iq_more_q8_K(type, k, w, y);
        default: return NAN;
    }
}

/* ---------------------------------------------------------------- mat-mul drivers */

static uint8_t * quantize_activations(int vdt, const float * X, int64_t rows, int64_t K, size_t * row_bytes) {
    const size_t rb = oq_row_size(vdt, K);
    uint8_t * q = (uint8_t *)malloc(rb * (size_t)rows + 16);
    if (!q) return NULL;
    #pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        if (vdt == OQ_Q8_0)      oq_quantize_row_q8_0_simd(X + r * K, q + rb * r, K);
        else if (vdt == OQ_Q8_1) oq_quantize_row_q8_1_simd(X + r * K, q + rb * r, K);
        else                     quant_q8_K(X + r * K, q + rb * r, K);
    }
    *row_bytes = rb;
    return q;
}

int oq_mul_mat(int type, const void * W, const float * X, float * Y, int64_t M, int64_t N, int64_t K) {
    const int vdt = oq_vec_dot_type(type);
    if (vdt < 0 || K % oq_blck_size(type) != 0) return -1;
    size_t yrb; uint8_t * yq = quantize_activations(vdt, X, N, K, &yrb);
    if (!yq) return -2;
    const size_t wrb = oq_row_size(type, K);
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n)
            Y[n * M + m] = oq_vec_dot(type, K, (const uint8_t *)W + wrb * m, yq + yrb * n);
    free(yq);
    return 0;
}

int oq_mul_mat_f64(int type, const void * W, const float * X, float * Y, int64_t M, int64_t N, int64_t K) {
    if (K % oq_blck_size(type) != 0) return -1;
    const size_t wrb = oq_row_size(type, K);
    int err = 0;
    #pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        float * wf = (float *)malloc((size_t)K * 4);
        if (!wf || oq_dequantize_row(type, (const uint8_t *)W + wrb * m, wf, K) != 0) { err = 1; free(wf); continue; }
        for (int64_t n = 0; n < N; ++n) {
            double s = 0.0;
            for (int64_t k = 0; k < K; ++k) s += (double)wf[k] * (double)X[n * K + k];
            Y[n * M + m] = (float)s;
        }
        free(wf);
    }
    return err ? -2 : 0;
}

int oq_mul_mat_id(int type, const void * W, const float * X, const int32_t * ids, int64_t ids_stride, float * Y,
                  int64_t M, int64_t K, int64_t n_expert, int64_t n_used, int64_t nb1, int64_t n_tok) {
    const int vdt = oq_vec_dot_type(type);
    if (vdt < 0 || K % oq_blck_size(type) != 0) return -1;
    size_t yrb; uint8_t * yq = quantize_activations(vdt, X, nb1 * n_tok, K, &yrb);
    if (!yq) return -2;
    const size_t wrb = oq_row_size(type, K);
    int bad = 0;
    #pragma omp parallel for schedule(static) collapse(2)
    for (int64_t t = 0; t < n_tok; ++t)
        for (int64_t e = 0; e < n_used; ++e) {
            const int32_t x = ids[t * ids_stride + e];
            if (x < 0 || x >= n_expert) { bad = 1; continue; }
            const uint8_t * w = (const uint8_t *)W + wrb * (size_t)M * (size_t)x;
            const uint8_t * y = yq + yrb * (size_t)(t * nb1 + (e % nb1));
            float * out = Y + (t * n_used + e) * M;
            for (int64_t m = 0; m < M; ++m) out[m] = oq_vec_dot(type, K, w + wrb * m, y);
        }
    free(yq);
    return bad ? -3 : 0;
}
