/* oracle/quants_oracle.h — TEST INFRASTRUCTURE ONLY (see quants_oracle.c). */
#ifndef QUANTS_ORACLE_H
#define QUANTS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* type ids = enum ggml_type (reference include/ggml.h:351-390) */
enum {
    OQ_F32 = 0, OQ_F16 = 1, OQ_Q4_0 = 2, OQ_Q4_1 = 3, OQ_Q5_0 = 6, OQ_Q5_1 = 7, OQ_Q8_0 = 8, OQ_Q8_1 = 9,
    OQ_Q2_K = 10, OQ_Q3_K = 11, OQ_Q4_K = 12, OQ_Q5_K = 13, OQ_Q6_K = 14, OQ_Q8_K = 15, OQ_IQ4_NL = 20, OQ_IQ4_XS = 23, OQ_IQ2_XXS = 16, OQ_IQ3_XXS = 18, OQ_IQ1_S = 19, OQ_IQ2_XS = 17, OQ_IQ3_S = 21, OQ_IQ2_S = 22, OQ_IQ1_M = 29, OQ_TQ1_0 = 34, OQ_TQ2_0 = 35,
};

float    oq_fp16_to_fp32(uint16_t h);
uint16_t oq_fp32_to_fp16(float f);

int64_t oq_blck_size(int type);
size_t  oq_type_size(int type);
size_t  oq_row_size(int type, int64_t k);

/* weights -> f32 (dequantize_row_*) */
int  oq_dequantize_row(int type, const void * src, float * dst, int64_t k);
/* f32 -> blocks: Q4_0/Q8_0 (quantize_row_*_ref), Q8_K */
int  oq_quantize_row_ref(int type, const float * src, void * dst, int64_t k);
/* activation quantizers as executed by the CPU backend on x86 (AVX2 flavours of quantize_row_q8_0 / quantize_row_q8_1) */
void oq_quantize_row_q8_0_simd(const float * src, void * dst, int64_t k);
void oq_quantize_row_q8_1_simd(const float * src, void * dst, int64_t k);
/* which activation format the CPU backend pairs with a weight type (type_traits_cpu[].vec_dot_type) */
int  oq_vec_dot_type(int type);
/* integer block dot of one weight row with one quantized activation row */
float oq_vec_dot(int type, int64_t k, const void * wrow, const void * yq);

/* Y[n*M + m] = sum_k W[m][k] * X[n][k], W in `type` blocks, contiguous; activation quantized as the CPU backend does */
int oq_mul_mat(int type, const void * W, const float * X, float * Y, int64_t M, int64_t N, int64_t K);
/* same without quantizing activations: dequant(W) . X in double (the "ideal" answer; used for error budgets) */
int oq_mul_mat_f64(int type, const void * W, const float * X, float * Y, int64_t M, int64_t N, int64_t K);
/* MUL_MAT_ID: as[K,M,n_expert], b[K,nb1,n_tok], ids[n_used (row stride ids_stride), n_tok] -> c[M,n_used,n_tok] */
int oq_mul_mat_id(int type, const void * W, const float * X, const int32_t * ids, int64_t ids_stride, float * Y,
                  int64_t M, int64_t K, int64_t n_expert, int64_t n_used, int64_t nb1, int64_t n_tok);

#ifdef __cplusplus
}
#endif
#endif
