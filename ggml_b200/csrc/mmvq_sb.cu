// mmvq_sb.cu — the bandwidth-path quantized mat-vec (n = 1), second generation: "one lane per 256-weight task".
//
// Why (profiles/r01_gemv_q4k_v1.md): the first TMA kernel moved exactly the algorithmic bytes from DRAM but spent
// 5.0 M warp-instructions on a 45 M-weight matrix (6-bit scale decode repeated per 64 weights with a run-time
// sub-block index, activation quantization repeated by 444 small CTAs on their critical path, 12 warps per SM),
// i.e. it was issue/latency-bound at 30 % of the HBM roofline.  Here:
//   * a lane owns a whole TASK = 256 consecutive weights of one row (a K-quant superblock, or 8 Q4_0 / 4 Q8_0
//     blocks): the 12-byte scale pack is decoded once, every sub-block index is a compile-time constant, the high
//     nibbles are used in place (u8 dp4a of q & 0xF0 = 16 x the nibble dot) -> ~0.8 instructions per weight;
//   * LPR lanes cooperate on a row (tasks strided by LPR), so a row's dot product is finished by 4-5 shuffles inside
//     a (half-)warp and written straight to y: no cross-warp reduction, no block-wide barrier per stage;
//   * CTAs are fat (8 consumer warps + 1 producer warp, two CTAs per SM): the activation vector is quantized once
//     per CTA by all warps in parallel while the first TMA stages are in flight, into a task-interleaved layout
//     that makes every LDS.128 of it bank-conflict-free;
//   * a dedicated producer warp keeps a ring of TMA bulk copies (cp.async.bulk + mbarrier complete_tx) in flight;
//     consumers release stages through per-stage "empty" mbarriers; chunks after the first are handed out by an
//     atomic counter (self-resetting), so SMs stay balanced to one chunk.
// Weights are read once from HBM in the reference's packed layout.  Numerics are those of b200_quants.cuh
// (int8 activations quantized as ggml-cpu does, integer dots, f32 scaling); only the f32 summation order differs.
#include "b200_internal.h"
#include "b200_quants.cuh"

#include <atomic>
#include <cstdlib>

namespace b200 {

// ----------------------------------------------------------------------------- PTX helpers (as mmvq.cu)
__device__ __forceinline__ uint32_t sb_smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sb_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void sb_mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(uint64_t * bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SB_DONE;\n"
        "bra SB_WAIT;\n"
        "SB_DONE:\n"
        "}\n" ::"r"(sb_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void sb_tma_g2s(void * dst_smem, const void * src_gmem, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sb_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(sb_smem_u32(bar)) : "memory");
}
// programmatic dependent launch: let the next kernel's prologue start / wait for the previous kernel's results
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// mixed-sign dp4a: bytes of a are unsigned, bytes of b signed
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ----------------------------------------------------------------------------- task geometry
// TASK_W weights per task, TASK_B bytes; LPR lanes per row.
template <int T> struct sbfmt;
template <> struct sbfmt<T_Q4_K> { static constexpr int TASK_W = 256, TASK_B = 144, LPR = 16, KQ = 1; };
template <> struct sbfmt<T_Q5_K> { static constexpr int TASK_W = 256, TASK_B = 176, LPR = 16, KQ = 1; };
template <> struct sbfmt<T_Q6_K> { static constexpr int TASK_W = 256, TASK_B = 210, LPR = 16, KQ = 1; };
template <> struct sbfmt<T_Q4_0> { static constexpr int TASK_W = 256, TASK_B = 144, LPR = 16, KQ = 0; };
template <> struct sbfmt<T_Q8_0> { static constexpr int TASK_W = 128, TASK_B = 136, LPR = 32, KQ = 0; };

// Task-interleaved activation record in shared memory (ntask = K / 256 "act tasks" of 256 values each):
//   q   : chunk j (16 int8 = values 16j..16j+15 of act-task t) at (j * ntask + t) * 16,           j = 0..15
//   s32 : eight int32 sums of 32 values of act-task t:  two 16-byte chunks at off_s32 + (jj * ntask + t) * 16
//   s16 : sixteen int16 sums of 16 values:              two chunks at off_s16 + (jj * ntask + t) * 16
//   d   : Q8_K family: one float per act-task at off_d + 4 t;  Q8_0 family: eight floats, chunks at off_d + (jj * ntask + t) * 16
// consecutive lanes (tasks) read consecutive 16-byte chunks -> conflict-free LDS.128, lanes on the other row broadcast.
struct sb_act {
    int32_t ntask, off_s32, off_s16, off_d, bytes;
};
__host__ __device__ inline sb_act make_sb_act(int64_t K) {
    sb_act A;
    A.ntask = (int32_t)(K / 256);
    A.off_s32 = 16 * A.ntask * 16;
    A.off_s16 = A.off_s32 + 2 * A.ntask * 16;
    A.off_d   = A.off_s16 + 2 * A.ntask * 16;
    A.bytes   = A.off_d + 2 * A.ntask * 16;
    return A;
}

// one warp quantizes act-task t (256 values) into the interleaved record
template <bool KQ> __device__ __forceinline__ void sb_quantize_task(const float4 a, const float4 b, uint8_t * rec, const sb_act & A, int t) {
    const int lane = threadIdx.x & 31;
    const float v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
    int q[8];
    if constexpr (KQ) {
        float amax = 0.0f, vmax = 0.0f; int imax = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = (i < 4 ? 4 * lane + i : 128 + 4 * lane + (i - 4));
            const float ax = fabsf(v[i]);
            if (ax > amax) { amax = ax; vmax = v[i]; imax = idx; }
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float oa = __shfl_xor_sync(0xffffffffu, amax, o), ov = __shfl_xor_sync(0xffffffffu, vmax, o);
            const int   oi = __shfl_xor_sync(0xffffffffu, imax, o);
            if (oa > amax || (oa == amax && oi < imax)) { amax = oa; vmax = ov; imax = oi; }
        }
        float d = 0.0f;
        if (amax != 0.0f) {
            const float iscale = __fdiv_rn(-127.0f, vmax);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = min(127, __float2int_rn(iscale * v[i]));
            d = __fdiv_rn(1.0f, iscale);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = 0;
        }
        if (lane == 0) *(float *)(rec + A.off_d + 4 * t) = d;
    } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float * vv = v + 4 * half;
            float amax = fmaxf(fmaxf(fabsf(vv[0]), fabsf(vv[1])), fmaxf(fabsf(vv[2]), fabsf(vv[3])));
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            const float id = amax != 0.0f ? __fdiv_rn(127.0f, amax) : 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) q[4 * half + i] = __float2int_rn(vv[i] * id);
            if ((lane & 7) == 0) {
                const int blk = 4 * half + (lane >> 3);              // 32-block index inside the act-task
                *(float *)(rec + A.off_d + ((blk >> 2) * A.ntask + t) * 16 + (blk & 3) * 4) = __half2float(__float2half_rn(__fdiv_rn(amax, 127.0f)));
            }
        }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int * qq = q + 4 * half;
        const int w = 32 * half + lane;                              // word index inside the act-task
        *(uint32_t *)(rec + ((w >> 2) * A.ntask + t) * 16 + (w & 3) * 4) =
            (uint32_t)(qq[0] & 0xFF) | ((uint32_t)(qq[1] & 0xFF) << 8) | ((uint32_t)(qq[2] & 0xFF) << 16) | ((uint32_t)(qq[3] & 0xFF) << 24);
        int s = qq[0] + qq[1] + qq[2] + qq[3];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);                     // 16 values
        if ((lane & 3) == 0) {
            const int g16 = 8 * half + (lane >> 2);
            *(int16_t *)(rec + A.off_s16 + ((g16 >> 3) * A.ntask + t) * 16 + (g16 & 7) * 2) = (int16_t)s;
        }
        s += __shfl_xor_sync(0xffffffffu, s, 4);                     // 32 values
        if ((lane & 7) == 0) {
            const int g32 = 4 * half + (lane >> 3);
            *(int32_t *)(rec + A.off_s32 + ((g32 >> 2) * A.ntask + t) * 16 + (g32 & 3) * 4) = s;
        }
    }
}

__device__ __forceinline__ int4 lds128(const uint8_t * p) { return *(const int4 *)p; }

// ----------------------------------------------------------------------------- task dot products
// `w` points at the task's first byte in the shared-memory stage, `rec` at the activation record, `t` = task index in the row.
template <int T> __device__ __forceinline__ float task_dot(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t);

// 6-bit (scale, min) pair J of the 12-byte pack, J compile-time
template <int J> __device__ __forceinline__ void k4_sm(uint32_t s0, uint32_t s1, uint32_t s2, int & sc, int & mn) {
    if constexpr (J < 4) {
        sc = (s0 >> (8 * J)) & 63;
        mn = (s1 >> (8 * J)) & 63;
    } else {
        constexpr int jj = J - 4;
        sc = ((s2 >> (8 * jj)) & 0x0F) | (((s0 >> (8 * jj + 6)) & 3) << 4);
        mn = ((s2 >> (8 * jj + 4)) & 0x0F) | (((s1 >> (8 * jj + 6)) & 3) << 4);
    }
}

template <int C, bool FIVE>
__device__ __forceinline__ void q45_chunk(const uint8_t * qs, const uint32_t (&qh)[8], const uint8_t * rec, const sb_act & A, int t,
                                           uint32_t s0, uint32_t s1, uint32_t s2, const int (&s32)[8], int & acc_s, int & acc_m) {
    const int4 qa = lds128(qs + 32 * C), qb = lds128(qs + 32 * C + 16);
    const uint32_t q[8] = { (uint32_t)qa.x, (uint32_t)qa.y, (uint32_t)qa.z, (uint32_t)qa.w, (uint32_t)qb.x, (uint32_t)qb.y, (uint32_t)qb.z, (uint32_t)qb.w };
    int p0 = 0, p1 = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int4 ylo = lds128(rec + ((4 * C + h) * A.ntask + t) * 16);          // values 64C + 16h ..   (sub-block 2C)
        const int4 yhi = lds128(rec + ((4 * C + 2 + h) * A.ntask + t) * 16);      // values 64C + 32 + 16h (sub-block 2C+1)
        const int yl[4] = { ylo.x, ylo.y, ylo.z, ylo.w }, yh[4] = { yhi.x, yhi.y, yhi.z, yhi.w };
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t qq = q[4 * h + i];
            if constexpr (FIVE) {
                const uint32_t hb = qh[4 * h + i] >> (2 * C);
                p0 = __dp4a((int)((qq & 0x0F0F0F0F) | ((hb & 0x01010101) << 4)), yl[i], p0);
                p1 = __dp4a((int)(((qq >> 4) & 0x0F0F0F0F) | ((hb & 0x02020202) << 3)), yh[i], p1);
            } else {
                p0 = __dp4a((int)(qq & 0x0F0F0F0F), yl[i], p0);
                p1 = dp4a_us(qq & 0xF0F0F0F0u, yh[i], p1);                          // 16 x (high nibbles . y)
            }
        }
    }
    if constexpr (!FIVE) p1 >>= 4;                                                  // exact: a multiple of 16
    int sc0, m0, sc1, m1;
    k4_sm<2 * C>(s0, s1, s2, sc0, m0);
    k4_sm<2 * C + 1>(s0, s1, s2, sc1, m1);
    acc_s += sc0 * p0 + sc1 * p1;
    acc_m += m0 * s32[2 * C] + m1 * s32[2 * C + 1];
}

template <bool FIVE> __device__ __forceinline__ float q45_task(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t) {
    const int4 hdr = lds128(w);                                     // d | dmin | scales[12]
    const int4 sa = lds128(rec + A.off_s32 + t * 16), sb = lds128(rec + A.off_s32 + (A.ntask + t) * 16);
    const int s32[8] = { sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w };
    uint32_t qh[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    if constexpr (FIVE) {
        const int4 ha = lds128(w + 16), hb = lds128(w + 32);
        qh[0] = ha.x; qh[1] = ha.y; qh[2] = ha.z; qh[3] = ha.w; qh[4] = hb.x; qh[5] = hb.y; qh[6] = hb.z; qh[7] = hb.w;
    }
    const uint8_t * qs = w + (FIVE ? 48 : 16);
    const uint32_t s0 = hdr.y, s1 = hdr.z, s2 = hdr.w;
    int acc_s = 0, acc_m = 0;
    q45_chunk<0, FIVE>(qs, qh, rec, A, t, s0, s1, s2, s32, acc_s, acc_m);
    q45_chunk<1, FIVE>(qs, qh, rec, A, t, s0, s1, s2, s32, acc_s, acc_m);
    q45_chunk<2, FIVE>(qs, qh, rec, A, t, s0, s1, s2, s32, acc_s, acc_m);
    q45_chunk<3, FIVE>(qs, qh, rec, A, t, s0, s1, s2, s32, acc_s, acc_m);
    const float yd = *(const float *)(rec + A.off_d + 4 * t);
    const float d = h2f((uint32_t)hdr.x & 0xFFFF) * yd, dmin = h2f((uint32_t)hdr.x >> 16) * yd;
    return d * (float)acc_s - dmin * (float)acc_m;
}
template <> __device__ __forceinline__ float task_dot<T_Q4_K>(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t) { return q45_task<false>(w, rec, A, t); }
template <> __device__ __forceinline__ float task_dot<T_Q5_K>(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t) { return q45_task<true>(w, rec, A, t); }

// Q4_0: task = 8 blocks of 18 bytes = 144 bytes (16-byte aligned), act-task == task
template <> __device__ __forceinline__ float task_dot<T_Q4_0>(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t) {
    uint32_t ww[37];
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int4 v = lds128(w + 16 * i); ww[4 * i] = v.x; ww[4 * i + 1] = v.y; ww[4 * i + 2] = v.z; ww[4 * i + 3] = v.w; }
    ww[36] = 0;
    const int4 sa = lds128(rec + A.off_s32 + t * 16), sb = lds128(rec + A.off_s32 + (A.ntask + t) * 16);
    const int s32[8] = { sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w };
    const int4 da = lds128(rec + A.off_d + t * 16), db = lds128(rec + A.off_d + (A.ntask + t) * 16);
    const float yd[8] = { __int_as_float(da.x), __int_as_float(da.y), __int_as_float(da.z), __int_as_float(da.w),
                          __int_as_float(db.x), __int_as_float(db.y), __int_as_float(db.z), __int_as_float(db.w) };
    float acc = 0.0f;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        // block b starts at byte 18 b = word 4.5 b: even b word-aligned, odd b half-word shifted (all compile-time)
        constexpr int dummy = 0; (void)dummy;
        const int w0 = (18 * b) / 4;
        const bool odd = (b & 1) != 0;
        uint32_t q[4];
        uint32_t dbits;
        if (!odd) {
            dbits = ww[w0] & 0xFFFF;
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = __funnelshift_r(ww[w0 + i], ww[w0 + i + 1], 16);
        } else {
            dbits = ww[w0] >> 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = ww[w0 + 1 + i];
        }
        const int4 ylo = lds128(rec + ((2 * b) * A.ntask + t) * 16), yhi = lds128(rec + ((2 * b + 1) * A.ntask + t) * 16);
        const int yl[4] = { ylo.x, ylo.y, ylo.z, ylo.w }, yh[4] = { yhi.x, yhi.y, yhi.z, yhi.w };
        int p0 = 0, p1 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            p0 = __dp4a((int)(q[i] & 0x0F0F0F0F), yl[i], p0);
            p1 = dp4a_us(q[i] & 0xF0F0F0F0u, yh[i], p1);
        }
        const int s = p0 + (p1 >> 4) - 8 * s32[b];
        acc += (float)s * h2f(dbits) * yd[b];
    }
    return acc;
}

// Q8_0: task = 4 blocks of 34 bytes = 136 bytes (8-byte aligned); two tasks per 256-value act-task
template <> __device__ __forceinline__ float task_dot<T_Q8_0>(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t) {
    uint32_t ww[35];
#pragma unroll
    for (int i = 0; i < 17; ++i) { const uint2 v = *(const uint2 *)(w + 8 * i); ww[2 * i] = v.x; ww[2 * i + 1] = v.y; }
    ww[34] = 0;
    const int at = t >> 1, hf = t & 1;                                             // act-task, which half of it
    const int4 dv = lds128(rec + A.off_d + (hf * A.ntask + at) * 16);
    const float yd[4] = { __int_as_float(dv.x), __int_as_float(dv.y), __int_as_float(dv.z), __int_as_float(dv.w) };
    float acc = 0.0f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int w0 = (34 * b) / 4;
        const bool odd = (b & 1) != 0;                                             // 34 b mod 4 = 2 for odd b
        const uint32_t dbits = odd ? (ww[w0] >> 16) : (ww[w0] & 0xFFFF);
        const int4 y0 = lds128(rec + ((8 * hf + 2 * b) * A.ntask + at) * 16), y1 = lds128(rec + ((8 * hf + 2 * b + 1) * A.ntask + at) * 16);
        const int y[8] = { y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w };
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t q = odd ? ww[w0 + 1 + i] : __funnelshift_r(ww[w0 + i], ww[w0 + i + 1], 16);
            s = __dp4a((int)q, y[i], s);
        }
        acc += (float)s * (h2f(dbits) * yd[b]);
    }
    return acc;
}

// Q6_K: 210-byte superblocks are only 2-byte aligned: aligned words + one run-time funnel shift (0 or 16 bits)
template <> __device__ __forceinline__ float task_dot<T_Q6_K>(const uint8_t * w, const uint8_t * rec, const sb_act & A, int t) {
    const uint32_t sh = ((uint32_t)(uintptr_t)w & 2) * 8;
    const uint32_t * wa = (const uint32_t *)((uintptr_t)w & ~(uintptr_t)3);
    auto word = [&](int i) { return __funnelshift_r(wa[i], wa[i + 1], sh); };     // 32-bit word i of the superblock
    const int4 sa = lds128(rec + A.off_s16 + t * 16), sb = lds128(rec + A.off_s16 + (A.ntask + t) * 16);
    const uint32_t s16w[8] = { (uint32_t)sa.x, (uint32_t)sa.y, (uint32_t)sa.z, (uint32_t)sa.w, (uint32_t)sb.x, (uint32_t)sb.y, (uint32_t)sb.z, (uint32_t)sb.w };
    int tot = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t scw0 = word(48 + 2 * h), scw1 = word(48 + 2 * h + 1);       // scales[8h .. 8h+7]
#pragma unroll
        for (int j = 0; j < 2; ++j) {                                              // l-range 16 j .. 16 j + 15
            int p[4] = { 0, 0, 0, 0 };
            const int4 yv0 = lds128(rec + ((8 * h + j) * A.ntask + t) * 16), yv1 = lds128(rec + ((8 * h + j + 2) * A.ntask + t) * 16);
            const int4 yv2 = lds128(rec + ((8 * h + j + 4) * A.ntask + t) * 16), yv3 = lds128(rec + ((8 * h + j + 6) * A.ntask + t) * 16);
            const int ya[4] = { yv0.x, yv0.y, yv0.z, yv0.w }, yb[4] = { yv1.x, yv1.y, yv1.z, yv1.w };
            const int yc[4] = { yv2.x, yv2.y, yv2.z, yv2.w }, yd4[4] = { yv3.x, yv3.y, yv3.z, yv3.w };
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t la = word(16 * h + 4 * j + i), lb = word(16 * h + 8 + 4 * j + i), qh = word(32 + 8 * h + 4 * j + i);
                p[0] = __dp4a((int)((la & 0x0F0F0F0F)        | ((qh << 4) & 0x30303030)), ya[i], p[0]);
                p[1] = __dp4a((int)((lb & 0x0F0F0F0F)        | ((qh << 2) & 0x30303030)), yb[i], p[1]);
                p[2] = __dp4a((int)(((la >> 4) & 0x0F0F0F0F) | ( qh       & 0x30303030)), yc[i], p[2]);
                p[3] = __dp4a((int)(((lb >> 4) & 0x0F0F0F0F) | ((qh >> 2) & 0x30303030)), yd4[i], p[3]);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int idx = j + 2 * g;                                          // scales[8h + j + 2g]
                const int sc = (int)(int8_t)(((idx < 4 ? scw0 : scw1) >> (8 * (idx & 3))) & 0xFF);
                const int g16 = 8 * h + j + 2 * g;                                  // 16-group index in the superblock
                const int bs = (int)(int16_t)((s16w[g16 >> 1] >> (16 * (g16 & 1))) & 0xFFFF);
                tot += sc * (p[g] - 32 * bs);
            }
        }
    }
    const float d = h2f(word(52) & 0xFFFF) * *(const float *)(rec + A.off_d + 4 * t);
    return d * (float)tot;
}

// ----------------------------------------------------------------------------- kernel
constexpr int SB_CONSUMER_WARPS = 8;
constexpr int SB_MAX_STAGES = 6;

struct sb_params {
    const uint8_t * w; const float * x; float * y;
    int64_t M, K;
    int32_t row_bytes, rows_per_chunk, nchunks, stage_bytes, nstages, ntasks_row;
    unsigned int * counters;      // this launch's scheduling slot: [0] next chunk, [1] finished producers, [2] finished CTAs (all return to 0)
    unsigned int * ctl;           // device-global control words: [0] exchange epoch, [1] trace launch index
    int32_t src1_static;          // activations are not produced by the preceding kernel either: never wait for it (independent ops overlap)
    int32_t src0_static;          // weights are not produced by the preceding kernel: prefetch them before griddepcontrol.wait
    // row-sharded multi-GPU: every result is stored straight into each peer's full-length y over NVLink (world == 0: off)
    // fused epilogue (bias add and GELU of the following ggml nodes): y2 = y + bias, y3 = gelu(y2); null = off
    const float * ep_bias; float * ep_y2; float * ep_y3;
    unsigned long long * dbg;     // optional %globaltimer trace (GGML_B200_SB_DEBUG=1): 32 launches x 8 stamps
    int32_t world, rank;
    int64_t row_offset;
    uint32_t epoch;
    float *    y_peers[8];
    uint32_t * flag_peers[8];
    sb_act A;
};

template <int T>
__global__ void __launch_bounds__((SB_CONSUMER_WARPS + 1) * 32, 2) mmvq_sb_kernel(const sb_params p) {
    using F = sbfmt<T>;
    constexpr int LPR = F::LPR, RPW = 32 / LPR;                 // rows per warp pass
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t * stages = smem;
    uint8_t * rec    = stages + (size_t)p.nstages * p.stage_bytes;
    uint64_t * full  = (uint64_t *)(rec + p.A.bytes);
    uint64_t * empty = full + SB_MAX_STAGES;
    int * chunk_of   = (int *)(empty + SB_MAX_STAGES);          // chunk id held by each stage (-1 = end)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    unsigned long long * dbg = nullptr;
    if (p.dbg && blockIdx.x == 0) {
        __shared__ unsigned int dbg_slot;
        if (tid == 0) dbg_slot = atomicAdd(&p.ctl[1], 1u) % 32u;
        __syncthreads();
        dbg = p.dbg + dbg_slot * 8;
        if (tid == 0) dbg[0] = gtime();                          // CTA 0 entry
    }

    if (tid == 0) {
        for (int s = 0; s < p.nstages; ++s) { sb_mbar_init(&full[s], 1); sb_mbar_init(&empty[s], SB_CONSUMER_WARPS); }
        sb_fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int s, int chunk) {
        chunk_of[s] = chunk < p.nchunks ? chunk : -1;
        if (chunk < p.nchunks) {
            const int64_t row0 = (int64_t)chunk * p.rows_per_chunk;
            const int rows = (int)min((int64_t)p.rows_per_chunk, p.M - row0);
            const uint32_t bytes = (uint32_t)rows * (uint32_t)p.row_bytes;
            sb_mbar_expect_tx(&full[s], bytes);
            sb_tma_g2s(stages + (size_t)s * p.stage_bytes, p.w + (size_t)row0 * p.row_bytes, bytes, &full[s]);
        } else {
            sb_mbar_arrive(&full[s]);                             // publish the end marker
        }
    };

    if (warp == SB_CONSUMER_WARPS) {
        // ===== producer warp: weights do not depend on the previous kernel, start streaming immediately
        if (lane == 0) {
            if (!p.src0_static) pdl_wait();
            issue(0, (int)blockIdx.x);                            // first chunk is static
            if (dbg) dbg[1] = gtime();                            // first TMA issued
            // (the scheduling counters are per launch slot, so the producer never has to wait for the previous grid on their account)
            if (dbg) dbg[2] = gtime();                            // producer past griddepcontrol.wait                                           // the chunk counter belongs to the previous launch until it completes
            int it = 1;
            bool done = (int)blockIdx.x >= p.nchunks;
            while (!done) {
                const int s = it % p.nstages;
                if (it >= p.nstages) sb_mbar_wait(&empty[s], (uint32_t)((it / p.nstages) - 1) & 1u);
                const int chunk = (int)atomicAdd(&p.counters[0], 1u) + (int)gridDim.x;
                issue(s, chunk);
                done = chunk >= p.nchunks;
                ++it;
            }
            // last CTA to finish its scheduling resets the counters for the next launch
            __threadfence();
            if (atomicAdd(&p.counters[1], 1u) == gridDim.x - 1) { p.counters[0] = 0; p.counters[1] = 0; __threadfence(); }
        }
        return;
    }

    // ===== consumers: quantize the activation vector (needs the previous kernel's output)
    if (!p.src1_static) pdl_wait();
    if (dbg && tid == 0) dbg[3] = gtime();                       // consumers past griddepcontrol.wait
    // two act-tasks per warp per round, both loads in flight before either is processed (this phase is on the critical
    // path of a dependent launch: it can only start once the previous kernel's output is visible)
    for (int t0 = warp; t0 < p.A.ntask; t0 += 2 * SB_CONSUMER_WARPS) {
        const int t1 = t0 + SB_CONSUMER_WARPS;
        const float * x0 = p.x + (size_t)t0 * 256, * x1 = p.x + (size_t)t1 * 256;
        const float4 a0 = load_f4(x0 + 4 * lane), b0 = load_f4(x0 + 128 + 4 * lane);
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = a1;
        if (t1 < p.A.ntask) { a1 = load_f4(x1 + 4 * lane); b1 = load_f4(x1 + 128 + 4 * lane); }
        sb_quantize_task<F::KQ != 0>(a0, b0, rec, p.A, t0);
        if (t1 < p.A.ntask) sb_quantize_task<F::KQ != 0>(a1, b1, rec, p.A, t1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(SB_CONSUMER_WARPS * 32) : "memory");        // consumers only

    const int sub = lane / LPR, l = lane % LPR;
    for (int it = 0;; ++it) {
        const int s = it % p.nstages;
        sb_mbar_wait(&full[s], (uint32_t)(it / p.nstages) & 1u);
        if (dbg && tid == 0 && it == 0) dbg[4] = gtime();        // first stage landed
        const int chunk = chunk_of[s];
        if (chunk < 0) { if (dbg && tid == 0) dbg[5] = gtime(); break; }   // last stage done
        const int64_t row0 = (int64_t)chunk * p.rows_per_chunk;
        const int rows = (int)min((int64_t)p.rows_per_chunk, p.M - row0);
        const uint8_t * st = stages + (size_t)s * p.stage_bytes;
        // the row loop is warp-uniform (both half-warps iterate together): the shuffles below use the full mask
        for (int r0 = warp * RPW; r0 < rows; r0 += SB_CONSUMER_WARPS * RPW) {
            const int r = r0 + sub;
            float acc = 0.0f;
            if (r < rows) {
                const uint8_t * row = st + (size_t)r * p.row_bytes;
                for (int t = l; t < p.ntasks_row; t += LPR) acc += task_dot<T>(row + (size_t)t * F::TASK_B, rec, p.A, t);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (l == 0 && r < rows) {
                if (p.world == 0) {
                    p.y[row0 + r] = acc;
                    if (p.ep_bias) {
                        const float v2 = acc + p.ep_bias[row0 + r];
                        p.ep_y2[row0 + r] = v2;
                        if (p.ep_y3) p.ep_y3[row0 + r] = gelu_ggml(v2);
                    }
                }
                else p.y_peers[p.rank][p.row_offset + row0 + r] = acc;       // own slot of the gathered y; pushed to the peers below
            }
        }
        __syncwarp();
        if (lane == 0) sb_mbar_arrive(&empty[s]);
    }
    if (p.world > 0) {
        // fused gather: the CTA that finishes last pushes this rank's slice to every peer's gathered y with coalesced 16-byte
        // stores over NVLink (one CTA -> one system-scope fence covers all the remote stores), then raises this rank's epoch in
        // every peer's flag array.  The other SMs are already free for the next (overlapping) launch.
        __shared__ int s_last;
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(SB_CONSUMER_WARPS * 32) : "memory");
        if (tid == 0) {
            s_last = atomicAdd(&p.counters[2], 1u) == gridDim.x - 1;
            if (s_last) { p.counters[2] = 0; __threadfence(); }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(SB_CONSUMER_WARPS * 32) : "memory");
        if (s_last) {
            const float * src = p.y_peers[p.rank] + p.row_offset;
            const int64_t n = p.M;
            const bool vec = ((p.row_offset | n) & 3) == 0;
            for (int q = 0; q < p.world; ++q) {
                if (q == p.rank) continue;
                float * dst = p.y_peers[q] + p.row_offset;
                if (vec) { for (int64_t i = tid; i < n / 4; i += SB_CONSUMER_WARPS * 32) ((float4 *)dst)[i] = __ldcg((const float4 *)src + i); }
                else     { for (int64_t i = tid; i < n; i += SB_CONSUMER_WARPS * 32) dst[i] = __ldcg(src + i); }
            }
            __threadfence_system();
            asm volatile("bar.sync 1, %0;" ::"n"(SB_CONSUMER_WARPS * 32) : "memory");
            if (tid == 0) {
                // the exchange epoch lives on the device (ctl[0]) so that a CUDA graph can replay the launch
                // (atomic: independent launches overlap, two grids may finish together; the flags only ever grow)
                const uint32_t e = p.epoch ? p.epoch : atomicAdd(&p.ctl[0], 1u) + 1u;
                __threadfence_system();
                for (int q = 0; q < p.world; ++q)
                    asm volatile("red.release.sys.global.max.u32 [%0], %1;" ::"l"(p.flag_peers[q] + p.rank), "r"(e) : "memory");
            }
        }
    }
}

// wait until every rank has published `epoch` in this rank's flag array (one warp)
__global__ void gather_wait_kernel(const uint32_t * flags, int world, uint32_t epoch, const unsigned int * ctl) {
    const int q = threadIdx.x;
    if (epoch == 0) epoch = ctl[0];          // the epoch this rank published with its last fused mat-vec
    if (q < world) {
        uint32_t v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
        } while ((int32_t)(v - epoch) < 0);
    }
    __syncwarp();
    __threadfence_system();
}

struct sb_plan { sb_params p; int grid, smem; };

// device control block: [0,64) global control words, [64, 64 + 64*8) 64 per-launch scheduling slots, byte 4096.. trace
static unsigned int * sb_counters() {
    static unsigned int * ptr[64] = { nullptr };
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ptr[dev]) {
        if (cudaMalloc(&ptr[dev], 8192) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(ptr[dev], 0, 8192);
    }
    return ptr[dev];
}

template <int T> static bool make_sb_plan(const ggml_b200_mul_mat_args & a, sb_plan & pl) {
    using F = sbfmt<T>;
    if (a.N != 1 || a.ne02 != 1 || a.ne03 != 1 || a.ne12 != 1 || a.ne13 != 1) return false;
    if (a.K % 256 != 0 || a.K < 256 || a.M < 1 || a.K > 32768) return false;
    const size_t rb = row_bytes(a.type, a.K);
    if (a.nb01 != rb || ((uintptr_t)a.src0 & 15) != 0 || ((uintptr_t)a.src1 & 3) != 0) return false;
    if ((a.M * rb) % 16 != 0) return false;
    static const int env_stage_kb = getenv("GGML_B200_SB_STAGE_KB") ? atoi(getenv("GGML_B200_SB_STAGE_KB")) : 36;
    static const int env_stages   = getenv("GGML_B200_SB_STAGES")   ? atoi(getenv("GGML_B200_SB_STAGES"))   : 0;
    static const int env_ctas     = getenv("GGML_B200_SB_CTAS")     ? atoi(getenv("GGML_B200_SB_CTAS"))     : 1;
    constexpr int RPW = 32 / F::LPR;
    int granule = 1; while ((granule * rb) % 16 != 0) granule *= 2;
    int step = SB_CONSUMER_WARPS * RPW; while (step % granule != 0) step *= 2;
    int rpc = (int)(((size_t)env_stage_kb * 1024) / rb) / step * step; if (rpc < step) rpc = step;
    if ((size_t)rpc * rb > 100 * 1024) {                         // very long rows: fewer rows per chunk than one full pass
        rpc = granule; while ((size_t)(rpc + granule) * rb <= 48 * 1024) rpc += granule;
        if ((size_t)rpc * rb > 100 * 1024) return false;
    }
    sb_params & p = pl.p;
    p.w = (const uint8_t *)a.src0; p.x = a.src1; p.y = a.dst; p.M = a.M; p.K = a.K;
    p.row_bytes = (int)rb; p.rows_per_chunk = rpc; p.nchunks = (int)((a.M + rpc - 1) / rpc);
    p.stage_bytes = (int)(((size_t)rpc * rb + 127) & ~(size_t)127);
    p.nstages = env_stages;       // 0 = automatic (below)
    p.ntasks_row = (int)(a.K / F::TASK_W);
    p.A = make_sb_act(a.K);
    p.ctl = sb_counters();
    static std::atomic<unsigned> seq{0};
    p.counters = p.ctl ? p.ctl + 64 + (seq.fetch_add(1) % 64u) * 8 : nullptr;
    p.src0_static = (a.flags & GGML_B200_MM_SRC0_STATIC) ? 1 : 0;
    p.world = 0; p.rank = 0; p.row_offset = 0; p.epoch = 0;
    p.ep_bias = nullptr; p.ep_y2 = nullptr; p.ep_y3 = nullptr;
    static const bool env_dbg = getenv("GGML_B200_SB_DEBUG") && atoi(getenv("GGML_B200_SB_DEBUG")) != 0;
    p.dbg = (env_dbg && p.ctl) ? (unsigned long long *)(p.ctl + 1024) : nullptr;
    p.src1_static = (a.flags & GGML_B200_MM_SRC1_STATIC) ? 1 : 0;
    for (int q = 0; q < 8; ++q) { p.y_peers[q] = nullptr; p.flag_peers[q] = nullptr; }
    if (!p.counters) return false;
    auto smem_of = [&]() { return p.nstages * p.stage_bytes + p.A.bytes + 2 * SB_MAX_STAGES * 8 + SB_MAX_STAGES * 4 + 64; };
    int ctas = env_ctas < 1 ? 1 : env_ctas > 2 ? 2 : env_ctas;
    if (p.nstages <= 0) {
        // deepest ring that still lets TWO launches be co-resident on an SM (<= 113 KB each), so that programmatic dependent
        // launch can overlap the next mat-vec's prologue and first TMA round trip with this one's tail
        p.nstages = 4;
        while (p.nstages > 2 && smem_of() > 113 * 1024) p.nstages--;
    }
    if (p.nstages < 2) p.nstages = 2;
    if (p.nstages > SB_MAX_STAGES) p.nstages = SB_MAX_STAGES;
    while (smem_of() * ctas > 222 * 1024 && p.nstages > 2) p.nstages--;
    if (smem_of() * ctas > 222 * 1024) ctas = 1;
    if (smem_of() > 222 * 1024) return false;
    pl.smem = smem_of();
    pl.grid = sm_count() * ctas;
    if (pl.grid > p.nchunks) pl.grid = p.nchunks;
    return true;
}

template <int T> static int launch_sb(const ggml_b200_mul_mat_args & a, const ggml_b200_gather * ga, cudaStream_t st, const ggml_b200_epilogue * ep = nullptr) {
    sb_plan pl;
    if (!make_sb_plan<T>(a, pl)) { set_error("mul_mat: shape not eligible for the superblock mat-vec kernel"); return GGML_B200_EUNSUPPORTED; }
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA_TRY(cudaFuncSetAttribute(mmvq_sb_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
        attr_set = true;
    }
    if (ep && ep->bias) { pl.p.ep_bias = ep->bias; pl.p.ep_y2 = ep->dst_bias; pl.p.ep_y3 = ep->unary == 1 ? ep->dst_unary : nullptr; }
    if (ga) {
        pl.p.world = ga->world; pl.p.rank = ga->rank; pl.p.row_offset = ga->row_offset; pl.p.epoch = ga->epoch;
        for (int q = 0; q < ga->world; ++q) { pl.p.y_peers[q] = ga->y_peers[q]; pl.p.flag_peers[q] = ga->flag_peers[q]; }
    }
    static const bool use_pdl = !(getenv("GGML_B200_NO_PDL") && atoi(getenv("GGML_B200_NO_PDL")) != 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid); cfg.blockDim = dim3((SB_CONSUMER_WARPS + 1) * 32); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_sb_kernel<T>, pl.p));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

bool mmvq_sb_eligible(const ggml_b200_mul_mat_args & a) {
    sb_plan pl;
    switch (a.type) {
        case T_Q4_0: return make_sb_plan<T_Q4_0>(a, pl);
        case T_Q8_0: return make_sb_plan<T_Q8_0>(a, pl);
        case T_Q4_K: return make_sb_plan<T_Q4_K>(a, pl);
        case T_Q5_K: return make_sb_plan<T_Q5_K>(a, pl);
        case T_Q6_K: return make_sb_plan<T_Q6_K>(a, pl);
        default: return false;
    }
}

int launch_mmvq_sb(const ggml_b200_mul_mat_args & a, cudaStream_t st, const ggml_b200_gather * ga, const ggml_b200_epilogue * ep) {
    switch (a.type) {
        case T_Q4_0: return launch_sb<T_Q4_0>(a, ga, st, ep);
        case T_Q8_0: return launch_sb<T_Q8_0>(a, ga, st, ep);
        case T_Q4_K: return launch_sb<T_Q4_K>(a, ga, st, ep);
        case T_Q5_K: return launch_sb<T_Q5_K>(a, ga, st, ep);
        case T_Q6_K: return launch_sb<T_Q6_K>(a, ga, st, ep);
        default: set_error("mul_mat: unsupported weight type %d", a.type); return GGML_B200_EUNSUPPORTED;
    }
}

int debug_read_trace(unsigned long long * out) {
    unsigned int * c = sb_counters();
    if (!c) return GGML_B200_ECUDA;
    B200_CUDA_TRY(cudaDeviceSynchronize());
    B200_CUDA_TRY(cudaMemcpy(out, c + 1024, 32 * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return GGML_B200_OK;
}

int launch_gather_wait(const uint32_t * flags, int world, uint32_t epoch, cudaStream_t st) {
    gather_wait_kernel<<<1, 32, 0, st>>>(flags, world, epoch, sb_counters());
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

} // namespace b200
