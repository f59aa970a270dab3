// mmq_tc.cu — batched quantized mat-mul (n > 8) on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), sm_100a.
//
// Replaces the reference's mul_mat_q (src/ggml-cuda/mmq.cuh:2499-2655: int8 mma.sync tiles + stream-k fix-up) and its
// cuBLAS fallback (dequantize_block_* -> cublasGemmEx, src/ggml-cuda/ggml-cuda.cu:1158-1300); computes GGML_OP_MUL_MAT
// for block-quantized src0 and f32 src1 as ggml_compute_forward_mul_mat does (src/ggml-cpu/ggml-cpu.c:7428).
//
// Per-block scales cannot be interposed in a TMEM accumulation that runs over the whole K loop, so the scales are
// folded into the operand: W tiles are dequantized to bf16 in shared memory (UMMA K-major SWIZZLE_128B layout),
// X is converted to bf16 once, D accumulates in f32 in TMEM.  bf16 rounding of both operands gives NMSE ~3e-6
// against the CPU backend (gate 5e-4, tests/test-backend-ops.cpp:1915-1917).
//
// One CTA = one 128 (W rows) x BN (activation rows, <= 256) output tile over a K range (split-K so that the grid
// fills the 148 SMs).  Warp roles (10 warps):
//   warp 0      TMA producer: raw packed W bytes, 128 rows x one 256-weight chunk per row, with a 2-D tensor map
//               over the byte matrix [M][row_bytes] (coalesced whatever the block size), double buffered; and the
//               bf16 X tiles (BN x 64) with a SWIZZLE_128B tensor map into the operand ring.
//   warp 1      allocates TMEM, issues tcgen05.mma (one elected lane): 4 x (128 x BN x 16) per 64-wide K step,
//               tcgen05.commit releases operand stages and finally signals the epilogue.
//   warps 2-9   dequantize raw W -> bf16 -> swizzled A stage (generic-proxy stores + fence.proxy.async), then run
//               the epilogue: tcgen05.ld the accumulator (each warp its TMEM lane quarter and column half), and
//               either write Y, or hand a split-K partial to the CTA that owns the tile (flag in the workspace;
//               partial producers are scheduled first, so the wait cannot deadlock).
#include "b200_internal.h"
#include "b200_quants.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <mutex>

namespace b200 {

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t tc_smem(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t * b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void tc_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_expect_tx(uint64_t * b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tc_arrive(uint64_t * b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem(b)) : "memory"); }
__device__ __forceinline__ void tc_wait(uint64_t * b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nTC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE;\nbra TC_WAIT;\nTC_DONE:\n}\n" ::"r"(tc_smem(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_tma_2d(void * dst, const CUtensorMap * map, int c0, int c1, uint64_t * bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(tc_smem(dst)), "l"(map), "r"(tc_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_prefetch_map(const CUtensorMap * map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_tmem_alloc(uint32_t * dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *(uint32_t *)&v;
}

// ----------------------------------------------------------------------------- W chunk geometry
// a chunk = 256 weights of one row; RAW = bytes the TMA box copies per row per chunk (multiple of 16)
template <int T> struct tcfmt;
template <> struct tcfmt<T_Q4_0> { static constexpr int RAW = 144; };
template <> struct tcfmt<T_Q8_0> { static constexpr int RAW = 272; };
template <> struct tcfmt<T_Q4_K> { static constexpr int RAW = 144; };
template <> struct tcfmt<T_Q5_K> { static constexpr int RAW = 176; };

__device__ __forceinline__ float s8_to_f(uint32_t word, int byte) {        // signed byte -> float without I2F
    const uint32_t u = __byte_perm(word ^ 0x80808080u, 0x4B000000u, 0x7650 + byte) ;   // {b, 0x00, 0x00, 0x4B}
    return __uint_as_float(u) - 8388736.0f;                                  // 2^23 + 128
}
__device__ __forceinline__ float u8_to_f(uint32_t word, int byte) {
    const uint32_t u = __byte_perm(word, 0x4B000000u, 0x7650 + byte);
    return __uint_as_float(u) - 8388608.0f;
}

// dequantize the 64 weights of K-step KS (0..3) of a chunk: out = 8 x (8 bf16 = 16 bytes), chunk c holds k = 8c..8c+7
template <int T, int KS> __device__ __forceinline__ void dq64(const uint8_t * raw, uint4 (&out)[8]) {
    if constexpr (T == T_Q8_0) {
        // blocks 2KS, 2KS+1 at bytes 34 b; the 68 bytes [68 KS, 68 KS + 68) lie in 16-byte chunks [4 KS, 4 KS + 5)
        uint32_t w[20];
#pragma unroll
        for (int i = 0; i < 5; ++i) { const uint4 v = *(const uint4 *)(raw + 16 * (4 * KS + i)); w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
        constexpr int o = KS;                                   // word offset of block 2KS inside the window ((68 KS - 64 KS) / 4)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const uint32_t dbits = b == 0 ? (w[o] & 0xFFFF) : (w[o + 8] >> 16);
            const float d = h2f(dbits);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t q0, q1;
                if (b == 0) { q0 = __funnelshift_r(w[o + 2 * c], w[o + 2 * c + 1], 16); q1 = __funnelshift_r(w[o + 2 * c + 1], w[o + 2 * c + 2], 16); }
                else        { q0 = w[o + 9 + 2 * c]; q1 = w[o + 10 + 2 * c]; }
                uint4 r;
                r.x = pack_bf16(s8_to_f(q0, 0) * d, s8_to_f(q0, 1) * d); r.y = pack_bf16(s8_to_f(q0, 2) * d, s8_to_f(q0, 3) * d);
                r.z = pack_bf16(s8_to_f(q1, 0) * d, s8_to_f(q1, 1) * d); r.w = pack_bf16(s8_to_f(q1, 2) * d, s8_to_f(q1, 3) * d);
                out[4 * b + c] = r;
            }
        }
    } else if constexpr (T == T_Q4_0) {
        // blocks 2KS, 2KS+1: 36 bytes at 36 KS (4-byte aligned): 16-byte chunks covering [36 KS, 36 KS + 36)
        constexpr int c0 = (36 * KS) / 16, o = ((36 * KS) % 16) / 4;
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const uint4 v = *(const uint4 *)(raw + 16 * (c0 + i)); w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w; }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const uint32_t dbits = b == 0 ? (w[o] & 0xFFFF) : (w[o + 4] >> 16);
            const float d = h2f(dbits), m8 = -8.0f * d;
            uint32_t q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = b == 0 ? __funnelshift_r(w[o + i], w[o + i + 1], 16) : w[o + 5 + i];
            // byte j of qs: low nibble -> element j, high nibble -> element j + 16
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int hi = c >> 1;                          // chunks 0,1: elements 0..15 (low nibbles); 2,3: 16..31 (high)
                const uint32_t qa = (q[2 * (c & 1)] >> (4 * hi)) & 0x0F0F0F0F, qb = (q[2 * (c & 1) + 1] >> (4 * hi)) & 0x0F0F0F0F;
                uint4 r;
                r.x = pack_bf16(fmaf(u8_to_f(qa, 0), d, m8), fmaf(u8_to_f(qa, 1), d, m8)); r.y = pack_bf16(fmaf(u8_to_f(qa, 2), d, m8), fmaf(u8_to_f(qa, 3), d, m8));
                r.z = pack_bf16(fmaf(u8_to_f(qb, 0), d, m8), fmaf(u8_to_f(qb, 1), d, m8)); r.w = pack_bf16(fmaf(u8_to_f(qb, 2), d, m8), fmaf(u8_to_f(qb, 3), d, m8));
                out[4 * b + c] = r;
            }
        }
    } else {   // Q4_K / Q5_K: 64-chunk KS of the superblock: sub-blocks 2KS (low nibbles) and 2KS+1 (high nibbles)
        constexpr bool FIVE = (T == T_Q5_K);
        const uint4 hdr = *(const uint4 *)raw;
        const uint8_t * qs = raw + (FIVE ? 48 : 16) + 32 * KS;
        const uint4 qa = *(const uint4 *)qs, qb = *(const uint4 *)(qs + 16);
        const uint32_t q[8] = { qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w };
        uint32_t qh[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        if constexpr (FIVE) {
            const uint4 ha = *(const uint4 *)(raw + 16), hb = *(const uint4 *)(raw + 32);
            qh[0] = ha.x; qh[1] = ha.y; qh[2] = ha.z; qh[3] = ha.w; qh[4] = hb.x; qh[5] = hb.y; qh[6] = hb.z; qh[7] = hb.w;
        }
        const uint32_t s0 = hdr.y, s1 = hdr.z, s2 = hdr.w;
        const float dd = h2f(hdr.x & 0xFFFF), dm = h2f(hdr.x >> 16);
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
            constexpr int dummy = 0; (void)dummy;
            int sc, mn;
            const int J = 2 * KS + hi;
            if (J < 4) { sc = (s0 >> (8 * J)) & 63; mn = (s1 >> (8 * J)) & 63; }
            else { const int jj = J - 4; sc = ((s2 >> (8 * jj)) & 0x0F) | (((s0 >> (8 * jj + 6)) & 3) << 4); mn = ((s2 >> (8 * jj + 4)) & 0x0F) | (((s1 >> (8 * jj + 6)) & 3) << 4); }
            const float d = dd * (float)sc, m = -(dm * (float)mn);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t a = (q[2 * c] >> (4 * hi)) & 0x0F0F0F0F, b = (q[2 * c + 1] >> (4 * hi)) & 0x0F0F0F0F;
                if constexpr (FIVE) {
                    a |= ((qh[2 * c] >> (2 * KS + hi)) & 0x01010101) << 4;
                    b |= ((qh[2 * c + 1] >> (2 * KS + hi)) & 0x01010101) << 4;
                }
                uint4 r;
                r.x = pack_bf16(fmaf(u8_to_f(a, 0), d, m), fmaf(u8_to_f(a, 1), d, m)); r.y = pack_bf16(fmaf(u8_to_f(a, 2), d, m), fmaf(u8_to_f(a, 3), d, m));
                r.z = pack_bf16(fmaf(u8_to_f(b, 0), d, m), fmaf(u8_to_f(b, 1), d, m)); r.w = pack_bf16(fmaf(u8_to_f(b, 2), d, m), fmaf(u8_to_f(b, 3), d, m));
                out[4 * hi + c] = r;
            }
        }
    }
}

// ----------------------------------------------------------------------------- kernel
constexpr int TC_BM = 128, TC_BK = 64, TC_DQ_WARPS = 8, TC_THREADS = (2 + TC_DQ_WARPS) * 32;
constexpr int TC_MAX_STAGES = 4;

struct tc_params {
    float * y; float * partials; unsigned int * flags;
    int64_t M, N;
    int32_t BN, m_tiles, n_tiles, splitk, chunks_total, nstages, raw_words_per_chunk;   // raw_words_per_chunk: row advance per chunk in 4-byte elements
};

// dequantizer main loop of one thread: row `row` of the tile, K-steps 2*KH and 2*KH+1 of every chunk.  KH is a template
// parameter so that the code paths never have to be merged through register moves.
template <int T, int KH>
__device__ __forceinline__ void tc_dequant_loop(const tc_params & p, int nchunks, int row, int lane, uint8_t * a_ring, int a_bytes, const uint8_t * raw,
                                                uint64_t * full, uint64_t * empty, uint64_t * raw_full, uint64_t * raw_empty) {
    constexpr int RAW = tcfmt<T>::RAW;
    const uint32_t sw = (uint32_t)(row & 7);
    const int a_row_off = (row >> 3) * 1024 + (row & 7) * 128;
    for (int c = 0; c < nchunks; ++c) {
        const int rs = c & 1;
        tc_wait(&raw_full[rs], (uint32_t)(c >> 1) & 1u);
        const uint8_t * rr = raw + rs * TC_BM * RAW + row * RAW;
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            uint4 v[8];
            if (qq == 0) dq64<T, 2 * KH>(rr, v); else dq64<T, 2 * KH + 1>(rr, v);
            if (qq == 1) { __syncwarp(); if (lane == 0) tc_arrive(&raw_empty[rs]); }     // raw bytes are in registers now
            const int step = 4 * c + 2 * KH + qq, s = step % p.nstages;
            if (step >= p.nstages) tc_wait(&empty[s], (uint32_t)((step / p.nstages) - 1) & 1u);
            uint8_t * a_row = a_ring + s * a_bytes + a_row_off;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) *(uint4 *)(a_row + (((uint32_t)ch ^ sw) << 4)) = v[ch];
            tc_fence_async_smem();
            __syncwarp();
            if (lane == 0) tc_arrive(&full[s]);
        }
    }
}

template <int T>
__global__ void __launch_bounds__(TC_THREADS, 1)
mmq_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const tc_params p) {
    constexpr int RAW = tcfmt<T>::RAW;
    extern __shared__ __align__(1024) uint8_t smem[];
    // [A ring: nstages x 16 KB][B ring: nstages x BN*128][raw: 2 x 128*RAW][barriers]
    const int a_bytes = TC_BM * TC_BK * 2, b_bytes = p.BN * TC_BK * 2;
    uint8_t * a_ring = smem;
    uint8_t * b_ring = a_ring + p.nstages * a_bytes;
    uint8_t * raw    = b_ring + p.nstages * b_bytes;
    uint64_t * bars  = (uint64_t *)(raw + 2 * TC_BM * RAW);
    uint64_t * full = bars, * empty = bars + TC_MAX_STAGES, * raw_full = bars + 2 * TC_MAX_STAGES, * raw_empty = raw_full + 2, * acc_full = raw_empty + 2;
    uint32_t * tmem_slot = (uint32_t *)(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work item: partial producers (ks > 0) first, tile owners (ks == 0) last
    const int tiles = p.m_tiles * p.n_tiles;
    const int ks = p.splitk - 1 - (int)blockIdx.x / tiles;
    const int tile = (int)blockIdx.x % tiles, tm = tile % p.m_tiles, tn = tile / p.m_tiles;
    const int cbeg = (int)((int64_t)p.chunks_total * ks / p.splitk), cend = (int)((int64_t)p.chunks_total * (ks + 1) / p.splitk);
    const int nchunks = cend - cbeg, nsteps = 4 * nchunks;

    if (tid == 0) {
        for (int s = 0; s < p.nstages; ++s) { tc_mbar_init(&full[s], 4 + 1); tc_mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc_mbar_init(&raw_full[s], 1); tc_mbar_init(&raw_empty[s], TC_DQ_WARPS); }
        tc_mbar_init(acc_full, 1);
        tc_fence_init();
        tc_prefetch_map(&map_w); tc_prefetch_map(&map_x);
    }
    if (warp == 1) tc_tmem_alloc(tmem_slot, (uint32_t)(p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int rs = c & 1;
                if (c >= 2) tc_wait(&raw_empty[rs], (uint32_t)((c >> 1) - 1) & 1u);
                tc_expect_tx(&raw_full[rs], TC_BM * RAW);
                tc_tma_2d(raw + rs * TC_BM * RAW, &map_w, (cbeg + c) * p.raw_words_per_chunk, tm * TC_BM, &raw_full[rs]);
                for (int q = 0; q < 4; ++q) {
                    const int step = 4 * c + q, s = step % p.nstages;
                    if (step >= p.nstages) tc_wait(&empty[s], (uint32_t)((step / p.nstages) - 1) & 1u);
                    tc_expect_tx(&full[s], (uint32_t)b_bytes);
                    tc_tma_2d(b_ring + s * b_bytes, &map_x, ((cbeg + c) * 4 + q) * TC_BK, tn * p.BN, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        for (int step = 0; step < nsteps; ++step) {
            const int s = step % p.nstages;
            tc_wait(&full[s], (uint32_t)(step / p.nstages) & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t ad = tc_smem_desc(tc_smem(a_ring + s * a_bytes)), bd = tc_smem_desc(tc_smem(b_ring + s * b_bytes));
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                    tc_mma_bf16(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (step | k) != 0 ? 1u : 0u);   // +32 bytes per K=16
                tc_commit(&empty[s]);
                if (step == nsteps - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ===================== dequantizers: thread -> (row, pair of K-steps of the chunk); 4 warps per pair
        const int dq = tid - 64, row = dq & 127, ksel = dq >> 7, dwarp = dq >> 5;
        if (ksel == 0) tc_dequant_loop<T, 0>(p, nchunks, row, lane, a_ring, a_bytes, raw, full, empty, raw_full, raw_empty);
        else           tc_dequant_loop<T, 1>(p, nchunks, row, lane, a_ring, a_bytes, raw, full, empty, raw_full, raw_empty);
        // ===================== epilogue
        tc_wait(acc_full, 0);
        tc_fence_after();
        // each warp owns its TMEM lane quarter (warp % 4) and one of up to four column groups (multiples of 32 columns)
        const int lg = warp & 3;
        const int ngroups = p.BN >= 64 * (TC_DQ_WARPS / 4) / 2 && TC_DQ_WARPS >= 16 ? 4 : p.BN >= 64 ? 2 : 1;
        const int cgrp = dwarp >> 2;
        const bool active = cgrp < ngroups;
        const int ncol = p.BN / ngroups;
        const int chalf = active ? cgrp : 0;
        const int64_t m = (int64_t)tm * TC_BM + lg * 32 + lane;
        const int col0 = chalf * ncol;
        const int64_t n_base = (int64_t)tn * p.BN + col0;
        float * part = p.partials ? p.partials + ((size_t)tile * (p.splitk - 1)) * (size_t)(p.BN * TC_BM) : nullptr;
        if (ks > 0) {
            // split-K partial: [ks-1][n_local][m_local]
            float * dst = part + (size_t)(ks - 1) * (p.BN * TC_BM);
            if (active) {
                for (int c0 = 0; c0 < ncol; c0 += 32) {
                    float v[32];
                    tc_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(col0 + c0), v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) dst[(size_t)(col0 + c0 + i) * TC_BM + lg * 32 + lane] = v[i];
                }
            }
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(TC_DQ_WARPS * 32) : "memory");
            if (dq == 0) atomicAdd(&p.flags[tile], 1u);
        } else {
            if (p.splitk > 1) {
                if (dq == 0) { while (atomicAdd(&p.flags[tile], 0u) < (unsigned)(p.splitk - 1)) __nanosleep(64); __threadfence(); }
                asm volatile("bar.sync 1, %0;" ::"n"(TC_DQ_WARPS * 32) : "memory");
            }
            if (active) {
                for (int c0 = 0; c0 < ncol; c0 += 32) {
                    float v[32];
                    tc_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(col0 + c0), v);
                    for (int j = 1; j < p.splitk; ++j) {
                        const float * src = part + (size_t)(j - 1) * (p.BN * TC_BM);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += __ldcg(&src[(size_t)(col0 + c0 + i) * TC_BM + lg * 32 + lane]);
                    }
                    if (m < p.M) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { const int64_t n = n_base + c0 + i; if (n < p.N) p.y[(size_t)n * p.M + m] = v[i]; }
                    }
                }
            }
            if (p.splitk > 1) {
                asm volatile("bar.sync 1, %0;" ::"n"(TC_DQ_WARPS * 32) : "memory");
                if (dq == 0) p.flags[tile] = 0;                 // leave the flag clean for the next launch
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tc_tmem_dealloc(tmem, (uint32_t)(p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256)); }
}

// ----------------------------------------------------------------------------- X -> bf16 prologue
__global__ void __launch_bounds__(256) x_to_bf16_kernel(const float * __restrict__ x, size_t nb11, __nv_bfloat16 * __restrict__ xb, int64_t K, int64_t N) {
    const int64_t n = blockIdx.y;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (k >= K) return;
    const float * xr = (const float *)((const uint8_t *)x + n * nb11) + k;
    const float4 a = load_f4(xr), b = load_f4(xr + 4);
    uint4 o; o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
    *(uint4 *)(xb + n * K + k) = o;
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void * p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (encode_tiled_fn)p;
        else cudaGetLastError();
    });
    return fn;
}

struct tc_plan {
    int BN, m_tiles, n_tiles, splitk, chunks, nstages, smem, grid;
    size_t xb_bytes, partial_bytes, flags_bytes;
};

static bool make_tc_plan(const ggml_b200_mul_mat_args & a, tc_plan & pl) {
    if (a.type != T_Q4_0 && a.type != T_Q8_0 && a.type != T_Q4_K && a.type != T_Q5_K) return false;
    if (a.ne02 != 1 || a.ne03 != 1 || a.ne12 != 1 || a.ne13 != 1) return false;
    if (a.N < 16 || a.K % 256 != 0 || a.K < 256 || a.M < 1) return false;
    const size_t rb = row_bytes(a.type, a.K);
    if (a.nb01 != rb || (rb % 16) != 0 || ((uintptr_t)a.src0 & 15) != 0 || ((uintptr_t)a.src1 & 3) != 0 || (a.nb11 & 3) != 0) return false;
    if (a.M >= (1ll << 31) || a.N >= (1ll << 31) || rb >= (1ull << 31)) return false;
    if (!get_encode()) return false;
    const int raw = a.type == T_Q8_0 ? 272 : a.type == T_Q5_K ? 176 : 144;
    int BN = a.N >= 256 ? 256 : (int)((a.N + 15) / 16 * 16);
    if (BN > 128 && BN < 256) BN = 256;
    if (BN > 64 && BN < 128) BN = 128;
    if (BN < 32) BN = 32;                                    // TMEM allocations are powers of two >= 32 columns
    if (BN > 32 && BN < 64) BN = 64;
    pl.BN = BN;
    pl.m_tiles = (int)((a.M + TC_BM - 1) / TC_BM);
    pl.n_tiles = (int)((a.N + BN - 1) / BN);
    pl.chunks = (int)(a.K / 256);
    const int tiles = pl.m_tiles * pl.n_tiles;
    int splitk = sm_count() / tiles; if (splitk < 1) splitk = 1; if (splitk > 8) splitk = 8; if (splitk > pl.chunks) splitk = pl.chunks;
    static const int env_splitk = getenv("GGML_B200_TC_SPLITK") ? atoi(getenv("GGML_B200_TC_SPLITK")) : 0;
    if (env_splitk > 0 && env_splitk <= pl.chunks) splitk = env_splitk;
    pl.splitk = splitk;
    int nstages = TC_MAX_STAGES;
    auto smem_of = [&](int ns) { return ns * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 2 * TC_BM * raw + 256 + 1024; };
    while (nstages > 2 && smem_of(nstages) > 225 * 1024) nstages--;
    if (smem_of(nstages) > 225 * 1024) return false;
    pl.nstages = nstages; pl.smem = smem_of(nstages);
    pl.grid = tiles * splitk;
    pl.xb_bytes = ((size_t)a.N * a.K * 2 + 255) & ~(size_t)255;
    pl.partial_bytes = splitk > 1 ? (size_t)tiles * (splitk - 1) * BN * TC_BM * 4 : 0;
    pl.flags_bytes = ((size_t)tiles * 4 + 255) & ~(size_t)255;
    return true;
}

bool mmq_tc_eligible(const ggml_b200_mul_mat_args & a) { tc_plan pl; return make_tc_plan(a, pl); }
size_t mmq_tc_workspace(const ggml_b200_mul_mat_args & a) {
    tc_plan pl;
    if (!make_tc_plan(a, pl)) return 0;
    return pl.xb_bytes + pl.partial_bytes + pl.flags_bytes + 1024;
}

template <int T> static int launch_tc(const ggml_b200_mul_mat_args & a, const tc_plan & pl, cudaStream_t st) {
    const size_t need = pl.xb_bytes + pl.partial_bytes + pl.flags_bytes + 1024;
    if (!a.workspace || a.workspace_size < need) { set_error("mul_mat: workspace %zu < %zu", a.workspace_size, need); return GGML_B200_EWORKSPACE; }
    uint8_t * ws = (uint8_t *)(((uintptr_t)a.workspace + 255) & ~(uintptr_t)255);
    __nv_bfloat16 * xb = (__nv_bfloat16 *)ws;
    float * partials = pl.partial_bytes ? (float *)(ws + pl.xb_bytes) : nullptr;
    unsigned int * flags = (unsigned int *)(ws + pl.xb_bytes + pl.partial_bytes);

    // flags must start at zero: the tile owners leave them clean, but the workspace may be fresh memory
    B200_CUDA_TRY(cudaMemsetAsync(flags, 0, pl.flags_bytes, st));
    {
        dim3 grid((unsigned)((a.K / 8 + 255) / 256), (unsigned)a.N);
        x_to_bf16_kernel<<<grid, 256, 0, st>>>(a.src1, a.nb11, xb, a.K, a.N);
        B200_LAUNCH_CHECK();
    }
    const size_t rb = row_bytes(a.type, a.K);
    alignas(64) CUtensorMap map_w, map_x;
    {
        const cuuint64_t dims[2] = { (cuuint64_t)(rb / 4), (cuuint64_t)a.M };
        const cuuint64_t strides[1] = { (cuuint64_t)rb };
        const cuuint32_t box[2] = { (cuuint32_t)(tcfmt<T>::RAW / 4), (cuuint32_t)TC_BM };
        const cuuint32_t es[2] = { 1, 1 };
        CUresult r = get_encode()(&map_w, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)a.src0, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(W) failed: %d", (int)r); return GGML_B200_ECUDA; }
    }
    {
        const cuuint64_t dims[2] = { (cuuint64_t)a.K, (cuuint64_t)a.N };
        const cuuint64_t strides[1] = { (cuuint64_t)a.K * 2 };
        const cuuint32_t box[2] = { (cuuint32_t)TC_BK, (cuuint32_t)pl.BN };
        const cuuint32_t es[2] = { 1, 1 };
        CUresult r = get_encode()(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)xb, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(X) failed: %d", (int)r); return GGML_B200_ECUDA; }
    }
    tc_params p;
    p.y = a.dst; p.partials = partials; p.flags = flags; p.M = a.M; p.N = a.N;
    p.BN = pl.BN; p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.splitk = pl.splitk; p.chunks_total = pl.chunks; p.nstages = pl.nstages;
    p.raw_words_per_chunk = type_bytes(a.type) * (256 / type_qk(a.type)) / 4;
    static bool attr_set = false;
    if (!attr_set) { B200_CUDA_TRY(cudaFuncSetAttribute(mmq_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); attr_set = true; }
    mmq_tc_kernel<T><<<pl.grid, TC_THREADS, pl.smem, st>>>(map_w, map_x, p);
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int launch_mmq_tc(const ggml_b200_mul_mat_args & a, cudaStream_t st) {
    tc_plan pl;
    if (!make_tc_plan(a, pl)) { set_error("mul_mat: shape not eligible for the tcgen05 kernel"); return GGML_B200_EUNSUPPORTED; }
    switch (a.type) {
        case T_Q4_0: return launch_tc<T_Q4_0>(a, pl, st);
        case T_Q8_0: return launch_tc<T_Q8_0>(a, pl, st);
        case T_Q4_K: return launch_tc<T_Q4_K>(a, pl, st);
        case T_Q5_K: return launch_tc<T_Q5_K>(a, pl, st);
        default: set_error("mul_mat: unsupported weight type %d for the tcgen05 kernel", a.type); return GGML_B200_EUNSUPPORTED;
    }
}

} // namespace b200
