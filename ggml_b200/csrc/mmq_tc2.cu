// mmq_tc2.cu — batched quantized mat-mul on CTA PAIRS: tcgen05.mma.cta_group::2, one 256 (W rows) x BN (activation rows) tile per
// pair of SMs, the second generation of mmq_tc.cu's kernel (same operand preparation: W dequantized to fp16 in the UMMA
// SWIZZLE_128B layout, X converted to fp16 once with an exact power-of-two row scale, f32 accumulation in TMEM).
//
// Why pairs (profiles/r01_gemm_q8_0_tcgen05.md: 410 TFLOP/s = 0.28 of the measured peak at 4096^2 x 512, the tensor pipe 30 % busy):
// the one-CTA kernel pulls a whole BN x 64 activation stage per 128 W rows, so its operand feed (64 B/clk/SM of L2 traffic at BN = 256)
// and its 2-4 stage ring, not the tensor pipe, set the pace.  With cta_group::2
//   * each CTA stages only HALF of the activation tile (BN/2 rows): the MMA reads both halves, the L2 -> SM activation traffic
//     per flop halves and so does the stage size: 5 stages of 32 KB where the 256-row one-CTA tile had 2 stages of 64 KB;
//   * each CTA dequantizes only its own 128 W rows (two threads per row), so the dequantizers are never the bottleneck;
//   * one elected thread of the LEADER CTA issues every MMA for both SMs; tcgen05.commit multicasts "stage free" / "accumulator
//     ready" to the barriers of both CTAs;
//   * the activation conversion kernel and this kernel are chained with programmatic dependent launch: TMEM allocation, barrier set-up,
//     tensor-map prefetch, the raw W stream and the first dequantized stages run while the conversion is still in flight; only the
//     activation TMA waits for it.  Split-K flags live in a self-cleaning per-device block: no memset per call.
//
// Warp roles per CTA (10 warps): warp 0 = TMA producer (raw W units for the own 128 rows into a 2-deep raw ring; the own half of every
// activation stage, signalled on the LEADER's stage barrier); warp 1 = TMEM allocation (both CTAs) and, in the leader, the MMA issuer;
// warps 2-9 = dequantizers (two threads per row, each half of a unit's K-steps; the non-leader's warps arrive on the leader's stage
// barrier through the cluster), then the epilogue (each CTA its 128 accumulator lanes).
//
// Replaces the reference's mul_mat_q (src/ggml-cuda/mmq.cuh:2499-2655) for the shapes make_tc2_plan accepts; everything else stays on
// mmq_tc.cu.  Results are those of mmq_tc.cu (same operand values, f32 accumulation; only the summation grouping of split-K differs).
#include "b200_internal.h"
#include "b200_quants.cuh"
#include "b200_tc_dequant.cuh"
#include "b200_tc_ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace b200 {

constexpr int T2_BM = 128;                    // W rows per CTA (256 per pair)
constexpr int T2_BK = 64;
constexpr int T2_DQ_WARPS = 8, T2_THREADS = (2 + T2_DQ_WARPS) * 32;
constexpr int T2_MAX_STAGES = 8;

struct tc2_params {
    float * y; float * partials; unsigned int * flags; const float * inv_scale;
    int64_t M, N;
    int32_t BN, m_tiles, n_tiles, splitk, units_total, nstages, w_static;
};

template <int T, int KS>
__device__ __forceinline__ void tc2_dequant_step(int nstages, int step, bool valid, const uint32_t (&u)[tcfmt<T>::UNIT_WORDS], uint8_t * ring, int stage_bytes, int a_row_off,
                                                 uint32_t sw, int lane, uint32_t rank, uint64_t * full, uint64_t * empty) {
    const int s = step % nstages;
    if (step >= nstages) tc_wait(&empty[s], (uint32_t)((step / nstages) - 1) & 1u);
    if (valid) dq64<T, KS>(u, ring + s * stage_bytes + a_row_off, sw);
    tc_fence_async_all();                      // generic-proxy stores -> visible to the tensor core (async proxy), also from the peer SM
    __syncwarp();
    if (lane == 0) { if (rank == 0) tc_arrive(&full[s]); else tc_arrive_cluster(&full[s], 0); }
}

template <int T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
mmq_tc2_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const tc2_params p) {
    constexpr int RAW = tcfmt<T>::RAW, UK = tcfmt<T>::UNIT_KSTEPS;
    extern __shared__ __align__(1024) uint8_t smem[];
    // identical layout in both CTAs: [ring: nstages x (A 16 KB | B half BN/2 x 128 B)][raw: 2 x 128 x RAW][barriers][tmem slot]
    constexpr int a_bytes = T2_BM * T2_BK * 2;
    const int b_bytes = (p.BN / 2) * T2_BK * 2, stage_bytes = a_bytes + b_bytes;
    uint8_t * ring = smem;
    uint8_t * raw  = ring + p.nstages * stage_bytes;
    uint64_t * bars = (uint64_t *)(raw + 2 * T2_BM * RAW);
    uint64_t * full = bars, * empty = bars + T2_MAX_STAGES, * raw_full = bars + 2 * T2_MAX_STAGES, * raw_empty = raw_full + 2, * acc_full = raw_empty + 2;
    uint32_t * tmem_slot = (uint32_t *)(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = tc_cluster_ctarank();                   // 0 = leader
    tc_pdl_launch_dependents();
    // work item of the pair: partial producers (ks > 0) first, tile owners (ks == 0) last
    const int pair = (int)blockIdx.x >> 1;
    const int tiles = p.m_tiles * p.n_tiles;
    const int ks = p.splitk - 1 - pair / tiles;
    const int tile = pair % tiles, tm = tile % p.m_tiles, tn = tile / p.m_tiles;
    const int ubeg = (int)((int64_t)p.units_total * ks / p.splitk), uend = (int)((int64_t)p.units_total * (ks + 1) / p.splitk);
    const int nunits = uend - ubeg, nsteps = UK * nunits;
    const int64_t row_base = (int64_t)tm * (2 * T2_BM) + (int64_t)rank * T2_BM;     // first W row of this CTA
    const int64_t rows_left = p.M - row_base;                     // may be <= 0 for the second CTA of the last tile

    if (tid == 0) {
        // leader's stage barrier: 4 dequantizer warps of each CTA (the group that owns the K-step) + the leader's producer (expect_tx
        // for both activation halves); the non-leader's copy of it is unused.  empty / acc_full: one multicast commit each.
        for (int s = 0; s < p.nstages; ++s) { tc_mbar_init(&full[s], T2_DQ_WARPS / 2 * 2 + 1); tc_mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc_mbar_init(&raw_full[s], 1); tc_mbar_init(&raw_empty[s], T2_DQ_WARPS); }
        tc_mbar_init(acc_full, 1);
        tc_fence_init();
        tc_prefetch_map(&map_w); tc_prefetch_map(&map_x);
    }
    if (warp == 1) tc_tmem_alloc_pair(tmem_slot, tc_tmem_cols(p.BN));
    tc_fence_before();
    __syncthreads();
    tc_cluster_sync();                                            // the peer's barriers exist before anything is signalled on them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs)
        if (lane == 0) {
            if (!p.w_static) tc_pdl_wait();                       // W produced by the preceding kernel: nothing may be read before it is done
            bool x_ready = !p.w_static ? true : false;            // the fp16 activations are written by the conversion kernel just before this one
            for (int u = 0; u < nunits; ++u) {
                const int rs = u & 1;
                if (u >= 2) tc_wait(&raw_empty[rs], (uint32_t)((u >> 1) - 1) & 1u);
                tc_expect_tx(&raw_full[rs], T2_BM * RAW);
                int coord;                                        // first 4-byte word of the box: 16-byte aligned start at or below the unit
                if constexpr (tcfmt<T>::LOAD_BYTES == 2) coord = (((ubeg + u) * tcfmt<T>::UNIT_BYTES) & ~15) >> 2;
                else                       coord = (ubeg + u) * tcfmt<T>::STRIDE_WORDS - ((ubeg + u) & 1) * tcfmt<T>::ODD_BACK_WORDS;
                tc_tma_2d(raw + rs * T2_BM * RAW, &map_w, coord, (int)row_base, &raw_full[rs]);
                if (!x_ready) { tc_pdl_wait(); x_ready = true; }
                for (int q = 0; q < UK; ++q) {
                    const int step = UK * u + q, s = step % p.nstages;
                    if (step >= p.nstages) tc_wait(&empty[s], (uint32_t)((step / p.nstages) - 1) & 1u);
                    if (rank == 0) tc_expect_tx(&full[s], (uint32_t)(2 * b_bytes));
                    tc_tma_2d_pair(ring + s * stage_bytes + a_bytes, &map_x, ((ubeg + u) * UK + q) * T2_BK, tn * p.BN + (int)rank * (p.BN / 2), tc_cluster_addr(&full[s], 0));
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only)
        if (rank == 0) {
            // instruction descriptor: D = f32 (bit 4), A = B = f16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24 with M = 256 (the pair)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((2 * T2_BM) >> 4) << 24);
            for (int step = 0; step < nsteps; ++step) {
                const int s = step % p.nstages;
                tc_wait_cluster(&full[s], (uint32_t)(step / p.nstages) & 1u);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t ad = tc_smem_desc(tc_smem(ring + s * stage_bytes));
                    const uint64_t bd = tc_smem_desc(tc_smem(ring + s * stage_bytes + a_bytes));
#pragma unroll
                    for (int k = 0; k < T2_BK / 16; ++k)
                        tc_mma_f16_pair(tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (step | k) != 0 ? 1u : 0u);   // +32 bytes per K = 16
                    tc_commit_pair(&empty[s]);
                    if (step == nsteps - 1) tc_commit_pair(acc_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== dequantizers: two threads per row (warps 2-5 / 6-9), each half of the unit's K-steps
        const int dq = tid - 64, dwarp = dq >> 5;
        const int row = dq & 127, ksel = dq >> 7;
        const bool valid = row < rows_left;                       // rows past M: the TMA box is zero-filled, nothing to convert
        const uint32_t sw = (uint32_t)(row & 7);
        const int a_row_off = (row >> 3) * 1024 + (row & 7) * 128;
        uint32_t ub[tcfmt<T>::UNIT_WORDS];
        for (int u = 0; u < nunits; ++u) {
            const int rs = u & 1;
            tc_wait(&raw_full[rs], (uint32_t)(u >> 1) & 1u);
            int lead;                                             // bytes between the box start and the unit's first byte
            if constexpr (tcfmt<T>::LOAD_BYTES == 2) lead = ((ubeg + u) * tcfmt<T>::UNIT_BYTES) & 15;
            else                       lead = ((ubeg + u) & 1) * (4 * tcfmt<T>::ODD_BACK_WORDS);
            tc_load_unit<T>(raw + rs * T2_BM * RAW + row * RAW + lead, ub);
            __syncwarp();
            if (lane == 0) tc_arrive(&raw_empty[rs]);            // the unit is in registers: the buffer can be refilled
            if constexpr (UK == 4) {
                if (ksel == 0) {
                    tc2_dequant_step<T, 0>(p.nstages, UK * u + 0, valid, ub, ring, stage_bytes, a_row_off, sw, lane, rank, full, empty);
                    tc2_dequant_step<T, 1>(p.nstages, UK * u + 1, valid, ub, ring, stage_bytes, a_row_off, sw, lane, rank, full, empty);
                } else {
                    tc2_dequant_step<T, 2>(p.nstages, UK * u + 2, valid, ub, ring, stage_bytes, a_row_off, sw, lane, rank, full, empty);
                    tc2_dequant_step<T, 3>(p.nstages, UK * u + 3, valid, ub, ring, stage_bytes, a_row_off, sw, lane, rank, full, empty);
                }
            } else {
                if (ksel == 0) tc2_dequant_step<T, 0>(p.nstages, UK * u + 0, valid, ub, ring, stage_bytes, a_row_off, sw, lane, rank, full, empty);
                else           tc2_dequant_step<T, 1>(p.nstages, UK * u + 1, valid, ub, ring, stage_bytes, a_row_off, sw, lane, rank, full, empty);
            }
        }
        // ===================== epilogue: this CTA's 128 accumulator lanes x BN columns
        tc_wait(acc_full, 0);
        tc_fence_after();
        // each warp owns its TMEM lane quarter (hardware: warp id % 4) and one half of the columns
        const int lg = warp & 3, grp = dwarp >> 2;
        const int ncol = p.BN / 2, col0 = grp * ncol;
        const int64_t m = row_base + lg * 32 + lane;
        const int64_t n_base = (int64_t)tn * p.BN + col0;
        const uint32_t tacc = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)col0;
        const int mloc = lg * 32 + lane;
        const int fidx = tile * 2 + (int)rank;
        float * part = p.partials ? p.partials + ((size_t)fidx * (p.splitk - 1)) * (size_t)(p.BN * T2_BM) : nullptr;
        if (ks > 0) {
            // split-K partial: [ks-1][n_local][m_local]
            float * dst = part + (size_t)(ks - 1) * (p.BN * T2_BM);
            for (int c0 = 0; c0 < ncol; c0 += 32) {
                float v[32];
                tc_ld32(tacc + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) dst[(size_t)(col0 + c0 + i) * T2_BM + mloc] = v[i];
            }
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(T2_DQ_WARPS * 32) : "memory");
            if (dq == 0) atomicAdd(&p.flags[fidx], 1u);
        } else {
            if (p.splitk > 1) {
                if (dq == 0) { while (atomicAdd(&p.flags[fidx], 0u) < (unsigned)(p.splitk - 1)) __nanosleep(64); __threadfence(); }
                asm volatile("bar.sync 1, %0;" ::"n"(T2_DQ_WARPS * 32) : "memory");
            }
            for (int c0 = 0; c0 < ncol; c0 += 32) {
                float v[32];
                tc_ld32(tacc + (uint32_t)c0, v);
                for (int j = 1; j < p.splitk; ++j) {
                    const float * src = part + (size_t)(j - 1) * (p.BN * T2_BM);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __ldcg(&src[(size_t)(col0 + c0 + i) * T2_BM + mloc]);
                }
                if (m < p.M) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const int64_t n = n_base + c0 + i; if (n < p.N) p.y[(size_t)n * p.M + m] = v[i] * __ldcg(p.inv_scale + n); }
                }
            }
            if (p.splitk > 1) {
                asm volatile("bar.sync 1, %0;" ::"n"(T2_DQ_WARPS * 32) : "memory");
                if (dq == 0) p.flags[fidx] = 0;                  // leave the flag clean for the next launch that gets this slot
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    tc_cluster_sync();                                            // neither CTA retires (nor frees TMEM) while the pair still works
    if (warp == 1) { tc_fence_after(); tc_tmem_dealloc_pair(tmem, tc_tmem_cols(p.BN)); }
}

// ----------------------------------------------------------------------------- split-K flags: persistent, zero-initialised, self-cleaning
constexpr int T2_FLAG_SLOTS = 64, T2_FLAGS_PER_SLOT = 256;
static unsigned int * tc_flag_block() {
    static unsigned int * ptr[64] = { nullptr };
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("tc flags: cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lock(mu);
    if (!ptr[dev]) {
        unsigned int * p = nullptr;
        const size_t bytes = (size_t)T2_FLAG_SLOTS * T2_FLAGS_PER_SLOT * sizeof(unsigned int);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("tc flags: %s", cudaGetErrorString(e)); cudaGetLastError(); if (p) cudaFree(p); return nullptr; }
        ptr[dev] = p;
    }
    return ptr[dev];
}
int tc_prepare_device() { return tc_flag_block() ? GGML_B200_OK : GGML_B200_ECUDA; }
unsigned int * tc_flag_slot() {
    unsigned int * b = tc_flag_block();
    if (!b) return nullptr;
    static std::atomic<unsigned> seq{0};
    return b + (size_t)(seq.fetch_add(1, std::memory_order_relaxed) % T2_FLAG_SLOTS) * T2_FLAGS_PER_SLOT;
}

// ----------------------------------------------------------------------------- host side
struct tc2_plan {
    int BN, m_tiles, n_tiles, splitk, chunks, nstages, smem, grid;
    size_t xb_bytes, partial_bytes, scale_bytes;
};

static int tc2_raw_bytes(int type) {
    switch (type) {
        case T_Q5_K: case T_Q5_0: return 176;
        case T_Q6_K: return 224;
        case T_Q4_1: return 160;
        case T_Q5_1: return 192;
        case T_Q2_K: return 96;
        case T_Q3_K: return 128;
        default: return 144;                                         // Q4_0, Q8_0 (half units), Q4_K, IQ4_NL, IQ4_XS
    }
}

static bool make_tc2_plan(const ggml_b200_mul_mat_args & a, tc2_plan & pl) {
    static const int env_mode = getenv("GGML_B200_TC_PAIR") ? atoi(getenv("GGML_B200_TC_PAIR")) : 1;     // 0 = off (one-CTA kernel everywhere)
    if (env_mode == 0) return false;
    static const bool env_q6k_off = getenv("GGML_B200_TC_Q6K") && atoi(getenv("GGML_B200_TC_Q6K")) == 0;
    switch (a.type) {
        case T_Q4_0: case T_Q8_0: case T_Q4_K: case T_Q5_K: case T_Q4_1: case T_Q5_0: case T_Q5_1: case T_IQ4_NL: case T_IQ4_XS: case T_Q2_K: case T_Q3_K: break;
        case T_Q6_K: if (env_q6k_off) return false; break;
        default: return false;
    }
    if (a.ne02 != 1 || a.ne03 != 1 || a.ne12 != 1 || a.ne13 != 1) return false;
    // pairs pay off once a 256-row tile is mostly full and the batch fills at least a 64-column tile; smaller problems keep the one-CTA kernel
    if (a.N < 33 || a.M < 192 || a.K % 256 != 0 || a.K < 256) return false;
    const size_t rb = row_bytes(a.type, a.K);
    if (a.nb01 != rb || (rb % 16) != 0 || ((uintptr_t)a.src0 & 15) != 0 || ((uintptr_t)a.src1 & 3) != 0 || (a.nb11 & 3) != 0) return false;
    if (a.M >= (1ll << 31) || a.N >= (1ll << 31) || rb >= (1ull << 31)) return false;
    if (!tc_get_encode()) return false;
    static const int env_bn = getenv("GGML_B200_TC2_BN") ? atoi(getenv("GGML_B200_TC2_BN")) : 0;
    int BN = a.N > 128 ? 256 : a.N > 64 ? 128 : 64;
    if (env_bn == 64 || env_bn == 128 || env_bn == 256) BN = env_bn;
    pl.BN = BN;
    pl.n_tiles = (int)((a.N + BN - 1) / BN);
    pl.m_tiles = (int)((a.M + 2 * T2_BM - 1) / (2 * T2_BM));
    pl.chunks = (int)(a.K / (a.type == T_Q8_0 ? 128 : 256));     // raw units along K (tcfmt<T>::UNIT_KSTEPS x 64 weights each)
    const int tiles = pl.m_tiles * pl.n_tiles;
    const int pairs = sm_count() / 2;
    int splitk = pairs / tiles; if (splitk < 1) splitk = 1; if (splitk > 8) splitk = 8; if (splitk > pl.chunks) splitk = pl.chunks;
    static const int env_splitk = getenv("GGML_B200_TC_SPLITK") ? atoi(getenv("GGML_B200_TC_SPLITK")) : 0;
    if (env_splitk > 0 && env_splitk <= pl.chunks) splitk = env_splitk;
    if (splitk > 1 && tiles * 2 > T2_FLAGS_PER_SLOT) splitk = 1;
    pl.splitk = splitk;
    const int raw = tc2_raw_bytes(a.type);
    auto smem_of = [&](int ns) { return ns * (T2_BM * T2_BK * 2 + (BN / 2) * T2_BK * 2) + 2 * T2_BM * raw + 256 + 1024; };
    static const int env_stages = getenv("GGML_B200_TC2_STAGES") ? atoi(getenv("GGML_B200_TC2_STAGES")) : 0;
    int nstages = env_stages >= 2 && env_stages <= T2_MAX_STAGES ? env_stages : T2_MAX_STAGES;
    while (nstages > 2 && smem_of(nstages) > 227 * 1024) nstages--;
    if (smem_of(nstages) > 227 * 1024) return false;
    pl.nstages = nstages; pl.smem = smem_of(nstages);
    pl.grid = 2 * tiles * splitk;
    pl.xb_bytes = ((size_t)a.N * a.K * 2 + 255) & ~(size_t)255;
    pl.partial_bytes = splitk > 1 ? (size_t)tiles * 2 * (splitk - 1) * BN * T2_BM * 4 : 0;
    pl.scale_bytes = ((size_t)a.N * 4 + 255) & ~(size_t)255;
    return true;
}

bool mmq_tc2_eligible(const ggml_b200_mul_mat_args & a) { tc2_plan pl; return make_tc2_plan(a, pl); }
size_t mmq_tc2_workspace(const ggml_b200_mul_mat_args & a) {
    tc2_plan pl;
    if (!make_tc2_plan(a, pl)) return 0;
    return pl.xb_bytes + pl.partial_bytes + pl.scale_bytes + 1024;
}

template <int T> static int launch_tc2(const ggml_b200_mul_mat_args & a, const tc2_plan & pl, cudaStream_t st) {
    const size_t need = pl.xb_bytes + pl.partial_bytes + pl.scale_bytes + 1024;
    if (!a.workspace || a.workspace_size < need) { set_error("mul_mat: workspace %zu < %zu", a.workspace_size, need); return GGML_B200_EWORKSPACE; }
    uint8_t * ws = (uint8_t *)(((uintptr_t)a.workspace + 255) & ~(uintptr_t)255);
    __half * xb = (__half *)ws;
    float * partials = pl.partial_bytes ? (float *)(ws + pl.xb_bytes) : nullptr;
    float * inv_scale = (float *)(ws + pl.xb_bytes + pl.partial_bytes);
    unsigned int * flags = tc_flag_slot();
    if (!flags) return GGML_B200_ECUDA;
    static const bool use_pdl = !(getenv("GGML_B200_NO_PDL") && atoi(getenv("GGML_B200_NO_PDL")) != 0);

    { const int rc = tc_launch_x_to_f16(a.src1, a.nb11, xb, inv_scale, a.K, a.N, st, use_pdl); if (rc != GGML_B200_OK) return rc; }
    const size_t rb = row_bytes(a.type, a.K);
    alignas(64) CUtensorMap map_w, map_x;
    {
        const cuuint64_t dims[2] = { (cuuint64_t)(rb / 4), (cuuint64_t)a.M };
        const cuuint64_t strides[1] = { (cuuint64_t)rb };
        const cuuint32_t box[2] = { (cuuint32_t)(tcfmt<T>::RAW / 4), (cuuint32_t)T2_BM };
        const cuuint32_t es[2] = { 1, 1 };
        CUresult r = tc_get_encode()(&map_w, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)a.src0, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(W) failed: %d", (int)r); return GGML_B200_ECUDA; }
    }
    {
        const cuuint64_t dims[2] = { (cuuint64_t)a.K, (cuuint64_t)a.N };
        const cuuint64_t strides[1] = { (cuuint64_t)a.K * 2 };
        const cuuint32_t box[2] = { (cuuint32_t)T2_BK, (cuuint32_t)(pl.BN / 2) };
        const cuuint32_t es[2] = { 1, 1 };
        CUresult r = tc_get_encode()(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)xb, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(X) failed: %d", (int)r); return GGML_B200_ECUDA; }
    }
    tc2_params p;
    p.y = a.dst; p.partials = partials; p.flags = flags; p.inv_scale = inv_scale; p.M = a.M; p.N = a.N;
    p.BN = pl.BN; p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.splitk = pl.splitk; p.units_total = pl.chunks; p.nstages = pl.nstages;
    p.w_static = (a.flags & GGML_B200_MM_SRC0_STATIC) ? 1 : 0;
    static per_device_flag attr_set;
    if (!attr_set.test()) { B200_CUDA_TRY(cudaFuncSetAttribute(mmq_tc2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_set.set(); }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl.grid); cfg.blockDim = dim3(T2_THREADS); cfg.dynamicSmemBytes = (size_t)pl.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;             // the cluster shape is compiled in (__cluster_dims__)
    B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmq_tc2_kernel<T>, map_w, map_x, p));
    B200_LAUNCH_CHECK();
    return GGML_B200_OK;
}

int launch_mmq_tc2(const ggml_b200_mul_mat_args & a, cudaStream_t st) {
    tc2_plan pl;
    if (!make_tc2_plan(a, pl)) { set_error("mul_mat: shape not eligible for the CTA-pair tcgen05 kernel"); return GGML_B200_EUNSUPPORTED; }
    switch (a.type) {
        case T_Q4_0:   return launch_tc2<T_Q4_0>(a, pl, st);
        case T_Q8_0:   return launch_tc2<T_Q8_0>(a, pl, st);
        case T_Q4_K:   return launch_tc2<T_Q4_K>(a, pl, st);
        case T_Q5_K:   return launch_tc2<T_Q5_K>(a, pl, st);
        case T_Q6_K:   return launch_tc2<T_Q6_K>(a, pl, st);
        case T_Q4_1:   return launch_tc2<T_Q4_1>(a, pl, st);
        case T_Q5_0:   return launch_tc2<T_Q5_0>(a, pl, st);
        case T_Q5_1:   return launch_tc2<T_Q5_1>(a, pl, st);
        case T_IQ4_NL: return launch_tc2<T_IQ4_NL>(a, pl, st);
        case T_IQ4_XS: return launch_tc2<T_IQ4_XS>(a, pl, st);
        case T_Q2_K:   return launch_tc2<T_Q2_K>(a, pl, st);
        case T_Q3_K:   return launch_tc2<T_Q3_K>(a, pl, st);
        default: set_error("mul_mat: unsupported weight type %d for the tcgen05 kernel", a.type); return GGML_B200_EUNSUPPORTED;
    }
}

} // namespace b200
