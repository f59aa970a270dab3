// cuda_shim.h — TEST INFRASTRUCTURE ONLY.  Just enough of the CUDA device vocabulary to compile the *decode logic* of
// ggml_b200/csrc/b200_quants.cuh (format traits, unit dot products, element decoders) with the host compiler, so that indexing /
// bit-twiddling mistakes in a block format are caught by the CPU-only test suite, before any GPU time is spent.
// Nothing here is part of the product; warp collectives are stubs (the activation quantizers are not exercised on the host).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <immintrin.h>
#include <pthread.h>
#include <thread>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct int4   { int x, y, z, w; };
struct uint4  { unsigned x, y, z, w; };
struct uint2  { unsigned x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
struct dim3_shim { unsigned x = 0, y = 0, z = 0; };
static thread_local dim3_shim threadIdx;          // one host thread per emulated lane (warp_emu below)
static dim3_shim blockDim, blockIdx;

struct __half { uint16_t x; };
static inline __half __ushort_as_half(unsigned short u) { return __half{u}; }
static inline float  __half2float(__half h) { return _cvtsh_ss(h.x); }
static inline __half __float2half_rn(float f) { return __half{(uint16_t)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT)}; }

// packed halves: every operation is computed exactly (double) and rounded once to fp16, as the hardware does
struct __half2 { __half x, y; };
static inline __half h_rn(double v) { return __float2half_rn((float)v); }   // float(v) is exact or innocuously double-rounded: |mantissa| >= 2 * 11 + 2
static inline __half2 __float2half2_rn(float f) { const __half h = __float2half_rn(f); return __half2{h, h}; }
static inline __half2 __hsub2(__half2 a, __half2 b) { return __half2{h_rn((double)__half2float(a.x) - __half2float(b.x)), h_rn((double)__half2float(a.y) - __half2float(b.y))}; }
static inline __half2 __hmul2(__half2 a, __half2 b) { return __half2{h_rn((double)__half2float(a.x) * __half2float(b.x)), h_rn((double)__half2float(a.y) * __half2float(b.y))}; }
static inline __half2 __hfma2(__half2 a, __half2 b, __half2 c) {
    return __half2{h_rn((double)__half2float(a.x) * __half2float(b.x) + __half2float(c.x)), h_rn((double)__half2float(a.y) * __half2float(b.y) + __half2float(c.y))};
}

static inline int __dp4a(int a, int b, int c) {                       // signed x signed bytes
    for (int i = 0; i < 4; ++i) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
}
static inline int __dp4a(unsigned a, int b, int c) { return __dp4a((int)a, b, c); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (shift & 31));
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t sel) {     // 3-bit selectors, as the CUDA intrinsic
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int   __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline int   __float2int_rn(float v) { return (int)nearbyintf(v); }
// Warp collectives: when a warp is being emulated (warp_emu::run: 32 host threads in lockstep, one per lane) a shuffle is an
// exchange through a shared slot array between two barriers; outside of it, the identity (single-lane code paths).
namespace warp_emu {
    inline pthread_barrier_t & barrier() { static pthread_barrier_t b; return b; }
    inline uint64_t * slots() { static uint64_t s[32]; return s; }
    inline bool & active() { static bool a = false; return a; }
    template <typename F> inline void run(F && fn) {            // fn() is executed by 32 lanes; every lane must reach the same shuffles
        pthread_barrier_init(&barrier(), nullptr, 32);
        active() = true;
        blockDim.x = 32;
        std::thread th[32];
        for (int l = 0; l < 32; ++l) th[l] = std::thread([&, l] { threadIdx.x = (unsigned)l; fn(); });
        for (auto & t : th) t.join();
        active() = false;
        pthread_barrier_destroy(&barrier());
    }
}
template <typename V> static inline V __shfl_xor_sync(unsigned, V v, int o) {
    static_assert(sizeof(V) <= 8, "shuffle payload");
    if (!warp_emu::active()) return v;
    const int lane = (int)(threadIdx.x & 31);
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(V));
    warp_emu::slots()[lane] = raw;
    pthread_barrier_wait(&warp_emu::barrier());
    const uint64_t got = warp_emu::slots()[lane ^ o];
    pthread_barrier_wait(&warp_emu::barrier());
    V r; std::memcpy(&r, &got, sizeof(V));
    return r;
}
using std::min;
using std::max;
