// b200_sb_ptx.cuh — PTX helpers shared by the bandwidth mat-vec kernels (mmvq_sb.cu, mmvq_mma.cu): mbarriers, bulk copies (TMA),
// programmatic dependent launch, L2 prefetch.
#pragma once
#include <stdint.h>

namespace b200 {

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
// pull a byte range into L2 without occupying shared memory (bulk prefetch; 16-byte aligned address and size)
__device__ __forceinline__ void sb_prefetch_l2(const void * src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }


// device control block of the bandwidth kernels (mmvq_sb.cu): [0,64) global control words, [64, 64 + 64*8) 64 per-launch scheduling slots
unsigned int * sb_control_block();          // nullptr on error (set_error called)
unsigned int * sb_next_slot(unsigned int * ctl);   // the next of the 64 self-resetting scheduling slots

} // namespace b200
