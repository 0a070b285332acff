// tcgen05 / tensor-memory PTX wrappers shared by the policy kernels (ssd_policy.cu, ssd_policy_head.cu).  sm_100a only.
#pragma once
#include <cuda_fp16.h>

#include <cstdint>

#include "ssd_device.cuh"

namespace ssd {
namespace umma {

// UMMA shared-memory descriptor, K-major, no swizzle: LBO = distance of the two 8-element k-halves of one MMA,
// SBO = distance of consecutive 8-row groups; version 1 (sm_100) in bits 46-47.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16 |
           static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32 | 1ull << 46;
}
// instruction descriptor of kind::f16: fp16 x fp16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return 1u << 4 | static_cast<uint32_t>(n >> 3) << 17 | static_cast<uint32_t>(m >> 4) << 24;
}
// D[tmem] (+)= A[smem] * B[smem]: 128 x N x 16
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: 128 x N x 16, A = eight columns of the 128 lanes
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// (a suspend-time hint on try_wait measured slower: wake-ups come later)
__device__ __forceinline__ bool try_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!try_wait_sleep(bar, parity)) {
    }
}
// one lane waits, the warp follows (__syncwarp orders the other lanes' accesses after lane 0's acquire)
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) wait_sleep(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive columns of this thread's lane (TMEM lane = 32 * (warp % 4) + lane); no wait
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {  // {fp16(max(lo, 0)), fp16(max(hi, 0))}, lo in the low half
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// all 512 columns or a power of two >= 32; one warp allocates, the same warp frees
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

}  // namespace umma
}  // namespace ssd
