// Policy-side consumer of the observation tensor: the feature trunk of the reference's policy network,
// models/conv_to_fcnet_v2.py:36-66 -- Conv2D(6, 3x3, stride 1, 'valid') -> ReLU -> flatten -> Dense(32) -> ReLU -> Dense(32) -> ReLU --
// reading the uint8 observations the step kernel left in HBM (the (x - 128) / 255 of map_env.py:199 is folded into the first
// layer).  One fused kernel on the 5th-generation tensor cores: nothing but the [M, 32] features goes back to HBM.
//
//   per group of 128 agents (= UMMA M, one accumulator row per agent, one CTA):
//     for each of the 13 output rows i of the convolution:
//       A1[agent][k]  = fp16(1024 + obs[agent][45 i + k]),  k < 135: image rows i..i+2 are CONTIGUOUS bytes of the observation,
//                       so the im2col operand is a sliding window of the raw bytes (0x6400 | byte is the fp16 of 1024 + byte)
//       D1[128 x 80]  = A1[128 x 144] * B1[144 x 80]        banded weights: column (j, f) holds filter f at taps 3 (j + dj) + c
//       C [agent][n]  = fp16(relu(D1 / 255 + cb[n]))        cb folds the bias, the -128 / 255 and the 1024 offset
//       D2[128 x 32] += C[128 x 80] * W1_i[80 x 32]         Dense(32) accumulated over the 13 row blocks of its 1014 inputs
//     D3[128 x 32] = fp16(relu(D2 + b1)) * W2;  features = relu(D3 + b2)
//
// Operands live in shared memory in the canonical K-major no-swizzle layout (8 x 16-byte core matrices: element (row, k) at
// (k / 8) * rows * 16 + row * 16 + (k % 8) * 2), accumulators in tensor memory, all MMAs issued by one thread.
#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "ssd_internal.h"
#include "ssd_device.cuh"

#ifdef SSD_POLICY_TIMING  // profiles/micro/policy_timing.cu: cycles per phase of one thread per role of CTA 0
__device__ unsigned long long g_policy_cycles[3][8];
#define SSD_PT_DECL unsigned long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_t = clock64()
#define SSD_PT(slot) do { const long long n_ = clock64(); pt_[slot] += n_ - pt_t; pt_t = n_; } while (0)
#define SSD_PT_FLUSH(role) do { if (blockIdx.x == 0) for (int q_ = 0; q_ < 8; ++q_) g_policy_cycles[role][q_] = pt_[q_]; } while (0)
#else
#define SSD_PT_DECL
#define SSD_PT(slot)
#define SSD_PT_FLUSH(role)
#endif

namespace ssd {
namespace policy {

constexpr int V = 15, IMG = V * V * 3, ROWB = V * 3, CO = V - 2, NF = 6, FEAT = 32;
constexpr int GA = 128;                       // agents per group = UMMA M
constexpr int ROWK = 48;                      // one image row: 45 taps padded to 3 k-steps
constexpr int N1 = 80;                        // 13 * 6 = 78 conv outputs of one output row, padded
constexpr int K2 = 80, N2 = FEAT;             // one row block of Dense(32)
constexpr int K3 = FEAT, N3 = FEAT;
constexpr int RING = 4;                       // image-row operand blocks (and Dense weight blocks) in flight
constexpr int kProducerWarps = 4, kMmaWarp = 4, kThreads = 288;   // warps 0-3 build operands, warp 4 issues MMAs, warps 5-8 drain
// tensor memory (512 columns x 128 lanes x 32 bit): accumulators, and the A operands -- a lane is an agent, a column two fp16
constexpr int kTmemCols = 512;
constexpr int kColD1a = 0, kColD2 = 96, kColD1b = 128, kColD3 = 224;   // fp32 accumulators
constexpr int kColRing = 256;                 // + 32 * slot: one image row, 24 columns
constexpr int kColC = 384;                    // + 64 * buffer: relu(conv) of one output row, 40 columns
constexpr int kColX3 = 496;                   // fc2 operand, 16 columns

// shared-memory carve-up (bytes): the group's observation bytes and the B operands
constexpr int kB1Bytes = (3 * ROWK / 8) * N1 * 16;        // 23 040
constexpr int kB2Bytes = (K2 / 8) * N2 * 16;              // 5 120 per row block of Dense(32)
constexpr int kB3Bytes = (K3 / 8) * N3 * 16;              // 2 048
constexpr int kConstFloats = N1 + N2 + N3;                // cb[80], b1[32], b2[32]
constexpr int kHeadBytes = kB1Bytes + kB3Bytes + kConstFloats * 4;   // resident part of the blob: 25 664
constexpr int kBlobBytes = kHeadBytes + CO * kB2Bytes;    // + the 13 row blocks of Dense(32), streamed: 92 224
constexpr int kOffObs = 0;                                // 128 * 675 = 86 400 + slack for the padded taps of the last agent
constexpr int kOffB1 = 86528, kOffB3 = kOffB1 + kB1Bytes, kOffConst = kOffB3 + kB3Bytes;
constexpr int kOffW1 = kOffB1 + kHeadBytes;
constexpr int kOffBar = kOffW1 + RING * kB2Bytes;
enum { BAR_OBS = 0, BAR_ROW_FULL = 1, BAR_ROW_FREE = 5, BAR_D1_FULL = 9, BAR_D1_FREE = 11, BAR_C_FULL = 13, BAR_C_FREE = 15, BAR_W1_FREE = 17,
       BAR_D2 = 21, BAR_X3 = 22, BAR_D3 = 23, BAR_COUNT = 24 };
constexpr int kSmemBytes = kOffBar + BAR_COUNT * 8;
static_assert(kHeadBytes % 16 == 0 && kOffB1 % 128 == 0 && kOffW1 % 16 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "one CTA per SM");

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

// Three roles, each looping over the same groups g = blockIdx.x, blockIdx.x + gridDim.x, ...:
//   producers (warps 0-3, thread = agent = TMEM lane): image row r of the group -> fp16(1024 + byte) in ring block R % 4 of
//       tensor memory (R counts rows over all groups of the CTA); thread 0 also streams the Dense(32) weight block of output
//       row r - 2 into shared memory and, at the end of a group, the next group's observation bytes;
//   MMA thread (warp 4): conv(i) = 9 MMAs over ring blocks i..i+2 into D1[T % 2] (T counts output rows), then the
//       Dense(32) partial of output row i - 1 while the drain warps convert row i;
//   drain warps (5-8, thread = accumulator lane = agent): D1 -> relu -> fp16 -> C[T % 2] (tensor memory again: the A
//       operand of the Dense partial); at the end of a group D2 -> fc2 operand, D3 -> features in HBM.
// Barrier parities: the n-th use of a full barrier waits parity n & 1; the n-th reuse of a slot waits its free barrier on
// parity (n & 1) ^ 1 (passes at once for n = 0).
__global__ void __launch_bounds__(kThreads, 1) policy_features_kernel(const uint8_t* __restrict__ obs, long long M, const uint8_t* __restrict__ blob,
                                                                     float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    const float* s_const = reinterpret_cast<const float*>(smem + kOffConst);
    const long long n_groups = (M + GA - 1) / GA;

    if (tid == 0) {
        for (int b = 0; b < BAR_COUNT; ++b) {
            const bool by_warps = (b >= BAR_ROW_FULL && b < BAR_ROW_FREE) || (b >= BAR_D1_FREE && b < BAR_C_FREE) || b == BAR_X3;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + b)), "r"(by_warps ? 4 : 1) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {  // all of the SM's tensor memory (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < kHeadBytes / 16; i += kThreads)  // conv and fc2 weights, constants: once per CTA (the grid is persistent)
        reinterpret_cast<uint4*>(smem + kOffB1)[i] = reinterpret_cast<const uint4*>(blob)[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------------ producers
        const int t = tid;
        const uint32_t tlane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
        auto load_obs = [&](long long g) {  // thread 0: the group's rem * 675 contiguous bytes (a group starts 16-byte aligned)
            const long long a0 = g * GA;
            const int rem = static_cast<int>(M - a0 < GA ? M - a0 : GA);
            const uint8_t* src = obs + a0 * IMG;
            const uint32_t bytes = static_cast<uint32_t>(rem) * IMG, b16 = bytes & ~15u;
            for (uint32_t i = b16; i < bytes; ++i) smem[kOffObs + i] = src[i];
            mbar_expect_tx(bars + BAR_OBS, b16);
            bulk_g2s(smem + kOffObs, src, b16, bars + BAR_OBS);
        };
        if (t == 0 && blockIdx.x < n_groups) load_obs(blockIdx.x);
        uint32_t R = 0, T = 0, gi = 0;
        SSD_PT_DECL;
        for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
            mbar_wait(bars + BAR_OBS, gi & 1);
            SSD_PT(0);
#pragma unroll 1
            for (int r = 0; r < V; ++r, ++R) {
                const uint32_t slot = R & (RING - 1);
                uint32_t h[ROWK / 2];
                {   // fp16(1024 + byte) of this agent's 48-byte window: 0x6400 | byte
                    const int base = t * IMG + r * ROWB;
                    const uint32_t* w = reinterpret_cast<const uint32_t*>(smem + kOffObs + (base & ~3));
                    const uint32_t sh = static_cast<uint32_t>(base & 3) * 8;
                    uint32_t x[ROWK / 4 + 1];
#pragma unroll
                    for (int q = 0; q <= ROWK / 4; ++q) x[q] = w[q];
#pragma unroll
                    for (int q = 0; q < ROWK / 4; ++q) {
                        const uint32_t b4 = __funnelshift_r(x[q], x[q + 1], sh);
                        h[2 * q] = __byte_perm(b4, 0x64646464u, 0x4140);
                        h[2 * q + 1] = __byte_perm(b4, 0x64646464u, 0x4342);
                    }
                }
                SSD_PT(2);
                mbar_wait(bars + BAR_ROW_FREE + slot, ((R >> 2) & 1) ^ 1);
                SSD_PT(1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < ROWK / 2; c += 8) tmem_st8(tlane + kColRing + slot * 32 + c, h + c);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (warp == 0 && r >= 2) {  // the Dense(32) weight block of output row r - 2 rides on the same barrier
                        const uint32_t Tt = T + (r - 2), ws = Tt & (RING - 1);
                        mbar_wait(bars + BAR_W1_FREE + ws, ((Tt >> 2) & 1) ^ 1);  // read last by the partial of row Tt - 4
                        mbar_expect_tx(bars + BAR_ROW_FULL + slot, kB2Bytes);
                        bulk_g2s(smem + kOffW1 + ws * kB2Bytes, blob + kHeadBytes + (r - 2) * kB2Bytes, kB2Bytes, bars + BAR_ROW_FULL + slot);
                    } else {
                        mbar_arrive(bars + BAR_ROW_FULL + slot);
                    }
                }
                SSD_PT(3);
            }
            T += CO;
            asm volatile("bar.sync 1, 128;" ::: "memory");  // every producer is done with the observation bytes
            if (t == 0 && g + gridDim.x < n_groups) load_obs(g + gridDim.x);
            SSD_PT(4);
        }
        if (t == 0) SSD_PT_FLUSH(0);
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issue (one thread)
        if (lane == 0) {
            constexpr uint32_t kI1 = umma_idesc(GA, N1), kI2 = umma_idesc(GA, N2), kI3 = umma_idesc(GA, N3);
            const uint32_t sB1 = smem_u32(smem + kOffB1), sB3 = smem_u32(smem + kOffB3), sW1 = smem_u32(smem + kOffW1);
            uint32_t rows_seen = 0, T = 0, gi = 0;
            SSD_PT_DECL;
            auto dense_partial = [&](uint32_t Tp, int i_local) {  // D2 (+)= C[Tp % 2] * W1 block
                const uint32_t b = Tp & 1, ws = Tp & (RING - 1);
                mbar_wait(bars + BAR_C_FULL + b, (Tp >> 1) & 1);
                SSD_PT(3);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < K2 / 16; ++ks)
                    umma_f16_ts(tmem + kColD2, tmem + kColC + b * 64 + ks * 8, umma_desc(sW1 + ws * kB2Bytes + ks * 2 * N2 * 16, N2 * 16, 128), kI2,
                                (i_local | ks) != 0);
                umma_commit(bars + BAR_C_FREE + b);
                umma_commit(bars + BAR_W1_FREE + ws);
                SSD_PT(4);
            };
            for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
                const uint32_t R0 = gi * V;
#pragma unroll 1
                for (int i = 0; i < CO; ++i, ++T) {
                    while (rows_seen < R0 + i + 3) {
                        mbar_wait(bars + BAR_ROW_FULL + (rows_seen & (RING - 1)), (rows_seen >> 2) & 1);
                        ++rows_seen;
                    }
                    SSD_PT(0);
                    const uint32_t b = T & 1;
                    mbar_wait(bars + BAR_D1_FREE + b, ((T >> 1) & 1) ^ 1);
                    SSD_PT(1);
                    tc_fence_after();
#pragma unroll
                    for (int di = 0; di < 3; ++di)
#pragma unroll
                        for (int ks = 0; ks < ROWK / 16; ++ks)
                            umma_f16_ts(tmem + (b ? kColD1b : kColD1a), tmem + kColRing + ((R0 + i + di) & (RING - 1)) * 32 + ks * 8,
                                        umma_desc(sB1 + (di * (ROWK / 16) + ks) * 2 * N1 * 16, N1 * 16, 128), kI1, (di | ks) != 0);
                    umma_commit(bars + BAR_D1_FULL + b);
                    umma_commit(bars + BAR_ROW_FREE + ((R0 + i) & (RING - 1)));  // image row i is not needed again
                    if (i == CO - 1) {
                        umma_commit(bars + BAR_ROW_FREE + ((R0 + i + 1) & (RING - 1)));
                        umma_commit(bars + BAR_ROW_FREE + ((R0 + i + 2) & (RING - 1)));
                    }
                    SSD_PT(2);
                    if (i >= 1) dense_partial(T - 1, i - 1);
                }
                dense_partial(T - 1, CO - 1);
                umma_commit(bars + BAR_D2);
                mbar_wait(bars + BAR_X3, gi & 1);
                SSD_PT(5);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < K3 / 16; ++ks)
                    umma_f16_ts(tmem + kColD3, tmem + kColX3 + ks * 8, umma_desc(sB3 + ks * 2 * N3 * 16, N3 * 16, 128), kI3, ks > 0);
                umma_commit(bars + BAR_D3);
                SSD_PT(6);
            }
            SSD_PT_FLUSH(1);
        }
    } else {
        // ------------------------------------------------------------------ drain warps: thread = accumulator lane = agent
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
        float cb[N1];  // bias, -128 / 255 and the 1024 offset of the operand, per conv column
#pragma unroll
        for (int c = 0; c < N1; ++c) cb[c] = s_const[c];
        uint32_t T = 0, gi = 0;
        SSD_PT_DECL;
        for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, ++gi) {
            const long long a0 = g * GA;
            const int rem = static_cast<int>(M - a0 < GA ? M - a0 : GA);
#pragma unroll 1
            for (int i = 0; i < CO; ++i, ++T) {
                const uint32_t b = T & 1;
                mbar_wait(bars + BAR_D1_FULL + b, (T >> 1) & 1);
                SSD_PT(0);
                tc_fence_after();
                uint32_t acc[N1];
#pragma unroll
                for (int c0 = 0; c0 < N1; c0 += 16) tmem_ld16(trow + (b ? kColD1b : kColD1a) + c0, acc + c0);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + BAR_D1_FREE + b);
                SSD_PT(1);
                uint32_t h[N1 / 2];   // C = fp16(relu(D1 / 255 + cb))
#pragma unroll
                for (int c = 0; c < N1; c += 2)
                    h[c / 2] = pack_relu_h2(fmaf(__uint_as_float(acc[c]), 1.0f / 255.0f, cb[c]), fmaf(__uint_as_float(acc[c + 1]), 1.0f / 255.0f, cb[c + 1]));
                SSD_PT(2);
                mbar_wait(bars + BAR_C_FREE + b, ((T >> 1) & 1) ^ 1);
                SSD_PT(3);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < N1 / 2; c += 8) tmem_st8(trow + kColC + b * 64 + c, h + c);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + BAR_C_FULL + b);
                SSD_PT(4);
            }
            {   // fc2 operand = fp16(relu(D2 + b1))
                mbar_wait(bars + BAR_D2, gi & 1);
                tc_fence_after();
                uint32_t acc[N2], h[N2 / 2];
                tmem_ld16(trow + kColD2, acc);
                tmem_ld16(trow + kColD2 + 16, acc + 16);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < N2; c += 2) h[c / 2] = pack_relu_h2(__uint_as_float(acc[c]) + s_const[N1 + c], __uint_as_float(acc[c + 1]) + s_const[N1 + c + 1]);
                tmem_st8(trow + kColX3, h);
                tmem_st8(trow + kColX3 + 8, h + 8);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + BAR_X3);
            }
            {   // features = relu(D3 + b2), one 128-byte row per agent
                mbar_wait(bars + BAR_D3, gi & 1);
                tc_fence_after();
                uint32_t acc[N3];
                tmem_ld16(trow + kColD3, acc);
                tmem_ld16(trow + kColD3 + 16, acc + 16);
                tmem_ld_wait();
                tc_fence_before();
                if (row < rem) {
                    float4* dst = reinterpret_cast<float4*>(out + (a0 + row) * FEAT);
#pragma unroll
                    for (int c = 0; c < N3; c += 4)
                        dst[c / 4] = make_float4(fmaxf(__uint_as_float(acc[c]) + s_const[N1 + N2 + c], 0.f), fmaxf(__uint_as_float(acc[c + 1]) + s_const[N1 + N2 + c + 1], 0.f),
                                                 fmaxf(__uint_as_float(acc[c + 2]) + s_const[N1 + N2 + c + 2], 0.f), fmaxf(__uint_as_float(acc[c + 3]) + s_const[N1 + N2 + c + 3], 0.f));
                }
            }
            SSD_PT(5);
        }
        if (tid == 160) SSD_PT_FLUSH(2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace policy
}  // namespace ssd

struct SsdPolicy {
    int device = 0;
    int sms = 0;
    uint8_t* d_blob = nullptr;
};

#pragma GCC visibility push(default)
extern "C" {

int ssd_policy_create(int view_radius, int device, const float* conv_w, const float* conv_b, const float* fc1_w, const float* fc1_b,
                      const float* fc2_w, const float* fc2_b, ssd_policy_t* out) {
    using namespace ssd::policy;
    if (!out) return ssd::set_error(SSD_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!conv_w || !conv_b || !fc1_w || !fc1_b || !fc2_w || !fc2_b) return ssd::set_error(SSD_ERR_INVALID, "null weight pointer");
    if (2 * view_radius + 1 != V) return ssd::set_error(SSD_ERR_UNSUPPORTED, "the feature kernel is built for 15x15 observations (view radius 7)");
    std::vector<uint8_t> blob(kBlobBytes, 0);
    __half* b1 = reinterpret_cast<__half*>(blob.data());                          // resident: conv, fc2, constants
    __half* b3 = reinterpret_cast<__half*>(blob.data() + kB1Bytes);
    float* cst = reinterpret_cast<float*>(blob.data() + kB1Bytes + kB3Bytes);
    __half* b2 = reinterpret_cast<__half*>(blob.data() + kHeadBytes);             // streamed: the 13 row blocks of Dense(32)
    auto at = [](int rows, int n, int k) { return ((k >> 3) * rows + n) * 8 + (k & 7); };  // canonical K-major, no swizzle
    // banded conv weights: conv_w[di][dj][c][f] (Keras kernel layout) at tap k = 48 di + 3 (j + dj) + c of column n = 6 j + f
    for (int j = 0; j < CO; ++j)
        for (int f = 0; f < NF; ++f) {
            double sum16 = 0.0;
            for (int di = 0; di < 3; ++di)
                for (int dj = 0; dj < 3; ++dj)
                    for (int c = 0; c < 3; ++c) {
                        const __half h = __float2half_rn(conv_w[((di * 3 + dj) * 3 + c) * NF + f]);
                        b1[at(N1, j * NF + f, di * ROWK + (j + dj) * 3 + c)] = h;
                        sum16 += static_cast<double>(__half2float(h));
                    }
            // relu(conv((x - 128) / 255) + b) with the operand holding 1024 + x
            cst[j * NF + f] = static_cast<float>(static_cast<double>(conv_b[f]) - (1024.0 + 128.0) * sum16 / 255.0);
        }
    for (int i = 0; i < CO; ++i)  // Dense(32) kernel [1014][32], inputs flattened (i, j, f)
        for (int k = 0; k < CO * NF; ++k)
            for (int n = 0; n < N2; ++n) b2[i * (K2 * N2) + at(N2, n, k)] = __float2half_rn(fc1_w[(i * CO * NF + k) * N2 + n]);
    for (int k = 0; k < K3; ++k)
        for (int n = 0; n < N3; ++n) b3[at(N3, n, k)] = __float2half_rn(fc2_w[k * N3 + n]);
    for (int n = 0; n < N2; ++n) cst[N1 + n] = fc1_b[n];
    for (int n = 0; n < N3; ++n) cst[N1 + N2 + n] = fc2_b[n];

    SsdPolicy* p = new (std::nothrow) SsdPolicy();
    if (!p) return ssd::set_error(SSD_ERR_INVALID, "out of host memory");
    p->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_blob, kBlobBytes);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_blob, blob.data(), kBlobBytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) {
        if (p->d_blob) cudaFree(p->d_blob);
        delete p;
        return ssd::set_error(SSD_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = p;
    return SSD_OK;
}

int ssd_policy_features(ssd_policy_t p, const uint8_t* obs, int64_t num_agents, float* features, void* stream) {
    using namespace ssd::policy;
    if (!p || !obs || !features || num_agents < 0) return ssd::set_error(SSD_ERR_INVALID, "bad argument");
    if (reinterpret_cast<uintptr_t>(obs) % 16 != 0 || reinterpret_cast<uintptr_t>(features) % 16 != 0)
        return ssd::set_error(SSD_ERR_INVALID, "obs and features must be 16-byte aligned");
    if (num_agents == 0) return SSD_OK;
    const long long groups = (num_agents + GA - 1) / GA;
    const int grid = static_cast<int>(groups < p->sms ? groups : p->sms);
    policy_features_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(obs, num_agents, p->d_blob, features);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ssd::set_error(SSD_ERR_CUDA, cudaGetErrorString(e));
    return SSD_OK;
}

void ssd_policy_destroy(ssd_policy_t p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_blob);
    delete p;
}

}  // extern "C"
#pragma GCC visibility pop
