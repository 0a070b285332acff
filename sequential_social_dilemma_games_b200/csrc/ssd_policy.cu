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

namespace ssd {
namespace policy {

constexpr int V = 15, IMG = V * V * 3, ROWB = V * 3, CO = V - 2, NF = 6, FEAT = 32;
constexpr int GA = 128;                       // agents per group = UMMA M
constexpr int K1 = 144, N1 = 80;              // 3 * 45 = 135 taps padded to 9 k-steps; 13 * 6 = 78 columns padded
constexpr int K2 = 80, N2 = FEAT;             // one row block of Dense(32)
constexpr int K3 = FEAT, N3 = FEAT;
constexpr int kThreads = 128;
constexpr int kTmemCols = 256;                // D1 at column 0 (80 used), D2 at 128, D3 at 160
constexpr int kColD1 = 0, kColD2 = 128, kColD3 = 160;

// shared-memory carve-up (bytes)
constexpr int kObsBytes = GA * IMG;                       // 86 400, a multiple of 16
constexpr int kOffObs = 0;
constexpr int kOffA = 86528;                              // obs + slack for the padded taps of the last agent
constexpr int kABytes = (K1 / 8) * GA * 16;               // 36 864; also holds C (20 480) and the fc2 operand (8 192)
constexpr int kOffW = kOffA + kABytes;                    // the packed weights, in the order of the blob
constexpr int kB1Bytes = (K1 / 8) * N1 * 16;              // 23 040
constexpr int kB2Bytes = (K2 / 8) * N2 * 16;              // 5 120 per row block
constexpr int kB3Bytes = (K3 / 8) * N3 * 16;              // 2 048
constexpr int kConstFloats = N1 + N2 + N3;                // cb[80], b1[32], b2[32]
constexpr int kBlobBytes = kB1Bytes + CO * kB2Bytes + kB3Bytes + kConstFloats * 4;   // 92 224
constexpr int kOffB1 = kOffW, kOffB2 = kOffB1 + kB1Bytes, kOffB3 = kOffB2 + CO * kB2Bytes, kOffConst = kOffB3 + kB3Bytes;
constexpr int kOffBar = kOffW + ((kBlobBytes + 15) & ~15);
constexpr int kSmemBytes = kOffBar + 16;
static_assert(kBlobBytes % 16 == 0 && kOffA % 128 == 0 && kOffW % 128 == 0, "operand tiles must be 16-byte aligned");
static_assert(kSmemBytes <= 227 * 1024, "one CTA per SM");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

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
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// 16 consecutive accumulator columns of this thread's row (TMEM lane = 32 * (warp % 4) + lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
}
__device__ __forceinline__ uint32_t pack_relu_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(fmaxf(a, 0.f), fmaxf(b, 0.f));
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1) policy_features_kernel(const uint8_t* __restrict__ obs, long long M, const uint8_t* __restrict__ blob,
                                                                     float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t sA = smem_u32(smem + kOffA), sB1 = smem_u32(smem + kOffB1), sB2 = smem_u32(smem + kOffB2), sB3 = smem_u32(smem + kOffB3);
    const uint32_t bar = smem_u32(smem + kOffBar);
    const float* s_const = reinterpret_cast<const float*>(smem + kOffConst);

    if (warp == 0) {  // tensor memory for the three accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = t; i < kBlobBytes / 16; i += kThreads)  // packed weights: once per CTA (the grid is persistent)
        reinterpret_cast<uint4*>(smem + kOffW)[i] = reinterpret_cast<const uint4*>(blob)[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);  // this warp's quarter of the 128 lanes
    uint32_t parity = 0;
    constexpr uint32_t kI1 = umma_idesc(GA, N1), kI2 = umma_idesc(GA, N2), kI3 = umma_idesc(GA, N3);

    const long long n_groups = (M + GA - 1) / GA;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long a0 = g * GA;
        const int rem = static_cast<int>(M - a0 < GA ? M - a0 : GA);
        {   // the group's observations: rem * 675 contiguous bytes (the group starts 16-byte aligned: 128 * 675 = 5400 * 16)
            const uint8_t* src = obs + a0 * IMG;
            const int bytes = rem * IMG, vec = bytes >> 4;
            for (int i = t; i < vec; i += kThreads) reinterpret_cast<uint4*>(smem + kOffObs)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
            for (int i = (vec << 4) + t; i < bytes; i += kThreads) smem[kOffObs + i] = src[i];
        }
        __syncthreads();

#pragma unroll 1
        for (int i = 0; i < CO; ++i) {
            {   // A1: this agent's 144-byte window as fp16(1024 + byte), eight taps per 16-byte store
                const int base = t * IMG + i * ROWB;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(smem + kOffObs + (base & ~3));
                const uint32_t sh = static_cast<uint32_t>(base & 3) * 8;
                uint4* dst = reinterpret_cast<uint4*>(smem + kOffA + t * 16);
                uint32_t w0 = w[0];
#pragma unroll
                for (int kc = 0; kc < K1 / 8; ++kc) {
                    const uint32_t w1 = w[2 * kc + 1], w2 = w[2 * kc + 2];
                    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                    dst[kc * GA] = make_uint4(__byte_perm(lo, 0x64646464u, 0x4140), __byte_perm(lo, 0x64646464u, 0x4342),
                                              __byte_perm(hi, 0x64646464u, 0x4140), __byte_perm(hi, 0x64646464u, 0x4342));
                    w0 = w2;
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (t == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < K1 / 16; ++ks)
                    umma_f16(tmem + kColD1, umma_desc(sA + ks * 2 * GA * 16, GA * 16, 128), umma_desc(sB1 + ks * 2 * N1 * 16, N1 * 16, 128), kI1, ks > 0);
                umma_commit(bar);
            }
            bar_wait(bar, parity);
            parity ^= 1;
            tc_fence_after();
            {   // C = fp16(relu(D1 / 255 + cb)) into the (now free) operand buffer, as the K-major A operand of Dense(32)
                uint4* dst = reinterpret_cast<uint4*>(smem + kOffA + t * 16);
#pragma unroll
                for (int c0 = 0; c0 < N1; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + kColD1 + c0, v);
#pragma unroll
                    for (int q = 0; q < 16; ++q) v[q] = v[q] * (1.0f / 255.0f) + s_const[c0 + q];
                    dst[(c0 / 8) * GA] = make_uint4(pack_relu_h2(v[0], v[1]), pack_relu_h2(v[2], v[3]), pack_relu_h2(v[4], v[5]), pack_relu_h2(v[6], v[7]));
                    dst[(c0 / 8 + 1) * GA] = make_uint4(pack_relu_h2(v[8], v[9]), pack_relu_h2(v[10], v[11]), pack_relu_h2(v[12], v[13]), pack_relu_h2(v[14], v[15]));
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (t == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < K2 / 16; ++ks)
                    umma_f16(tmem + kColD2, umma_desc(sA + ks * 2 * GA * 16, GA * 16, 128),
                             umma_desc(sB2 + i * kB2Bytes + ks * 2 * N2 * 16, N2 * 16, 128), kI2, (i | ks) != 0);
                umma_commit(bar);
            }
            bar_wait(bar, parity);  // C is consumed: the buffer can take the next window
            parity ^= 1;
        }
        tc_fence_after();
        {   // fc2 operand = fp16(relu(D2 + b1))
            uint4* dst = reinterpret_cast<uint4*>(smem + kOffA + t * 16);
#pragma unroll
            for (int c0 = 0; c0 < N2; c0 += 16) {
                float v[16];
                tmem_ld16(trow + kColD2 + c0, v);
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] += s_const[N1 + c0 + q];
                dst[(c0 / 8) * GA] = make_uint4(pack_relu_h2(v[0], v[1]), pack_relu_h2(v[2], v[3]), pack_relu_h2(v[4], v[5]), pack_relu_h2(v[6], v[7]));
                dst[(c0 / 8 + 1) * GA] = make_uint4(pack_relu_h2(v[8], v[9]), pack_relu_h2(v[10], v[11]), pack_relu_h2(v[12], v[13]), pack_relu_h2(v[14], v[15]));
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < K3 / 16; ++ks)
                umma_f16(tmem + kColD3, umma_desc(sA + ks * 2 * GA * 16, GA * 16, 128), umma_desc(sB3 + ks * 2 * N3 * 16, N3 * 16, 128), kI3, ks > 0);
            umma_commit(bar);
        }
        bar_wait(bar, parity);
        parity ^= 1;
        tc_fence_after();
        {   // features = relu(D3 + b2), one 128-byte row per agent
            float4* dst = reinterpret_cast<float4*>(out + (a0 + t) * FEAT);
#pragma unroll
            for (int c0 = 0; c0 < N3; c0 += 16) {
                float v[16];
                tmem_ld16(trow + kColD3 + c0, v);
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = fmaxf(v[q] + s_const[N1 + N2 + c0 + q], 0.f);
                if (t < rem) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[c0 / 4 + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();  // the observation buffer and the accumulators are free for the next group
        tc_fence_after();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
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
    __half* b1 = reinterpret_cast<__half*>(blob.data());
    __half* b2 = reinterpret_cast<__half*>(blob.data() + kB1Bytes);
    __half* b3 = reinterpret_cast<__half*>(blob.data() + kB1Bytes + CO * kB2Bytes);
    float* cst = reinterpret_cast<float*>(blob.data() + kB1Bytes + CO * kB2Bytes + kB3Bytes);
    auto at = [](int rows, int n, int k) { return ((k >> 3) * rows + n) * 8 + (k & 7); };  // canonical K-major, no swizzle
    // banded conv weights: conv_w[di][dj][c][f] (Keras kernel layout) at tap k = 45 di + 3 (j + dj) + c of column n = 6 j + f
    for (int j = 0; j < CO; ++j)
        for (int f = 0; f < NF; ++f) {
            double sum16 = 0.0;
            for (int di = 0; di < 3; ++di)
                for (int dj = 0; dj < 3; ++dj)
                    for (int c = 0; c < 3; ++c) {
                        const __half h = __float2half_rn(conv_w[((di * 3 + dj) * 3 + c) * NF + f]);
                        b1[at(N1, j * NF + f, di * ROWB + (j + dj) * 3 + c)] = h;
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
