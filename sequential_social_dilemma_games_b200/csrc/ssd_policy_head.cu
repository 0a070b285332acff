// LSTM(128) + logits / value heads + action sampling of the reference's policy network (models/conv_to_fcnet_v2.py:68-92)
// as ONE kernel behind ssd_policy_lstm_heads: per group of 128 agents (= UMMA M, one TMEM lane per agent)
//
//   A[128 x 160]  = fp16([features | h])                                  built from HBM into shared memory
//   G[128 x 512]  = A * [W; U]            (Keras gate order i, f, c~, o)   20 MMAs 128 x 256 x 16, all of tensor memory
//   c' = sigmoid(f) c + sigmoid(i) tanh(c~),  h' = sigmoid(o) tanh(c')     drain: thread = agent, 8 warps split the units
//   Y[128 x 16]   = fp16(h') * [logits_w | value_w]                        8 MMAs 128 x 16 x 16 into the freed columns
//   action        = argmax(logits + Gumbel noise)                          Philox4x32-10 keyed (seed; agent, counter)
//
// The recurrent state lives in HBM in a TILED layout chosen for this kernel: [group of 128 agents][block of 16 units (8)]
// [agent in group (128)][16 floats] -- element (agent m, unit u) at ((m / 128 * 8 + u / 16) * 128 + m % 128) * 16 + u % 16.
// A thread owns an agent and works through 16 units at a time, so a warp's loads and stores of c / h are 2 KB contiguous
// (with row-major [M][128] state they were 64-byte pieces 512 bytes apart and the cell update ran at 2.9 TB/s of HBM).
// HBM traffic per agent: 128 B features + 1 KB (h, c) in, 1 KB (h', c') + logits / value / action out -- one pass, against
// ~7.5 KB for the unfused gate GEMMs + cell update.  The B operand [W; U] (160 KB as fp16) stays resident in shared memory;
// the grid is persistent (one CTA per SM).
#include <cuda_fp16.h>

#include <cstdio>
#include <new>
#include <vector>

#include "ssd_internal.h"
#include "ssd_policy.h"
#include "ssd_umma.cuh"

#ifdef SSD_POLICY_TIMING  // profiles/micro/policy_head_timing.cu: cycles per phase of thread 0 of CTA 0
__device__ unsigned long long g_head_cycles[8];
#define SSD_HT_DECL unsigned long long ht_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long ht_t = clock64()
#define SSD_HT(slot) do { const long long n_ = clock64(); ht_[slot] += n_ - ht_t; ht_t = n_; } while (0)
#define SSD_HT_FLUSH do { if (blockIdx.x == 0 && threadIdx.x == 0) for (int q_ = 0; q_ < 8; ++q_) g_head_cycles[q_] = ht_[q_]; } while (0)
#else
#define SSD_HT_DECL
#define SSD_HT(slot)
#define SSD_HT_FLUSH
#endif

namespace ssd {
namespace policy_head {
using namespace ssd::umma;

constexpr int GA = 128, U = 128, KX = 32, K = KX + U, NG = 4 * U, NH = 16;
constexpr int kThreads = 256;
constexpr int kBBytes = (K / 8) * NG * 16;         // 163 840: [W; U] K-major, element (n, k) at ((k / 8) * 512 + n) * 16 + (k % 8) * 2
constexpr int kBHBytes = (U / 8) * NH * 16;        // 4 096: [logits_w | value_w | 0]
constexpr int kConstFloats = NG + NH;              // gate bias, head bias
constexpr int kBlobBytes = kBBytes + kBHBytes + kConstFloats * 4;   // 170 048
constexpr int kOffB = 0, kOffBH = kBBytes, kOffConst = kOffBH + kBHBytes;
constexpr int kOffA = (kOffConst + kConstFloats * 4 + 127) & ~127;   // [128 x 160] fp16, later fp16(h') [128 x 128]
constexpr int kABytes = (K / 8) * GA * 16;         // 40 960
constexpr int kOffBar = kOffA + kABytes;
constexpr int kSmemBytes = kOffBar + 16;
static_assert(kBlobBytes % 16 == 0 && kSmemBytes <= 227 * 1024, "one CTA per SM");

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// one MUFU op each (tanh.approx.f32, ~2^-11 relative error: below the fp16 rounding of the GEMM operands); the cell update
// is transcendental-bound: 5 per unit, 640 per agent
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

__global__ void __launch_bounds__(kThreads, 1)
lstm_heads_kernel(const float* __restrict__ feat, const float* h_in, const float* c_in, float* h_out, float* c_out, float* __restrict__ logits,
                  float* __restrict__ value, int8_t* __restrict__ actions, long long M, int num_outputs, const uint8_t* __restrict__ blob,
                  uint32_t seed_lo, uint32_t seed_hi, uint32_t counter) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    const float* s_bias = reinterpret_cast<const float*>(smem + kOffConst);
    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    for (int i = tid; i < (kBlobBytes >> 4); i += kThreads) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(blob)[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t sA = smem_u32(smem + kOffA), sB = smem_u32(smem + kOffB), sBH = smem_u32(smem + kOffBH);
    constexpr uint32_t kIG = umma_idesc(GA, 256), kIH = umma_idesc(GA, NH);
    uint32_t parity = 0;

    const long long n_groups = (M + GA - 1) / GA;
    SSD_HT_DECL;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long a0 = g * GA;
        const int rem = static_cast<int>(M - a0 < GA ? M - a0 : GA);
        {   // A = fp16([features | h]).  HBM is read in long contiguous runs (features: a 128-byte row per eight lanes; h: the tiled
            // state, 64 bytes per lane, 2 KB per warp -- a thread walking a row-major row in 32-byte steps reached 2.3 TB/s).
            // All loads before the stores: the compiler keeps a global load below a store through a generic pointer.
            const int seg = tid & 7, r8 = tid >> 3;          // features: 32 rows per pass, 4 passes, a 128-byte row per eight lanes
            const int hrow = tid & (GA - 1), hb0 = (tid >> 7) * 4;   // h: thread = (agent, four of the eight 16-unit blocks), 64 bytes per block
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 vx[4], vh[4][4];
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int row = ps * 32 + r8;
                vx[ps] = row < rem ? __ldcs(reinterpret_cast<const float4*>(feat + (a0 + row) * KX) + seg) : z;
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float4* src = reinterpret_cast<const float4*>(h_in + ((g * (U / 16) + hb0 + b) * GA + hrow) * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) vh[b][j] = hrow < rem ? __ldcs(src + j) : z;
            }
#pragma unroll
            for (int ps = 0; ps < 4; ++ps)   // elements 4 seg .. 4 seg + 3 of the 32 features: half of operand chunk seg / 2
                *reinterpret_cast<uint2*>(smem + kOffA + (seg >> 1) * (GA * 16) + (ps * 32 + r8) * 16 + (seg & 1) * 8) =
                    make_uint2(pack_h2(vx[ps].x, vx[ps].y), pack_h2(vx[ps].z, vx[ps].w));
#pragma unroll
            for (int b = 0; b < 4; ++b)      // units 16 (hb0 + b) .. + 15: operand chunks 4 + 2 (hb0 + b) and the next
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    *reinterpret_cast<uint4*>(smem + kOffA + (KX / 8 + 2 * (hb0 + b) + hf) * (GA * 16) + hrow * 16) =
                        make_uint4(pack_h2(vh[b][2 * hf].x, vh[b][2 * hf].y), pack_h2(vh[b][2 * hf].z, vh[b][2 * hf].w),
                                   pack_h2(vh[b][2 * hf + 1].x, vh[b][2 * hf + 1].y), pack_h2(vh[b][2 * hf + 1].z, vh[b][2 * hf + 1].w));
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        SSD_HT(0);
        if (warp == 0 && elect_one()) {  // gates: two column halves of 256, k-steps interleaved
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < K / 16; ++ks)
#pragma unroll
                for (int half = 0; half < 2; ++half)
                    umma_f16(tmem + half * 256, umma_desc(sA + ks * 2 * GA * 16, GA * 16, 128),
                             umma_desc(sB + half * 256 * 16 + ks * 2 * NG * 16, NG * 16, 128), kIG, ks > 0);
            umma_commit(bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        tc_fence_after();
        SSD_HT(1);
        {   // cell update: thread = agent (TMEM lane), warps 0-3 take units 0-63, warps 4-7 units 64-127
            const int q = warp & 3, row = q * 32 + lane, uh = warp >> 2;
            const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
            const bool live = row < rem;
            const long long tile0 = ((g * (U / 16) + uh * 4) * GA + row) * 16;   // this agent's 16 floats of unit block 4 uh; + GA * 16 per block
            float4 cn[4];   // the next 16 units of c, fetched one iteration ahead
#pragma unroll
            for (int e = 0; e < 4; ++e) cn[e] = live ? __ldcs(reinterpret_cast<const float4*>(c_in + tile0) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (int ub = 0; ub < 4; ++ub) {
                const int u0 = uh * 64 + ub * 16;
                uint32_t gi[16], gf[16], gg[16], go[16];
                tmem_ld16(trow + u0, gi);
                tmem_ld16(trow + U + u0, gf);
                tmem_ld16(trow + 2 * U + u0, gg);
                tmem_ld16(trow + 3 * U + u0, go);
                const float c[16] = {cn[0].x, cn[0].y, cn[0].z, cn[0].w, cn[1].x, cn[1].y, cn[1].z, cn[1].w,
                                     cn[2].x, cn[2].y, cn[2].z, cn[2].w, cn[3].x, cn[3].y, cn[3].z, cn[3].w};
                if (ub < 3) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        cn[e] = live ? __ldcs(reinterpret_cast<const float4*>(c_in + tile0 + (ub + 1) * (GA * 16)) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                tmem_ld_wait();
                float c2[16], hn[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float si = sigmoid_fast(__uint_as_float(gi[e]) + s_bias[u0 + e]), sf = sigmoid_fast(__uint_as_float(gf[e]) + s_bias[U + u0 + e]);
                    const float so = sigmoid_fast(__uint_as_float(go[e]) + s_bias[3 * U + u0 + e]);
                    c2[e] = sf * c[e] + si * tanh_fast(__uint_as_float(gg[e]) + s_bias[2 * U + u0 + e]);
                    hn[e] = so * tanh_fast(c2[e]);
                }
                if (live) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __stcs(reinterpret_cast<float4*>(c_out + tile0 + ub * (GA * 16)) + e, make_float4(c2[4 * e], c2[4 * e + 1], c2[4 * e + 2], c2[4 * e + 3]));
                        __stcs(reinterpret_cast<float4*>(h_out + tile0 + ub * (GA * 16)) + e, make_float4(hn[4 * e], hn[4 * e + 1], hn[4 * e + 2], hn[4 * e + 3]));
                    }
                }
                // fp16(h') is the A operand of the heads: the gate MMAs are complete, their operand buffer is free
                uint4* dst = reinterpret_cast<uint4*>(smem + kOffA + (u0 / 8) * (GA * 16) + row * 16);
                dst[0] = make_uint4(pack_h2(hn[0], hn[1]), pack_h2(hn[2], hn[3]), pack_h2(hn[4], hn[5]), pack_h2(hn[6], hn[7]));
                dst[GA] = make_uint4(pack_h2(hn[8], hn[9]), pack_h2(hn[10], hn[11]), pack_h2(hn[12], hn[13]), pack_h2(hn[14], hn[15]));
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();  // every gate column has been read: tensor memory can take the heads
        SSD_HT(2);
        if (warp == 0 && elect_one()) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < U / 16; ++ks)
                umma_f16(tmem, umma_desc(sA + ks * 2 * GA * 16, GA * 16, 128), umma_desc(sBH + ks * 2 * NH * 16, NH * 16, 128), kIH, ks > 0);
            umma_commit(bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        tc_fence_after();
        SSD_HT(3);
        if (warp < 4) {  // heads: logits, value, sampled action
            const int row = warp * 32 + lane;
            uint32_t y[16];
            tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16), y);
            tmem_ld_wait();
            if (row < rem) {
                const long long m = a0 + row;
                const float* hb = s_bias + NG;
                float best = -3.0e38f;
                int arg = 0;
                uint4 rnd = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int a = 0; a < NH - 1; ++a) {
                    if (a < num_outputs) {
                        const float l = __uint_as_float(y[a]) + hb[a];
                        logits[m * num_outputs + a] = l;
                        if (actions != nullptr) {  // Gumbel-max: argmax(l - log(-log u)), u uniform in (0, 1)
                            if ((a & 3) == 0) rnd = philox4x32_10(static_cast<uint32_t>(m), static_cast<uint32_t>(m >> 32), counter, a >> 2, seed_lo, seed_hi);
                            const float u = (static_cast<float>(pick_word(rnd, a & 3) >> 8) + 0.5f) * (1.0f / 16777216.0f);
                            const float s = l - __logf(-__logf(u));
                            if (s > best) { best = s; arg = a; }
                        }
                    }
                }
                // y[] is indexed statically: read the value column with a select chain
                float v = 0.f;
#pragma unroll
                for (int a = 0; a < NH; ++a) if (a == num_outputs) v = __uint_as_float(y[a]) + hb[a];
                value[m] = v;
                if (actions != nullptr) actions[m] = static_cast<int8_t>(arg);
            }
        }
        tc_fence_before();
        __syncthreads();  // tensor memory and the operand buffer are free for the next group
        tc_fence_after();
        SSD_HT(4);
    }
    SSD_HT_FLUSH;
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}

}  // namespace policy_head
}  // namespace ssd

#pragma GCC visibility push(default)
extern "C" {

int ssd_policy_set_head(ssd_policy_t p, int units, int num_outputs, const float* lstm_w, const float* lstm_u, const float* lstm_b,
                        const float* logits_w, const float* logits_b, const float* value_w, const float* value_b) {
    using namespace ssd::policy_head;
    if (!p || !lstm_w || !lstm_u || !lstm_b || !logits_w || !logits_b || !value_w || !value_b) return ssd::set_error(SSD_ERR_INVALID, "null argument");
    if (units != U) return ssd::set_error(SSD_ERR_UNSUPPORTED, "the fused LSTM kernel is built for cell_size 128 (conv_to_fcnet_v2.py's custom_options)");
    if (num_outputs < 1 || num_outputs > NH - 1) return ssd::set_error(SSD_ERR_UNSUPPORTED, "1..15 policy outputs");
    std::vector<uint8_t> blob(kBlobBytes, 0);
    __half* b = reinterpret_cast<__half*>(blob.data() + kOffB);
    __half* bh = reinterpret_cast<__half*>(blob.data() + kOffBH);
    float* cst = reinterpret_cast<float*>(blob.data() + kOffConst);
    auto at = [](int rows, int n, int k) { return ((k >> 3) * rows + n) * 8 + (k & 7); };  // canonical K-major, no swizzle
    for (int n = 0; n < NG; ++n) {   // Keras LSTM: kernel [32][4u], recurrent kernel [u][4u], gates i, f, c~, o
        for (int k = 0; k < KX; ++k) b[at(NG, n, k)] = __float2half_rn(lstm_w[k * NG + n]);
        for (int k = 0; k < U; ++k) b[at(NG, n, KX + k)] = __float2half_rn(lstm_u[k * NG + n]);
        cst[n] = lstm_b[n];
    }
    for (int k = 0; k < U; ++k) {
        for (int n = 0; n < num_outputs; ++n) bh[at(NH, n, k)] = __float2half_rn(logits_w[k * num_outputs + n]);
        bh[at(NH, num_outputs, k)] = __float2half_rn(value_w[k]);
    }
    for (int n = 0; n < num_outputs; ++n) cst[NG + n] = logits_b[n];
    cst[NG + num_outputs] = value_b[0];
    cudaError_t e = cudaSetDevice(p->device);
    if (e == cudaSuccess && !p->d_head_blob) e = cudaMalloc(&p->d_head_blob, kBlobBytes);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_head_blob, blob.data(), kBlobBytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(lstm_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return ssd::set_error(SSD_ERR_CUDA, cudaGetErrorString(e));
    p->units = units;
    p->num_outputs = num_outputs;
    return SSD_OK;
}

int ssd_policy_lstm_heads(ssd_policy_t p, const float* features, const float* h_in, const float* c_in, float* h_out, float* c_out, float* logits,
                          float* value, int8_t* actions, int64_t num_agents, uint64_t seed, uint32_t counter, void* stream) {
    using namespace ssd::policy_head;
    if (!p || !features || !h_in || !c_in || !h_out || !c_out || !logits || !value || num_agents < 0) return ssd::set_error(SSD_ERR_INVALID, "bad argument");
    if (!p->d_head_blob) return ssd::set_error(SSD_ERR_INVALID, "ssd_policy_set_head has not been called");
    const void* ptrs[5] = {features, h_in, c_in, h_out, c_out};
    for (const void* q : ptrs)
        if (reinterpret_cast<uintptr_t>(q) % 16 != 0) return ssd::set_error(SSD_ERR_INVALID, "features and state pointers must be 16-byte aligned");
    if (num_agents == 0) return SSD_OK;
    if (cudaSetDevice(p->device) != cudaSuccess) return ssd::set_error(SSD_ERR_CUDA, "cudaSetDevice failed");  // the blob and the stream live there
    const long long groups = (num_agents + GA - 1) / GA;
    const int grid = static_cast<int>(groups < p->sms ? groups : p->sms);
    lstm_heads_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
        features, h_in, c_in, h_out, c_out, logits, value, actions, num_agents, p->num_outputs, p->d_head_blob, static_cast<uint32_t>(seed),
        static_cast<uint32_t>(seed >> 32), counter);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ssd::set_error(SSD_ERR_CUDA, cudaGetErrorString(e));
    return SSD_OK;
}

}  // extern "C"
#pragma GCC visibility pop
