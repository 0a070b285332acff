// How long does a tcgen05.mma (kind::f16, M = 128, K = 16) take when it accumulates into the tile of its predecessor, against
// MMAs that rotate over independent accumulators?  One CTA, one issuing thread, operands are whatever is in memory.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../sequential_social_dilemma_games_b200/csrc -o umma_chain umma_chain.cu
#include <cstdio>
#include "ssd_umma.cuh"

using namespace ssd;
using namespace ssd::umma;

template <int N, int CHAINS, bool A_TMEM>
__global__ void chain_kernel(int n_mma, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp == 0 && elect_one()) {
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 8192);
        constexpr uint32_t idesc = umma_idesc(128, N);
        const uint64_t a_desc = umma_desc(sA, 2048, 128), b_desc = umma_desc(sB, N * 16, 128);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const uint32_t d = tmem + (i % CHAINS) * 96;   // accumulators 96 columns apart (N <= 96), A operand above them
                if (A_TMEM) umma_f16_ts(d, tmem + 448, b_desc, idesc, i >= CHAINS);
                else umma_f16(d, a_desc, b_desc, idesc, i >= CHAINS);
            }
            umma_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            out[0] = t1 - t0;   // issue
            out[1] = t2 - t0;   // issue -> all complete (barrier seen by the issuing thread)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}

template <int N, int CHAINS, bool A_TMEM>
void run(const char* what) {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(chain_kernel<N, CHAINS, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    printf("%-44s", what);
    for (int n : {1, 4, 8, 16, 32, 64}) {
        chain_kernel<N, CHAINS, A_TMEM><<<1, 128, 65536>>>(n, d);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("  n=%-2d %5lld (%4.0f/MMA)", n, h[1], (double)h[1] / n);
    }
    printf("   [%s]\n", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    printf("cycles from first issue to completion seen; M = 128, K = 16, kind::f16\n");
    run<80, 1, true>("N=80  A in TMEM  1 chain (dependent)");
    run<80, 2, true>("N=80  A in TMEM  2 chains interleaved");
    run<80, 4, true>("N=80  A in TMEM  4 chains interleaved");
    run<80, 1, false>("N=80  A in smem  1 chain (dependent)");
    run<80, 4, false>("N=80  A in smem  4 chains interleaved");
    run<32, 1, true>("N=32  A in TMEM  1 chain (dependent)");
    run<32, 4, true>("N=32  A in TMEM  4 chains interleaved");
    run<16, 1, true>("N=16  A in TMEM  1 chain (dependent)");
    return 0;
}
