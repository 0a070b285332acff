// Microbenchmark: what HBM bandwidth does a B200 sustain for write-only / read-only / copy / an 18:82 read:write mix
// (the step kernel's DRAM mix), with plain 16-byte stores and with TMA bulk stores of 1440-byte chunks from shared memory?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hbm_mix hbm_mix.cu ; run: ./hbm_mix
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void write_k(uint4* dst, size_t n) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void read_k(const uint4* src, size_t n, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void copy_k(const uint4* src, uint4* dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// reads n_r uint4, writes n_w uint4 (interleaved per thread in proportion)
__global__ void mix_k(const uint4* src, uint4* dst, size_t n_w, int ratio, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_w; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = make_uint4(i, 2, 3, 4);
        if ((i / 32) % ratio == 0) { uint4 r = src[(i / 32 / ratio) * 32 + (i & 31)]; acc ^= r.x; }
        dst[i] = v;
    }
    if (acc == 0x12345678u) *sink = acc;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// one warp = one 1440-byte staging buffer; fill it, bulk-store it, wait for the read, repeat
__global__ void bulk_write_k(uint8_t* dst, size_t n_chunks, int chunk) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    uint8_t* st = sm + warp * ((chunk + 127) & ~127);
    for (size_t c = blockIdx.x * (size_t)nw + warp; c < n_chunks; c += (size_t)gridDim.x * nw) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int i = lane * 4; i < chunk; i += 128) *reinterpret_cast<uint32_t*>(st + i) = (uint32_t)c + i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + c * chunk), "r"(smem_u32(st)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const size_t bytes = 4ull << 30;  // 4 GiB per buffer
    uint8_t *a, *b; uint32_t* sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t n = bytes / 16;
    const int grid = 148 * 16, thr = 256;
    float ms;
#define TIME(label, moved, ...) do { for (int w = 0; w < 2; ++w) { __VA_ARGS__; } cudaEventRecord(e0); for (int r = 0; r < 5; ++r) { __VA_ARGS__; } cudaEventRecord(e1); \
        CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); printf("%-44s %8.1f GB/s\n", label, (double)(moved) * 5 / ms / 1e6); } while (0)
    TIME("cudaMemsetAsync 4 GiB (write only)", bytes, cudaMemsetAsync(a, 3, bytes));
    TIME("st.v4 write only", bytes, (write_k<<<grid, thr>>>((uint4*)a, n)));
    TIME("ld.v4 read only", bytes, (read_k<<<grid, thr>>>((const uint4*)a, n, sink)));
    TIME("copy (read + write, both counted)", 2 * bytes, (copy_k<<<grid, thr>>>((const uint4*)a, (uint4*)b, n)));
    TIME("cudaMemcpyAsync D2D (both counted)", 2 * bytes, cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
    TIME("mix 1 read : 4 write (both counted)", bytes + bytes / 4, (mix_k<<<grid, thr>>>((const uint4*)a, (uint4*)b, n, 4, sink)));
    TIME("mix 1 read : 5 write (both counted)", bytes + bytes / 5, (mix_k<<<grid, thr>>>((const uint4*)a, (uint4*)b, n, 5, sink)));
    for (int chunk : {1440, 4096}) {
        const size_t nc = bytes / chunk;
        for (int cps : {8, 16}) {
            const int smem = 4 * ((chunk + 127) & ~127);
            cudaFuncSetAttribute(bulk_write_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            char label[96]; snprintf(label, sizeof label, "bulk store %d B chunks, %d CTAs/SM x 4 warps", chunk, cps);
            TIME(label, nc * chunk, (bulk_write_k<<<148 * cps, 128, smem>>>(b, nc, chunk)));
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
