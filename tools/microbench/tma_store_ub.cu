// How fast can an SM issue small cp.async.bulk shared->global copies (SASS UBLKCP.G.S)?  Decides whether the
// digit-run scatter of the binning pass can be handed to the TMA engine.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 2) bulk_store(uint8_t* out, size_t out_bytes, int bytes, int iters, int issuers, long long* clk)
{
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x < issuers) {
        uint32_t s = (blockIdx.x * 512 + threadIdx.x) * 2654435761u + 1;
        const size_t slots = out_bytes / 2048;
        for (int it = 0; it < iters; ++it) {
            s = s * 1664525u + 1013904223u;
            uint8_t* dst = out + (size_t)(s % slots) * 2048 + ((s >> 20) & 7) * 16;   // 16B-aligned, pseudo-random
            const uint32_t src = smem_addr(sm + ((threadIdx.x * 1024 + it * 16) & 0xfff0 & ~(0)) % (65536 - 1024));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src & ~15u), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if ((it & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// same traffic with ordinary stores: 32 lanes write 128 B rows (coalesced), for comparison
__global__ void __launch_bounds__(512, 2) plain_store(uint8_t* out, size_t out_bytes, int bytes, int iters, long long* clk)
{
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u;
    __syncthreads();
    long long t0 = clock64();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t s = (blockIdx.x * 16 + warp) * 2654435761u + 1;
    const size_t slots = out_bytes / 2048;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        uint32_t* dst = (uint32_t*)(out + (size_t)(s % slots) * 2048 + ((s >> 20) & 7) * 16 + 4 * ((s >> 25) & 3));
        for (int b = lane * 4; b < bytes; b += 128) dst[b / 4] = ((uint32_t*)sm)[(warp * 1024 + b) / 4];
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t out_bytes = (size_t)4 << 30;
    uint8_t* out; CK(cudaMalloc(&out, out_bytes));
    long long* clk; CK(cudaMalloc(&clk, sizeof(long long) * sms * 2));
    CK(cudaFuncSetAttribute(bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(plain_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = sms * 2;
    for (int issuers : {32, 256}) for (int bytes : {64, 128, 256, 512, 1024}) {
        const int iters = 4096 / (issuers >= 32 ? 8 : 1);
        bulk_store<<<grid, 512, 65536>>>(out, out_bytes, bytes, 16, issuers, clk);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        bulk_store<<<grid, 512, 65536>>>(out, out_bytes, bytes, iters, issuers, clk);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)grid * issuers * iters;
        printf("bulk S->G  issuers/CTA=%3d bytes=%4d : %.3f ms  %.1f Mops/s/SM  %.1f clk/op/SM(@1.9GHz)  %.0f GB/s\n", issuers, bytes, ms,
               ops / ms / 1e3 / sms, ms * 1e-3 * 1.9e9 / (ops / sms), ops * bytes / ms / 1e6);
    }
    for (int bytes : {128, 256, 512, 1024}) {
        const int iters = 2048;
        plain_store<<<grid, 512, 65536>>>(out, out_bytes, bytes, 16, clk);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        plain_store<<<grid, 512, 65536>>>(out, out_bytes, bytes, iters, clk);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)grid * 16 * iters;
        printf("plain STG  warp-runs bytes=%4d : %.3f ms  %.1f Mruns/s/SM  %.0f GB/s\n", bytes, ms, ops / ms / 1e3 / sms, ops * bytes / ms / 1e6);
    }
    return 0;
}
