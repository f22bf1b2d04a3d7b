// Which store flavour keeps NVLink efficient when a kernel writes 32-key (128 B) runs into a PEER GPU's memory?
// Single process, 2 GPUs, cudaDeviceEnablePeerAccess.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// mode 0: st.global (default)   1: st.global.cs   2: st.global.cg   3: 16 B vector stores (aligned only)
// each warp writes RUN consecutive u32 starting at a pseudo-random (optionally misaligned) offset
template <int MODE>
__global__ void __launch_bounds__(512) runs(uint32_t* __restrict__ dst, size_t n, int run, int mis, int iters)
{
    const int lane = threadIdx.x & 31;
    uint32_t s = (blockIdx.x * 16 + (threadIdx.x >> 5)) * 2654435761u + 7;
    const size_t slots = n / 4096;
    for (int it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        size_t base = (size_t)(s % slots) * 4096 + ((s >> 22) & 31) * 32 + (mis ? 1 + ((s >> 27) & 1) * 2 : 0);
        if (MODE == 3) {
            for (int i = lane * 4; i < run; i += 128) *reinterpret_cast<uint4*>(dst + base + i) = make_uint4(s, it, lane, i);
        } else {
            for (int i = lane; i < run; i += 32) {
                uint32_t* p = dst + base + i;
                if (MODE == 0) *p = s + i;
                if (MODE == 1) __stcs(p, s + i);
                if (MODE == 2) __stcg(p, s + i);
            }
        }
    }
}

template <int MODE>
void bench(const char* name, uint32_t* dst, size_t n, int run, int mis)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = 148 * 4, iters = 2000;
    runs<MODE><<<grid, 512>>>(dst, n, run, mis, 10);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    runs<MODE><<<grid, 512>>>(dst, n, run, mis, iters);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    const double bytes = (double)grid * 16 * iters * run * 4;
    printf("%-28s run=%3d keys mis=%d : %8.3f ms  %8.1f GB/s\n", name, run, mis, ms, bytes / ms / 1e6);
}

int main()
{
    int nd; CK(cudaGetDeviceCount(&nd));
    const size_t n = (size_t)1 << 29;  // 2 GiB of u32 per buffer
    uint32_t *local, *peer;
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, n * 4));
    if (nd < 2) { printf("single GPU: local only\n"); peer = nullptr; }
    else {
        CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, n * 4));
        CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
    }
    for (int where = 0; where < (peer ? 2 : 1); ++where) {
        uint32_t* dst = where ? peer : local;
        printf("---- destination: %s\n", where ? "PEER (NVLink)" : "local HBM");
        for (int mis = 0; mis < 2; ++mis) for (int run : {32, 64, 256}) {
            bench<0>("st.global", dst, n, run, mis);
            bench<1>("st.global.cs", dst, n, run, mis);
            bench<2>("st.global.cg", dst, n, run, mis);
            if (!mis) bench<3>("st.global.v4 (16 B/lane)", dst, n, run, mis);
        }
    }
    return 0;
}
