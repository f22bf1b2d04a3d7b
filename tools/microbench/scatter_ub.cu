// Memory-system ceiling of the radix-scatter ACCESS PATTERN with no ranking work at all: every CTA reads a 16,384-key
// tile with 32 coalesced loads per thread in flight and writes it as NB runs (one per "digit") into NB output streams,
// exactly the pattern of a DigitBinningPass.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int THREADS = 512, K = 32, T = THREADS * K;

__global__ void __launch_bounds__(THREADS, 2) pattern(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, int nb, int mis)
{
    const size_t tiles = n / T;
    const int run = T / nb;            // keys per run (>= 32 for nb <= 512)
    const size_t stream = n / nb;      // keys per output stream
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint32_t key[K];
        const uint32_t* src = in + t * T + warp * (32 * K) + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) key[i] = __ldcs(src + i * 32);
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int idx = warp * (32 * K) + i * 32 + lane;   // position inside the tile
            const int b = idx / run, r = idx % run;
            size_t pos = (size_t)b * stream + t * run + r + (mis ? (size_t)((b * 7 + 3) % 8) * mis : 0);
            if (pos >= n) pos -= n;
            __stcs(out + pos, key[i]);
        }
    }
}

int main()
{
    const size_t n = (size_t)1 << 30;
    uint32_t *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 1, n * 4)); CK(cudaMemset(b, 2, n * 4));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int grid_mult : {2, 443}) for (int mis : {0, 1}) for (int nb : {1, 16, 64, 256, 512}) {
        const int grid = grid_mult == 2 ? sms * 2 : (int)(n / T);   // persistent vs one CTA per tile
        pattern<<<grid, THREADS>>>(a, b, n, nb, mis); CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0)); pattern<<<grid, THREADS>>>(a, b, n, nb, mis); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("%s streams=%4d run=%5d keys mis=%d : %.3f ms  %.0f GB/s (r+w)\n", grid_mult == 2 ? "persistent  " : "cta-per-tile", nb, T / nb, mis, best, 8.0 * n / best / 1e6);
    }
    return 0;
}
