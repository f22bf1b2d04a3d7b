// Microbenchmarks that decide the ranking design of the digit-binning kernel on B200.
// Measures warp-level multisplit variants in isolation (keys in registers, warp-private
// shared-memory histograms), reporting keys per SM-clock.  Not part of the product path.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int K = 16;        // keys per lane per tile
constexpr int WARPS = 16;    // warps per CTA
constexpr int THREADS = WARPS * 32;

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

__device__ __forceinline__ uint32_t match_ballot8(uint32_t d) {
    uint32_t mask = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool p = (d >> b) & 1;
        const uint32_t bal = __ballot_sync(0xffffffffu, p);
        mask &= p ? bal : ~bal;
    }
    return mask;
}

__device__ __forceinline__ uint32_t match_hw(uint32_t d) {
    return __match_any_sync(0xffffffffu, d);
}

// variant 0: ballots + leader atomicAdd (returning) + shfl
// variant 1: match.any + leader atomicAdd + shfl
// variant 2: match.any + leader LDS/STS + shfl
// variant 3: ballots + leader LDS/STS + shfl
// variant 4: match.any only (throughput of MATCH)
// variant 5: ballots only
// variant 6: every lane atomicAdd (returning), no match (unstable; throughput probe of ATOMS)
// variant 7: every lane red (non-returning atomic) probe
template <int V>
__global__ void __launch_bounds__(THREADS) rank_kernel(uint32_t* out, int iters, uint32_t seed, long long* clk)
{
    __shared__ uint32_t hist[WARPS * 256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* wh = hist + warp * 256;
    for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) hist[i] = 0;
    __syncthreads();
    uint32_t keys[K];
    uint32_t s = seed ^ (blockIdx.x * THREADS + threadIdx.x) * 2654435761u;
#pragma unroll
    for (int i = 0; i < K; ++i) { s = s * 1664525u + 1013904223u; keys[i] = s >> 8; }
    uint32_t acc = 0;
    const uint32_t lt = lanemask_lt();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int shift = (it & 1) * 8;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const uint32_t d = (keys[i] >> shift) & 255u;
            if (V == 0 || V == 1 || V == 2 || V == 3) {
                const uint32_t m = (V == 0 || V == 3) ? match_ballot8(d) : match_hw(d);
                const uint32_t below = __popc(m & lt);
                uint32_t pre = 0;
                if (below == 0) {
                    if (V == 0 || V == 1) pre = atomicAdd(&wh[d], __popc(m));
                    else { pre = wh[d]; wh[d] = pre + __popc(m); }
                }
                pre = __shfl_sync(0xffffffffu, pre, __ffs(m) - 1);
                acc += pre + below;
            } else if (V == 4) {
                acc += match_hw(d);
            } else if (V == 5) {
                acc += match_ballot8(d);
            } else if (V == 6) {
                acc += atomicAdd(&wh[d], 1u);
            } else if (V == 7) {
                atomicAdd(&wh[d], 1u);
            }
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (V == 7) acc += wh[lane];
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int ctas_per_sm, int iters)
{
    int dev = 0, sms = 0; CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * ctas_per_sm;
    uint32_t* out; long long* clk;
    CK(cudaMalloc(&out, sizeof(uint32_t) * grid * THREADS));
    CK(cudaMalloc(&clk, sizeof(long long) * grid));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    rank_kernel<V><<<grid, THREADS>>>(out, 10, 1u, clk);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    rank_kernel<V><<<grid, THREADS>>>(out, iters, 7u, clk);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* h = (long long*)malloc(sizeof(long long) * grid);
    CK(cudaMemcpy(h, clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double keys_per_cta = (double)iters * K * THREADS;
    // keys per SM-clock: ctas_per_sm CTAs share the SM for ~avg clocks
    printf("%-34s ctas/sm=%d  %.3f keys/clk/SM  (%.1f clk per warp-round)  %.3f ms  => %.1f Gkeys/s chip\n",
           name, ctas_per_sm, keys_per_cta * ctas_per_sm / avg, avg / ((double)iters * K * WARPS * ctas_per_sm) ,
           ms, keys_per_cta * grid / (ms * 1e6));
    free(h); CK(cudaFree(out)); CK(cudaFree(clk));
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int c = 1; c <= 4; c *= 2) {
        run<0>("ballot8+ATOMS+shfl", c, 2000);
        run<1>("match.any+ATOMS+shfl", c, 2000);
        run<2>("match.any+LDS/STS+shfl", c, 2000);
        run<3>("ballot8+LDS/STS+shfl", c, 2000);
        run<4>("match.any only", c, 2000);
        run<5>("ballot8 only", c, 2000);
        run<6>("ATOMS.ADD ret all lanes", c, 2000);
        run<7>("ATOMS.ADD noret all lanes", c, 2000);
    }
    return 0;
}
