// Memory microbenchmarks: what fraction of the copy peak does a radix-scatter write pattern reach on B200?
// Not part of the product path.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void copy_v4(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n4)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
    }
    for (; i < n4; i += stride) out[i] = in[i];
}

__global__ void read_v4(const uint4* __restrict__ in, uint32_t* __restrict__ out, size_t n4)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n4; i += stride) { uint4 a = in[i]; acc += a.x ^ a.y ^ a.z ^ a.w; }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void write_v4(uint4* __restrict__ out, size_t n4)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) out[i] = make_uint4(i, 1, 2, 3);
}

// Emulated radix scatter: tile t (TILE keys, read coalesced) is written as 256 runs of RUN=TILE/256 keys;
// run b of tile t lands at out[b*(n/256) + t*RUN + shift(b)] where shift de-aligns the runs by `mis` keys.
template <int TILE>
__global__ void scatter_runs(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, int mis)
{
    constexpr int RUN = TILE / 256;
    const size_t ntiles = n / TILE;
    const size_t binsz = n / 256;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t* src = in + t * TILE;
        for (int i = threadIdx.x; i < TILE; i += blockDim.x) {
            const uint32_t v = src[i];
            const int b = i / RUN, r = i % RUN;
            size_t pos = (size_t)b * binsz + t * RUN + r + (mis ? ((b * 7 + 3) % 8) * (size_t)mis : 0);
            if (pos >= n) pos -= n;
            out[pos] = v;
        }
    }
}

template <typename F>
float timeit(F f, int reps = 5)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    const size_t n = (size_t)1 << 30;
    uint32_t *a, *b;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    CK(cudaMemset(a, 1, n * 4)); CK(cudaMemset(b, 2, n * 4));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const double gb = n * 4 / 1e9;
    for (int mult : {4, 8, 16}) for (int thr : {256, 512, 1024}) {
        float ms = timeit([&] { copy_v4<<<sms * mult, thr>>>((uint4*)a, (uint4*)b, n / 4); });
        printf("copy_v4   grid=%dxSM thr=%d  %.3f ms  %.1f GB/s (r+w)\n", mult, thr, ms, 2 * gb / ms * 1e3);
    }
    { float ms = timeit([&] { CK(cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice)); });
      printf("cudaMemcpy D2D            %.3f ms  %.1f GB/s (r+w)\n", ms, 2 * gb / ms * 1e3); }
    { float ms = timeit([&] { read_v4<<<sms * 8, 512>>>((uint4*)a, b, n / 4); });
      printf("read_v4                   %.3f ms  %.1f GB/s (r)\n", ms, gb / ms * 1e3); }
    { float ms = timeit([&] { write_v4<<<sms * 8, 512>>>((uint4*)b, n / 4); });
      printf("write_v4                  %.3f ms  %.1f GB/s (w)\n", ms, gb / ms * 1e3); }
    { float ms = timeit([&] { CK(cudaMemsetAsync(b, 0, n * 4)); });
      printf("cudaMemset                %.3f ms  %.1f GB/s (w)\n", ms, gb / ms * 1e3); }
    for (int mis : {0, 1, 3}) {
        { float ms = timeit([&] { scatter_runs<4096><<<sms * 4, 512>>>(a, b, n, mis); });
          printf("scatter tile=4096 run=16 mis=%d   %.3f ms  %.1f GB/s (r+w)\n", mis, ms, 2 * gb / ms * 1e3); }
        { float ms = timeit([&] { scatter_runs<8192><<<sms * 4, 512>>>(a, b, n, mis); });
          printf("scatter tile=8192 run=32 mis=%d   %.3f ms  %.1f GB/s (r+w)\n", mis, ms, 2 * gb / ms * 1e3); }
        { float ms = timeit([&] { scatter_runs<16384><<<sms * 4, 512>>>(a, b, n, mis); });
          printf("scatter tile=16384 run=64 mis=%d  %.3f ms  %.1f GB/s (r+w)\n", mis, ms, 2 * gb / ms * 1e3); }
        { float ms = timeit([&] { scatter_runs<32768><<<sms * 4, 512>>>(a, b, n, mis); });
          printf("scatter tile=32768 run=128 mis=%d %.3f ms  %.1f GB/s (r+w)\n", mis, ms, 2 * gb / ms * 1e3); }
    }
    return 0;
}
