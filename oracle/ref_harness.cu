// ref_harness.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin C-ABI over the UNMODIFIED reference kernels, compiled from the sources where they lie under
// /root/reference (oracle/Makefile target `ref`; output only in oracle/_ref/, git-ignored).  Nothing from
// the reference is copied into this repository: this file only #includes the reference headers at build
// time and launches the reference's own __global__ functions in the order its dispatcher does.
//
//   launch order restated from Sort/OneSweepDispatcher.cuh:301-336 (keys) and :338-363 (pairs);
//   input generation from OneSweepDispatcher.cuh:100-104 (InitRandom<<<256,256>>>);
//   validation from OneSweepDispatcher.cuh:365-377 (Validate<<<ceil(n/4096),256>>>).
//
// The reference's dispatcher methods are private and own their buffers, hence this harness.
// Differences from the dispatcher, all on the safe side: descriptor arrays get tiles+1 slots (the
// reference writes slot tile+1 unconditionally), no cudaDeviceSynchronize inside the sort.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "Sort/OneSweep.cuh"
#include "UtilityKernels.cuh"

namespace {
constexpr uint32_t kRadix = 256, kPasses = 4, kPartSize = 7680, kHistPartSize = 65536;
constexpr uint32_t kHistThreads = 128, kBinThreads = 512;
inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

struct RefState {
    uint32_t max_n = 0;
    uint32_t *index = nullptr, *ghist = nullptr, *pass[4] = {nullptr, nullptr, nullptr, nullptr}, *err = nullptr;
};
}  // namespace

extern "C" {

void* ref_create(uint32_t max_n)
{
    RefState* s = new RefState();
    s->max_n = max_n;
    const size_t slots = (size_t)div_up(max_n, kPartSize) + 1;
    if (cudaMalloc(&s->index, kPasses * sizeof(uint32_t)) != cudaSuccess) return nullptr;
    if (cudaMalloc(&s->ghist, kPasses * kRadix * sizeof(uint32_t)) != cudaSuccess) return nullptr;
    for (int p = 0; p < 4; ++p)
        if (cudaMalloc(&s->pass[p], slots * kRadix * sizeof(uint32_t)) != cudaSuccess) return nullptr;
    if (cudaMalloc(&s->err, sizeof(uint32_t)) != cudaSuccess) return nullptr;
    return s;
}

void ref_destroy(void* h)
{
    RefState* s = (RefState*)h;
    if (!s) return;
    cudaFree(s->index); cudaFree(s->ghist); cudaFree(s->err);
    for (int p = 0; p < 4; ++p) cudaFree(s->pass[p]);
    delete s;
}

int ref_init_random_keys(uint32_t* d_sort, uint32_t n, uint32_t and_count, uint32_t seed)
{
    InitRandom<<<256, 256>>>(d_sort, and_count, seed, n);
    return (int)cudaGetLastError();
}

int ref_init_random_pairs(uint32_t* d_sort, uint32_t* d_payload, uint32_t n, uint32_t and_count, uint32_t seed)
{
    InitRandom<<<256, 256>>>(d_sort, d_payload, and_count, seed, n);
    return (int)cudaGetLastError();
}

static void ref_clear(RefState* s, uint32_t tiles)
{
    cudaMemsetAsync(s->index, 0, kPasses * sizeof(uint32_t));
    cudaMemsetAsync(s->ghist, 0, kRadix * kPasses * sizeof(uint32_t));
    for (int p = 0; p < 4; ++p) cudaMemsetAsync(s->pass[p], 0, (size_t)kRadix * (tiles + 1) * sizeof(uint32_t));
}

// sorts d_sort in place (result in d_sort), d_alt scratch; default stream, asynchronous.
int ref_sort_keys(void* h, uint32_t* d_sort, uint32_t* d_alt, uint32_t n)
{
    RefState* s = (RefState*)h;
    if (!s || n > s->max_n || n < 2) return -1;
    const uint32_t tiles = div_up(n, kPartSize);
    ref_clear(s, tiles);
    OneSweep::GlobalHistogram<<<div_up(n, kHistPartSize), kHistThreads>>>(d_sort, s->ghist, n);
    OneSweep::Scan<<<kPasses, kRadix>>>(s->ghist, s->pass[0], s->pass[1], s->pass[2], s->pass[3]);
    OneSweep::DigitBinningPassKeysOnly<<<tiles, kBinThreads>>>(d_sort, d_alt, s->pass[0], s->index, n, 0);
    OneSweep::DigitBinningPassKeysOnly<<<tiles, kBinThreads>>>(d_alt, d_sort, s->pass[1], s->index, n, 8);
    OneSweep::DigitBinningPassKeysOnly<<<tiles, kBinThreads>>>(d_sort, d_alt, s->pass[2], s->index, n, 16);
    OneSweep::DigitBinningPassKeysOnly<<<tiles, kBinThreads>>>(d_alt, d_sort, s->pass[3], s->index, n, 24);
    return (int)cudaGetLastError();
}

int ref_sort_pairs(void* h, uint32_t* d_sort, uint32_t* d_pay, uint32_t* d_alt, uint32_t* d_alt_pay, uint32_t n)
{
    RefState* s = (RefState*)h;
    if (!s || n > s->max_n || n < 2) return -1;
    const uint32_t tiles = div_up(n, kPartSize);
    ref_clear(s, tiles);
    OneSweep::GlobalHistogram<<<div_up(n, kHistPartSize), kHistThreads>>>(d_sort, s->ghist, n);
    OneSweep::Scan<<<kPasses, kRadix>>>(s->ghist, s->pass[0], s->pass[1], s->pass[2], s->pass[3]);
    OneSweep::DigitBinningPassPairs<<<tiles, kBinThreads>>>(d_sort, d_pay, d_alt, d_alt_pay, s->pass[0], s->index, n, 0);
    OneSweep::DigitBinningPassPairs<<<tiles, kBinThreads>>>(d_alt, d_alt_pay, d_sort, d_pay, s->pass[1], s->index, n, 8);
    OneSweep::DigitBinningPassPairs<<<tiles, kBinThreads>>>(d_sort, d_pay, d_alt, d_alt_pay, s->pass[2], s->index, n, 16);
    OneSweep::DigitBinningPassPairs<<<tiles, kBinThreads>>>(d_alt, d_alt_pay, d_sort, d_pay, s->pass[3], s->index, n, 24);
    return (int)cudaGetLastError();
}

// copies the global histogram [4][256] computed by the last sort to host (for kernel-level parity).
int ref_get_global_histogram(void* h, uint32_t* host_out)
{
    RefState* s = (RefState*)h;
    return (int)cudaMemcpy(host_out, s->ghist, kPasses * kRadix * sizeof(uint32_t), cudaMemcpyDeviceToHost);
}

// returns the reference's error count (adjacent inversions), or -1 on CUDA error.
long long ref_validate_keys(void* h, uint32_t* d_sort, uint32_t n)
{
    RefState* s = (RefState*)h;
    cudaMemset(s->err, 0, sizeof(uint32_t));
    Validate<<<div_up(n, 4096), 256>>>(d_sort, s->err, n);
    uint32_t e = 0;
    if (cudaMemcpy(&e, s->err, sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)e;
}

long long ref_validate_pairs(void* h, uint32_t* d_sort, uint32_t* d_pay, uint32_t n)
{
    RefState* s = (RefState*)h;
    cudaMemset(s->err, 0, sizeof(uint32_t));
    Validate<<<div_up(n, 4096), 256>>>(d_sort, d_pay, s->err, n);
    uint32_t e = 0;
    if (cudaMemcpy(&e, s->err, sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (long long)e;
}

// The reference's timing protocol (OneSweepDispatcher.cuh:193-239): per iteration InitRandom(seed+i), then
// cudaEvent around the dispatch (memsets included); iteration 0 is a discarded warm-up.  Returns total ms of
// `iters` timed sorts, or <0 on error.
float ref_batch_timing_keys(void* h, uint32_t* d_sort, uint32_t* d_alt, uint32_t n, uint32_t iters, uint32_t seed)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float total = 0.f;
    for (uint32_t i = 0; i <= iters; ++i) {
        InitRandom<<<256, 256>>>(d_sort, 0, i + seed, n);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        if (ref_sort_keys(h, d_sort, d_alt, n) != 0) return -1.f;
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
        if (i) total += ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return cudaGetLastError() == cudaSuccess ? total : -1.f;
}

float ref_batch_timing_pairs(void* h, uint32_t* d_sort, uint32_t* d_pay, uint32_t* d_alt, uint32_t* d_alt_pay,
                             uint32_t n, uint32_t iters, uint32_t seed)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float total = 0.f;
    for (uint32_t i = 0; i <= iters; ++i) {
        InitRandom<<<256, 256>>>(d_sort, d_pay, 0, i + seed, n);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        if (ref_sort_pairs(h, d_sort, d_pay, d_alt, d_alt_pay, n) != 0) return -1.f;
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
        if (i) total += ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return cudaGetLastError() == cudaSuccess ? total : -1.f;
}

}  // extern "C"
