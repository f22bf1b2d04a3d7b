// main_cli.cpp -- the reference's test/benchmark driver (GPUSortingCUDA/GPUSortingCUDA.cu:16-57) for the B200 path:
// TestAllKeysOnly / TestAllPairs sweeps, then BatchTiming at 2^28 with 100 iterations, seed 10 -- through the C-ABI.
// Build: make -C gpusorting_b200/csrc cli (also done by __graft_entry__.build()).
// Run (GPU box): gpusorting_b200/lib/onesweep_b200_cli [log2n=28] [iters=100] [sweep_step=1] [max_log2=28]
// (the reference's main is the defaults; tests/test_gpu_parity.py::test_cli_runs_the_reference_protocol uses a coarse sweep)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/OneSweepB200.hpp"

static uint32_t g_sweep_step = 1, g_max_log2 = 28;

static bool test_sweep(bool pairs, uint32_t max_n)
{
    OneSweepSorterB200 s(max_n, 4, pairs ? 4 : 0);
    uint32_t *keys = nullptr, *vals = nullptr;
    cudaMalloc(&keys, (size_t)max_n * 4);
    if (pairs) cudaMalloc(&vals, (size_t)max_n * 4);
    unsigned passed = 0, total = 0;
    auto one = [&](uint32_t n, uint32_t seed) {
        osb200_init_random_u32(keys, vals, n, 0, seed, 0, nullptr);
        if (pairs) s.SortPairs(keys, vals, n); else s.SortKeys(keys, n);
        bool ok = s.Validate(keys, n) == 0 && (!pairs || s.Validate(vals, n) == 0);  // payload == key, as the reference tests
        passed += ok; ++total;
        if (!ok) printf("\n Test failed at size %u \n", n);
    };
    for (uint32_t n = 7680; n <= 15360; n += g_sweep_step) { one(n, n); if (!(n & 255)) { printf("."); fflush(stdout); } }
    for (uint32_t e = g_max_log2 >= 28 ? 26 : g_max_log2; e <= g_max_log2 && (1u << e) <= max_n; ++e) one(1u << e, e);
    printf("\n%u/%u %s\n\n", passed, total, passed == total ? "All tests passed." : "Test failed.");
    cudaFree(keys); cudaFree(vals);
    return passed == total;
}

static void batch_timing(bool pairs, uint32_t size, uint32_t batch, uint32_t seed)
{
    OneSweepSorterB200 s(size, 4, pairs ? 4 : 0);
    uint32_t *keys = nullptr, *vals = nullptr;
    cudaMalloc(&keys, (size_t)size * 4);
    if (pairs) cudaMalloc(&vals, (size_t)size * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float total = 0.f;
    for (uint32_t i = 0; i <= batch; ++i) {
        osb200_init_random_u32(keys, vals, size, 0, i + seed, 0, nullptr);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        if (pairs) s.SortPairs(keys, vals, size); else s.SortKeys(keys, size);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (i) total += ms;
    }
    printf("Total time elapsed: %f\nEstimated speed at %u 32-bit elements: %E %s/sec\n\n", total / 1000.f, size,
           size / (total / 1000.f) * batch, pairs ? "pairs" : "keys");
    cudaFree(keys); cudaFree(vals);
}

int main(int argc, char** argv)
{
    const uint32_t log2n = argc > 1 ? atoi(argv[1]) : 28, iters = argc > 2 ? atoi(argv[2]) : 100;
    if (argc > 3) g_sweep_step = atoi(argv[3]) > 0 ? atoi(argv[3]) : 1;
    if (argc > 4) g_max_log2 = atoi(argv[4]);
    const uint32_t cap = 1u << (g_max_log2 > log2n ? g_max_log2 : log2n);
    try {
        printf("Beginning B200 OneSweep keys validation test: \n");
        bool ok = test_sweep(false, cap);
        batch_timing(false, 1u << log2n, iters, 10);
        printf("Beginning B200 OneSweep pairs validation test: \n");
        ok = test_sweep(true, cap) && ok;
        batch_timing(true, 1u << log2n, iters, 10);
        return ok ? 0 : 1;
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
