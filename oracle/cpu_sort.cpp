// cpu_sort.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.c header).
//
// The reference has no CPU sort; BASELINE.json's north_star asks for "CPU std::sort on the
// box's host cores" as the reported baseline and bit-exact check, so this file provides it:
// std::sort (keys), std::stable_sort by key (pairs -- the unique stable answer an LSD radix sort with
// in-order ranking must produce, OneSweep.cu:207-253), and libstdc++ parallel-mode sort on all cores.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>
#include <parallel/algorithm>
#include <omp.h>

extern "C" {

void orc_std_sort_u32(uint32_t* keys, uint64_t n) { std::sort(keys, keys + n); }
void orc_std_sort_u64(uint64_t* keys, uint64_t n) { std::sort(keys, keys + n); }

// all-core sort; threads<=0 -> omp default
void orc_parallel_sort_u32(uint32_t* keys, uint64_t n, int threads)
{
    if (threads > 0) omp_set_num_threads(threads);
    __gnu_parallel::sort(keys, keys + n);
}
void orc_parallel_sort_u64(uint64_t* keys, uint64_t n, int threads)
{
    if (threads > 0) omp_set_num_threads(threads);
    __gnu_parallel::sort(keys, keys + n);
}

// stable sort of (key, value) pairs by key only
void orc_std_stable_sort_pairs_u32(uint32_t* keys, uint32_t* vals, uint64_t n)
{
    std::vector<uint64_t> kv(n);
    for (uint64_t i = 0; i < n; ++i) kv[i] = (uint64_t(keys[i]) << 32) | uint64_t(i);
    std::sort(kv.begin(), kv.end());  // (key, original index) ascending == stable by key
    std::vector<uint32_t> v(vals, vals + n);
    for (uint64_t i = 0; i < n; ++i) { keys[i] = uint32_t(kv[i] >> 32); vals[i] = v[uint32_t(kv[i])]; }
}

}  // extern "C"
