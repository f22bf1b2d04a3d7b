// OneSweepB200.hpp -- header-only C++ mirror of the reference's interface on top of the C-ABI.
//
//   namespace OneSweep { Sort(keys[, values], n[, stream]) }   the north-star call shape: the only public `Sort` of the
//       reference is GPUSortingUnity/Runtime/OneSweep.cs:297-306,358-370; in CUDA, OneSweep is a namespace of kernels
//       (GPUSortingCUDA/Sort/OneSweep.cuh:22-53) driven by private dispatcher methods.
//   class OneSweepSorterB200                                    RAII owner of one osb200 handle
//       (reference: OneSweepDispatcher ctor/dtor, Sort/OneSweepDispatcher.cuh:42-83).
// Errors: std::runtime_error carrying osb200_status_string() -- the reference ignores CUDA errors entirely.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>

#include "onesweep_b200.h"

class OneSweepSorterB200 {
  public:
    OneSweepSorterB200(uint64_t max_n, int key_bytes, int value_bytes) : max_n_(max_n)
    {
        check(osb200_create(&h_, max_n, key_bytes, value_bytes), "osb200_create");
    }
    ~OneSweepSorterB200() { if (h_) osb200_destroy(h_); }
    OneSweepSorterB200(const OneSweepSorterB200&) = delete;
    OneSweepSorterB200& operator=(const OneSweepSorterB200&) = delete;

    void SortKeys(uint32_t* d_keys, uint64_t n, void* stream = nullptr) { check(osb200_sort_keys_u32(h_, d_keys, n, stream), "osb200_sort_keys_u32"); }
    void SortKeys(uint64_t* d_keys, uint64_t n, void* stream = nullptr) { check(osb200_sort_keys_u64(h_, d_keys, n, stream), "osb200_sort_keys_u64"); }
    void SortPairs(uint32_t* d_keys, uint32_t* d_values, uint64_t n, void* stream = nullptr)
    {
        check(osb200_sort_pairs_u32(h_, d_keys, d_values, n, stream), "osb200_sort_pairs_u32");
    }
    // stable sort on the key bits [begin_bit, end_bit) only; d_values may be null
    void SortBits(void* d_keys, uint32_t* d_values, uint64_t n, int begin_bit, int end_bit, void* stream = nullptr)
    {
        check(osb200_sort_bits(h_, d_keys, d_values, n, begin_bit, end_bit, stream), "osb200_sort_bits");
    }
    // signed / float keys and descending order (osb200_key_type)
    void SortKeysTyped(void* d_keys, uint64_t n, int key_type, bool descending, void* stream = nullptr)
    {
        check(osb200_sort_keys_typed(h_, d_keys, n, key_type, descending ? 1 : 0, stream), "osb200_sort_keys_typed");
    }
    // every segment [offsets[i], offsets[i+1]) sorted ascending and stable in place, one thread block per segment
    // (reference: SplitSort, SegSort/SplitSort/SplitSort.cuh:702-938); d_values may be null; max_segment_len <= 16,384
    void SegmentedSort(uint32_t* d_keys, uint32_t* d_values, const uint64_t* d_segment_offsets, uint64_t num_segments,
                       uint32_t max_segment_len, void* stream = nullptr)
    {
        check(osb200_segmented_sort_u32(h_, d_keys, d_values, d_segment_offsets, num_segments, max_segment_len, stream),
              "osb200_segmented_sort_u32");
    }
    void SetOption(const char* key, int64_t value) { check(osb200_set_option(h_, key, value), "osb200_set_option"); }
    uint64_t Validate(const void* d_keys, uint64_t n, void* stream = nullptr)
    {
        uint64_t e = 0;
        check(osb200_validate(h_, d_keys, n, &e, stream), "osb200_validate");
        return e;
    }
    uint64_t max_n() const { return max_n_; }
    osb200_handle handle() const { return h_; }

    static void check(int status, const char* what)
    {
        if (status != OSB200_OK) throw std::runtime_error(std::string(what) + ": " + osb200_status_string(status));
    }

  private:
    osb200_handle h_ = nullptr;
    uint64_t max_n_;
};

namespace OneSweep {
namespace detail {
// one cached sorter per (key width, pairs); grown on demand.  Not thread-safe across concurrent sorts of the same kind
// (neither is the reference's dispatcher: shared tickets/descriptors).
inline OneSweepSorterB200& sorter(int key_bytes, int value_bytes, uint64_t n)
{
    static std::mutex mu;
    static std::map<std::tuple<int, int>, std::unique_ptr<OneSweepSorterB200>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto& slot = cache[{key_bytes, value_bytes}];
    if (!slot || slot->max_n() < n) slot.reset(new OneSweepSorterB200(n ? n : 1, key_bytes, value_bytes));
    return *slot;
}
}  // namespace detail

inline void Sort(uint32_t* d_keys, uint64_t n, void* stream = nullptr) { detail::sorter(4, 0, n).SortKeys(d_keys, n, stream); }
inline void Sort(uint64_t* d_keys, uint64_t n, void* stream = nullptr) { detail::sorter(8, 0, n).SortKeys(d_keys, n, stream); }
inline void Sort(uint32_t* d_keys, uint32_t* d_values, uint64_t n, void* stream = nullptr)
{
    detail::sorter(4, 4, n).SortPairs(d_keys, d_values, n, stream);
}
}  // namespace OneSweep
