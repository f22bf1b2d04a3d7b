// osb_internal.h -- internal (non-exported) interfaces shared by osb_host.cu and osb_sharded.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/onesweep_b200.h"

// One stable digit-binning pass of the handle's kernels with caller-provided digit counts:
//   d_hist256   device [256] u64 counts of the digit at `shift` over d_in (the pass scans them itself)
//   out_base    if non-null: device [256] u64 "virtual element index" of the first key of every digit relative to
//               d_out (used by the fused NVLink exchange, where d_out is nullptr and the indices encode peer
//               addresses); if null the exclusive scan of d_hist256 is used (ordinary pass into d_out).
int osb_internal_binning_pass(osb200_handle h, const void* d_in, void* d_out, uint64_t n, uint32_t shift,
                              const unsigned long long* d_hist256, const unsigned long long* out_base,
                              cudaStream_t stream);
// 256-bin histogram of the digit at `shift` (d_hist256 is overwritten).
int osb_internal_digit_histogram(osb200_handle h, const void* d_in, uint64_t n, uint32_t shift,
                                 unsigned long long* d_hist256, cudaStream_t stream);
