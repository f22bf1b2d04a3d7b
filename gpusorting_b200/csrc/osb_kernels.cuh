// osb_kernels.cuh -- launch interface of the sm_100a OneSweep kernels (implemented in osb_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "osb_common.cuh"

namespace osb {

// Kernel variants of the digit-binning pass (osb200_set_option "variant").
enum BinningVariant : int {
    kVariantTilePerCta = 0,  // one CTA per partition tile, keys loaded straight into registers
    kVariantPersistent = 1,  // persistent CTAs, TMA (cp.async.bulk) double-buffered tile staging
    kVariantWide = 2,        // 16,384-key tiles, two-phase atomic ranking, compact reductions + one-shot lookback
};
constexpr int kNumVariants = 3;
enum RankMode : int {
    kRankAtomic = 0,  // one shared-memory atomicAdd per key (lane-ordered on sm_100, verified at create)
    kRankBallot = 1,  // 8 ballots per key (the reference's warp-level multisplit, OneSweep.cu:208-253)
};

struct BinningConfig {
    int variant = kVariantTilePerCta;
    int rank_mode = kRankAtomic;
    int sm_count = 148;
    KeyCodec codec;  // typed keys: encode-on-load / decode-on-store for THIS launch (flags 0 = plain unsigned keys)
    // --- per-launch pass parameters (kVariantWide honours all of them; the other variants need the defaults) ---
    uint32_t digit_bits = 8;          // width of this pass's digit, 1..8 (narrower: last place of a begin/end-bit sort)
    const SortPlan* plan = nullptr;   // device plan: skip flag, ping-pong parity, dynamic codec flags; null = as launched
    uint32_t place = 0;               // index of this pass in the plan
    uint32_t spin_cap = 2048;         // lookback polls of one predecessor before the digit thread re-reduces that tile itself
    uint32_t debug_stall_every = 0;   // test hook: tiles with tile % N == N-1 never publish their reduction (0 = off)
    bool hot_passes = false;          // also enqueue the HOT instantiation (the plan decides which of the two runs the pass)
};

// keys per partition tile for a key width / pairs flag / variant (host needs it to size descriptors)
uint32_t binning_tile_keys(int key_bytes, bool pairs, const BinningConfig& cfg);

// One-time per-process kernel attribute setup (dynamic shared memory opt-in). Returns cudaError_t.
cudaError_t configure_kernels();

// GlobalHistogram (reference: OneSweep::GlobalHistogram, Sort/OneSweep.cu:44-123).
// ghist[place*256 + digit] += counts; caller zeroes ghist first.
cudaError_t launch_global_histogram(const void* keys, uint64_t n, int key_bytes, unsigned long long* ghist,
                                    int sm_count, cudaStream_t stream, const KeyCodec* codec = nullptr);

// Single-place histogram (used by the sharded path for the most significant digit): hist256[digit] += counts.
cudaError_t launch_digit_histogram(const void* keys, uint64_t n, int key_bytes, uint32_t shift,
                                   unsigned long long* hist256, int sm_count, cudaStream_t stream);

// Scan (reference: OneSweep::Scan, Sort/OneSweep.cu:125-162): per place exclusive prefix of ghist -> gbase.
// With plan != null the kernel also writes the device launch plan: a place is skipped when one of its bins holds all n
// keys (allow_skip), and the first/last executed places are recorded for the typed-key codec.
cudaError_t launch_scan(const unsigned long long* ghist, unsigned long long* gbase, int places, cudaStream_t stream,
                        SortPlan* plan = nullptr, uint64_t n = 0, bool allow_skip = false, bool allow_hot = false);

// GlobalHistogram of a begin_bit/end_bit sort: place p counts the digit (key >> (begin_bit + 8p)) & mask_p, mask_p = 255
// except for the last place, which keeps last_bits bits.  (The byte-aligned full-width case uses launch_global_histogram.)
cudaError_t launch_global_histogram_bits(const void* keys, uint64_t n, int key_bytes, unsigned long long* ghist, int sm_count,
                                         cudaStream_t stream, const KeyCodec* codec, uint32_t begin_bit, int places,
                                         uint32_t last_bits);

// If the plan says an odd number of passes ran, the sorted data sits in the alt buffers: move it to the caller's.
cudaError_t launch_copy_back(const SortPlan* plan, const void* alt_keys, void* keys, const uint32_t* alt_vals, uint32_t* vals,
                             uint64_t n, int key_bytes, int sm_count, cudaStream_t stream);

// DigitBinningPass (reference: OneSweep::DigitBinningPassKeysOnly / Pairs, Sort/OneSweep.cu:164-600).
//   gbase_place: [256] exclusive global digit bases for this digit place
//   desc:        [tiles][256] 64-bit tile descriptors (never cleared; `epoch` distinguishes launches)
//   agg16:       [tiles][256] 16-bit tile reductions (flag:1|count:15) used by kVariantWide; zeroed per sort
//   ticket:      one zeroed u32 (dynamic tile id counter, reference `index[]`)
cudaError_t launch_digit_binning(const void* in, void* out, const uint32_t* in_val, uint32_t* out_val, uint64_t n,
                                 int key_bytes, uint32_t shift, const unsigned long long* gbase_place, uint64_t* desc,
                                 uint16_t* agg16, uint32_t* ticket, uint32_t epoch, const BinningConfig& cfg,
                                 cudaStream_t stream);

// Segment sort / small-n path: one CTA sorts one segment (<= segment_sort_capacity keys) in shared memory, all digit passes
// in one launch.  seg_off == nullptr: the single segment [0, single_n).  max_len (an upper bound of the segment lengths)
// picks the geometry; segments longer than the capacity are skipped by the kernel (the caller must not pass them).
uint32_t segment_sort_capacity(int key_bytes, bool small);
cudaError_t launch_segment_sort(void* keys, uint32_t* vals, int key_bytes, const unsigned long long* seg_off, uint64_t num_segments,
                                uint64_t single_n, uint32_t max_len, uint32_t begin_bit, uint32_t places, uint32_t last_bits,
                                const KeyCodec* codec, int rank_mode, int sm_count, cudaStream_t stream);

// Validate (reference: Validate, UtilityKernels.cuh:403-429): err_count += #(keys[i] > keys[i+1]).
cudaError_t launch_validate(const void* keys, uint64_t n, int key_bytes, unsigned long long* err_count, int sm_count,
                            cudaStream_t stream);

// InitRandom (reference: InitRandom, UtilityKernels.cuh:53-117, launched <<<256,256>>> by
// OneSweepDispatcher.cuh:100-104): the reference's deterministic test-input generator.  65,536 hybrid
// Tausworthe/LCG streams; stream g writes elements g, g+65536, ...; each element ANDs and_count+1 draws.
// payload (may be null) receives a copy of the key (reference pairs overload) or, if payload_is_index, i.
cudaError_t launch_init_random(uint32_t* keys, uint32_t* payload, uint64_t n, uint32_t and_count, uint32_t seed,
                               bool payload_is_index, cudaStream_t stream);

// Device self-test: does a shared-memory atomicAdd hand out its return values in ascending lane order among
// the lanes of one warp instruction that hit the same address?  (kRankAtomic depends on it.)
// mismatches (device u64) receives the number of violations found.
cudaError_t launch_atomic_order_selftest(unsigned long long* mismatches, int sm_count, cudaStream_t stream);

}  // namespace osb
