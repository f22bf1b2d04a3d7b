/*
 * oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's OneSweep radix-sort path (b0nes164/GPUSorting,
 * GPUSortingCUDA/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (gpusorting_b200/) never does.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path, so the
 * oracle is pinned against outputs of the reference itself: oracle/ref_harness.cu compiles the
 * reference's own OneSweep.cu / UtilityKernels.cuh from /root/reference into oracle/_ref/, runs them on
 * the B200 box, and tests/golden/ref_*.json hold the digests that run produced
 * (generator: tests/golden/make_ref_golden.py).  tests/test_oracle.py checks every function below
 * against those fixtures.
 *
 * Every function cites the reference file:line whose behaviour it restates
 * (paths relative to /root/reference/GPUSortingCUDA/).
 */
#define _GNU_SOURCE
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_RADIX 256u

/* ---------------------------------------------------------------------------------------------
 * Input generator.  Restates InitRandom, UtilityKernels.cuh:53-83 (keys) and :85-117 (pairs), as
 * launched by OneSweepDispatcher.cuh:100-104,215-219 with <<<256,256>>>: 65,536 generator threads,
 * thread g seeds z_j = (4g+j)*seed (32-bit wraparound), takes one warm-up step of the hybrid
 * Tausworthe/LCG generator (UtilityKernels.cuh:26-33), then writes elements g, g+65536, ...; each
 * element is the AND of (and_count+1) successive draws (Thearling & Smith entropy reduction,
 * UtilityKernels.cuh:42-52,70-81).
 * ------------------------------------------------------------------------------------------- */
#define ORC_GENERATORS 65536u

typedef struct { uint32_t z1, z2, z3, z4; } orc_taus;

static inline uint32_t orc_taus_next(orc_taus* s)
{
    s->z1 = ((s->z1 & 4294967294u) << 12) ^ (((s->z1 << 13) ^ s->z1) >> 19);
    s->z2 = ((s->z2 & 4294967288u) << 4) ^ (((s->z2 << 2) ^ s->z2) >> 25);
    s->z3 = ((s->z3 & 4294967280u) << 17) ^ (((s->z3 << 3) ^ s->z3) >> 11);
    s->z4 = s->z4 * 1664525u + 1013904223u;
    return s->z1 ^ s->z2 ^ s->z3 ^ s->z4;
}

static inline void orc_taus_seed(orc_taus* s, uint32_t g, uint32_t seed)
{
    s->z1 = (g << 2) * seed;
    s->z2 = ((g << 2) + 1u) * seed;
    s->z3 = ((g << 2) + 2u) * seed;
    s->z4 = ((g << 2) + 3u) * seed;
    (void)orc_taus_next(s); /* the warm-up step at UtilityKernels.cuh:65-68 */
}

/* keys[i] for i in [0,n): UtilityKernels.cuh:53-83.  n is 64-bit here so that the same stream can be
 * extended past the reference's uint32 size for the u64 / sharded configs (our definition). */
void orc_init_random_u32(uint32_t* keys, uint64_t n, uint32_t and_count, uint32_t seed)
{
    const uint64_t gens = n < ORC_GENERATORS ? n : ORC_GENERATORS;
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < (int64_t)gens; ++g) {
        orc_taus s;
        orc_taus_seed(&s, (uint32_t)g, seed);
        for (uint64_t i = (uint64_t)g; i < n; i += ORC_GENERATORS) {
            uint32_t t = 0xffffffffu;
            for (uint32_t k = 0; k <= and_count; ++k) t &= orc_taus_next(&s);
            keys[i] = t;
        }
    }
}

/* Pairs overload, UtilityKernels.cuh:85-117: the reference sets payload = key (:115). */
void orc_init_random_pairs_u32(uint32_t* keys, uint32_t* payload, uint64_t n, uint32_t and_count, uint32_t seed)
{
    orc_init_random_u32(keys, n, and_count, seed);
    memcpy(payload, keys, (size_t)n * sizeof(uint32_t));
}

/* u64 keys: hi = draw 2i, lo = draw 2i+1 of the same u32 stream of length 2n (our definition; the CUDA
 * reference has no 64-bit path). */
void orc_init_random_u64(uint64_t* keys, uint64_t n, uint32_t and_count, uint32_t seed)
{
    uint32_t* tmp = (uint32_t*)malloc((size_t)(2 * n) * sizeof(uint32_t));
    if (!tmp) return;
    orc_init_random_u32(tmp, 2 * n, and_count, seed);
    for (uint64_t i = 0; i < n; ++i) keys[i] = ((uint64_t)tmp[2 * i] << 32) | tmp[2 * i + 1];
    free(tmp);
}

/* ---------------------------------------------------------------------------------------------
 * GlobalHistogram, Sort/OneSweep.cu:44-123: counts of every 8-bit digit at each digit place in one
 * read of the keys; output layout [place][digit] (SEC/THIRD/FOURTH_RADIX_START = 256/512/768,
 * OneSweep.cu:21-23,116-122).  key_bytes = 4 (reference) or 8 (ours), places = key_bytes.
 * ------------------------------------------------------------------------------------------- */
void orc_global_histogram(const void* keys, uint64_t n, int key_bytes, uint64_t* hist /* [key_bytes*256] */)
{
    memset(hist, 0, (size_t)key_bytes * ORC_RADIX * sizeof(uint64_t));
    const uint8_t* p = (const uint8_t*)keys;
    for (uint64_t i = 0; i < n; ++i)
        for (int b = 0; b < key_bytes; ++b) /* little-endian: byte b is digit place b */
            hist[(size_t)b * ORC_RADIX + p[i * (uint64_t)key_bytes + b]]++;
}

/* Scan, Sort/OneSweep.cu:125-162: per digit place, exclusive prefix sum over the 256 bins; this is the
 * value the reference stores (tagged FLAG_INCLUSIVE) in descriptor slot 0 of each pass. */
void orc_scan_exclusive(const uint64_t* hist, int places, uint64_t* excl)
{
    for (int pl = 0; pl < places; ++pl) {
        uint64_t run = 0;
        for (uint32_t d = 0; d < ORC_RADIX; ++d) {
            excl[(size_t)pl * ORC_RADIX + d] = run;
            run += hist[(size_t)pl * ORC_RADIX + d];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * One DigitBinningPass, Sort/OneSweep.cu:164-344 (keys) / :346-600 (pairs): a STABLE counting sort on
 * the 8-bit digit at `shift`.  The reference reaches this through warp-level multisplit + chained scan;
 * its net effect on memory is exactly: dst[ excl[digit(k)] + (number of earlier keys with the same
 * digit) ] = k, in input order (in-order ranking OneSweep.cu:207-253, padding keeps order :195-205).
 * ------------------------------------------------------------------------------------------- */
void orc_binning_pass_u32(const uint32_t* src, uint32_t* dst, const uint32_t* src_val, uint32_t* dst_val,
                          uint64_t n, uint32_t shift)
{
    uint64_t cnt[ORC_RADIX];
    memset(cnt, 0, sizeof(cnt));
    for (uint64_t i = 0; i < n; ++i) cnt[(src[i] >> shift) & 255u]++;
    uint64_t run = 0;
    for (uint32_t d = 0; d < ORC_RADIX; ++d) { uint64_t c = cnt[d]; cnt[d] = run; run += c; }
    if (src_val) {
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t pos = cnt[(src[i] >> shift) & 255u]++;
            dst[pos] = src[i];
            dst_val[pos] = src_val[i];
        }
    } else {
        for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> shift) & 255u]++] = src[i];
    }
}

void orc_binning_pass_u64(const uint64_t* src, uint64_t* dst, uint64_t n, uint32_t shift)
{
    uint64_t cnt[ORC_RADIX];
    memset(cnt, 0, sizeof(cnt));
    for (uint64_t i = 0; i < n; ++i) cnt[(src[i] >> shift) & 255u]++;
    uint64_t run = 0;
    for (uint32_t d = 0; d < ORC_RADIX; ++d) { uint64_t c = cnt[d]; cnt[d] = run; run += c; }
    for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> shift) & 255u]++] = src[i];
}

/* Full sort = the launch plan of OneSweepDispatcher.cuh:311-336 (keys) / :338-363 (pairs): passes at
 * radixShift 0,8,16,24 ping-ponging sort->alt->sort->alt->sort, result back in `keys`.
 * `alt` (and `alt_val`) are caller-provided scratch of n elements.  Returns 0. */
int orc_onesweep_keys_u32(uint32_t* keys, uint32_t* alt, uint64_t n)
{
    orc_binning_pass_u32(keys, alt, NULL, NULL, n, 0);
    orc_binning_pass_u32(alt, keys, NULL, NULL, n, 8);
    orc_binning_pass_u32(keys, alt, NULL, NULL, n, 16);
    orc_binning_pass_u32(alt, keys, NULL, NULL, n, 24);
    return 0;
}

int orc_onesweep_pairs_u32(uint32_t* keys, uint32_t* vals, uint32_t* alt, uint32_t* alt_val, uint64_t n)
{
    orc_binning_pass_u32(keys, alt, vals, alt_val, n, 0);
    orc_binning_pass_u32(alt, keys, alt_val, vals, n, 8);
    orc_binning_pass_u32(keys, alt, vals, alt_val, n, 16);
    orc_binning_pass_u32(alt, keys, alt_val, vals, n, 24);
    return 0;
}

/* 64-bit keys: the same plan extended to 8 digit places (no CUDA reference). */
int orc_onesweep_keys_u64(uint64_t* keys, uint64_t* alt, uint64_t n)
{
    for (uint32_t pl = 0; pl < 8; pl += 2) {
        orc_binning_pass_u64(keys, alt, n, pl * 8);
        orc_binning_pass_u64(alt, keys, n, pl * 8 + 8);
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Validate, UtilityKernels.cuh:403-429 (keys): number of adjacent inversions keys[i] > keys[i+1];
 * :432-479 (pairs): additionally counts inversions in the payload array itself (valid for the
 * reference's payload==key inputs only).
 * ------------------------------------------------------------------------------------------- */
uint64_t orc_validate_keys_u32(const uint32_t* keys, uint64_t n)
{
    uint64_t err = 0;
    for (uint64_t i = 1; i < n; ++i) err += keys[i - 1] > keys[i];
    return err;
}

uint64_t orc_validate_keys_u64(const uint64_t* keys, uint64_t n)
{
    uint64_t err = 0;
    for (uint64_t i = 1; i < n; ++i) err += keys[i - 1] > keys[i];
    return err;
}

uint64_t orc_validate_pairs_u32(const uint32_t* keys, const uint32_t* vals, uint64_t n)
{
    uint64_t err = 0;
    for (uint64_t i = 1; i < n; ++i) err += (keys[i - 1] > keys[i]) + (vals[i - 1] > vals[i]);
    return err;
}

/* ---------------------------------------------------------------------------------------------
 * Host-parallel port of the same algorithm (used only as the timed CPU baseline, bench.py
 * --impl reference / cpu_baseline): per-thread histograms of contiguous chunks give each thread the
 * chunk-exclusive digit offsets -- the CPU analogue of the per-tile reductions chained in
 * OneSweep.cu:257-327 -- so the scatter stays stable.  4 (or 8) passes, ping-pong as above.
 * ------------------------------------------------------------------------------------------- */
#define ORC_WC 16 /* keys buffered per bin before a burst store (64 B for u32) */

static void orc_par_pass(const void* src, void* dst, const uint32_t* sv, uint32_t* dv, uint64_t n,
                         uint32_t shift, int key_bytes, int threads, uint64_t* table /* threads*256 */)
{
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num();
#else
        const int t = 0;
#endif
        const uint64_t lo = n * (uint64_t)t / (uint64_t)threads, hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
        uint64_t* c = table + (size_t)t * ORC_RADIX;
        memset(c, 0, ORC_RADIX * sizeof(uint64_t));
        if (key_bytes == 4) { const uint32_t* s = (const uint32_t*)src; for (uint64_t i = lo; i < hi; ++i) c[(s[i] >> shift) & 255u]++; }
        else                { const uint64_t* s = (const uint64_t*)src; for (uint64_t i = lo; i < hi; ++i) c[(s[i] >> shift) & 255u]++; }
#pragma omp barrier
#pragma omp single
        {
            uint64_t run = 0;
            for (uint32_t d = 0; d < ORC_RADIX; ++d)
                for (int tt = 0; tt < threads; ++tt) { uint64_t v = table[(size_t)tt * ORC_RADIX + d]; table[(size_t)tt * ORC_RADIX + d] = run; run += v; }
        }
        if (key_bytes == 4 && !sv) {
            /* software write-combining: random 4-byte stores become 64-byte bursts (order within a bin is kept) */
            const uint32_t* s = (const uint32_t*)src; uint32_t* o = (uint32_t*)dst;
            uint32_t buf[ORC_RADIX][ORC_WC];
            uint8_t fill[ORC_RADIX];
            memset(fill, 0, sizeof(fill));
            for (uint64_t i = lo; i < hi; ++i) {
                const uint32_t k = s[i], d = (k >> shift) & 255u;
                buf[d][fill[d]++] = k;
                if (fill[d] == ORC_WC) { memcpy(o + c[d], buf[d], sizeof(buf[d])); c[d] += ORC_WC; fill[d] = 0; }
            }
            for (uint32_t d = 0; d < ORC_RADIX; ++d) { memcpy(o + c[d], buf[d], fill[d] * sizeof(uint32_t)); c[d] += fill[d]; }
        } else if (key_bytes == 4) {
            const uint32_t* s = (const uint32_t*)src; uint32_t* o = (uint32_t*)dst;
            for (uint64_t i = lo; i < hi; ++i) { uint64_t p = c[(s[i] >> shift) & 255u]++; o[p] = s[i]; dv[p] = sv[i]; }
        } else {
            const uint64_t* s = (const uint64_t*)src; uint64_t* o = (uint64_t*)dst;
            uint64_t buf[ORC_RADIX][ORC_WC / 2];
            uint8_t fill[ORC_RADIX];
            memset(fill, 0, sizeof(fill));
            for (uint64_t i = lo; i < hi; ++i) {
                const uint64_t k = s[i]; const uint32_t d = (uint32_t)(k >> shift) & 255u;
                buf[d][fill[d]++] = k;
                if (fill[d] == ORC_WC / 2) { memcpy(o + c[d], buf[d], sizeof(buf[d])); c[d] += ORC_WC / 2; fill[d] = 0; }
            }
            for (uint32_t d = 0; d < ORC_RADIX; ++d) { memcpy(o + c[d], buf[d], fill[d] * sizeof(uint64_t)); c[d] += fill[d]; }
        }
    }
}

/* threads the baseline may use: OpenMP's limit capped by the CPUs this process is allowed to run on */
int orc_host_threads(void)
{
#ifdef _OPENMP
    int t = omp_get_max_threads();
#else
    int t = 1;
#endif
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int a = CPU_COUNT(&set);
        if (a > 0 && a < t) t = a;
    }
    return t < 1 ? 1 : t;
}

/* key_bytes 4|8; vals/alt_val may be NULL (keys only). threads<=0 -> all. Result in keys(/vals). */
int orc_onesweep_parallel(void* keys, void* alt, uint32_t* vals, uint32_t* alt_val, uint64_t n, int key_bytes, int threads)
{
    if (threads <= 0) threads = orc_host_threads();
    uint64_t* table = (uint64_t*)malloc((size_t)threads * ORC_RADIX * sizeof(uint64_t));
    if (!table) return -1;
    for (int pl = 0; pl < key_bytes; pl += 2) {
        orc_par_pass(keys, alt, vals, alt_val, n, (uint32_t)pl * 8u, key_bytes, threads, table);
        orc_par_pass(alt, keys, alt_val, vals, n, (uint32_t)pl * 8u + 8u, key_bytes, threads, table);
    }
    free(table);
    return 0;
}

/* FNV-1a style 64-bit digest of a buffer (for golden fixtures: order-sensitive). */
uint64_t orc_digest(const void* buf, uint64_t bytes)
{
    const uint8_t* p = (const uint8_t*)buf;
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
