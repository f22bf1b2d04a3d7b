// osb_sharded.cu -- multi-GPU sharded sort: one MSD bucket-exchange pass over NVLink, then a local OneSweep.
//
// No reference equivalent (the reference is single-device); this is BASELINE.json's fifth config.
// One process per GPU.  Every rank holds n_local unsorted keys; after the call rank r holds the r-th contiguous
// slice of the global ascending order.
//
//   1. 256-bin histogram of the most significant digit of the local keys            (digit_histogram_kernel)
//   2. all-gather of the R histograms                                                (ncclAllGather, 2 KB per rank)
//   3. plan: contiguous bucket ranges -> ranks, balanced on the global counts        (osb200_sharded_plan, host)
//   4. exchange pass, two implementations:
//        fused  (default): the ordinary DigitBinningPass kernel scatters straight into the peers' receive buffers
//                through CUDA-IPC-mapped NVLink addresses -- its per-digit output bases are "virtual element indices"
//                that encode peer addresses, so ranking, chained scan and the NVLink stores are ONE kernel and the
//                keys cross HBM once (read) + NVLink once (write);
//        staged: DigitBinningPass into a local send buffer, then ncclSend/ncclRecv per peer (baseline).
//   5. local OneSweep (osb200_sort_keys_u32) on the received keys.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <fcntl.h>
#include <nccl.h>
#include <sys/mman.h>
#include <unistd.h>

#include "../../include/onesweep_b200.h"
#include "osb_internal.h"

namespace {

constexpr int kRadix = 256;
constexpr int kMaxWorld = 64;

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? OSB200_OK : OSB200_ERR_CUDA - static_cast<int>(e); }
#define OSB_TRY(expr)                                    \
    do {                                                 \
        cudaError_t e__ = (expr);                        \
        if (e__ != cudaSuccess) return cuda_status(e__); \
    } while (0)
// EXPERIMENT, off unless OSB_SHARDED_RENDEZVOUS is set: a host rendezvous of the ranks (POSIX shared memory, busy-polled) at
// the entry of every sharded sort.  Round 2 measured, on this round's 4- and 8-GPU boxes, steps of 29 / 77 ms instead of round
// 1's 17 / 18 ms with every kernel at its expected duration (MSD histogram 0.73, exchange 4.6, local sort 11.05 ms) and NCCL's
// tiny collectives at 18-30 us in isolation: the time is skew between the ranks at the three collectives of a step
// (profiles/r02_sharded_host_wait.txt).  Not the cause: the NVML sampler, host enqueue time (50 + 120 us per call), the way
// the host waits (blocking or busy poll), NCCL's algorithm (NVLS / Ring / LL).  Starting the steps together with this
// rendezvous made it worse (46 ms at N = 4), so the skew does not come from the hosts entering the call at different times.
// Unresolved at the end of the round; N = 2 is unaffected (15.8 ms per step).
inline void host_rendezvous(std::atomic<unsigned long long>* arrivals, int world)
{
    if (!arrivals || world < 2) return;
    const unsigned long long ticket = arrivals->fetch_add(1, std::memory_order_acq_rel);
    const unsigned long long target = (ticket / static_cast<unsigned long long>(world) + 1) * static_cast<unsigned long long>(world);
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (arrivals->load(std::memory_order_acquire) < target) {
        for (int i = 0; i < 32; ++i) __builtin_ia32_pause();
        if ((++spins & 0xfffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(1)) return;
    }
}

// busy poll (see osb200_sharded_sort_keys_u32): returns as soon as the event has completed
inline cudaError_t spin_until(cudaEvent_t e)
{
    cudaError_t r;
    while ((r = cudaEventQuery(e)) == cudaErrorNotReady)
        for (int i = 0; i < 32; ++i) __builtin_ia32_pause();
    return r;
}
#define OSB_NCCL(expr)                                   \
    do {                                                 \
        ncclResult_t r__ = (expr);                       \
        if (r__ != ncclSuccess) return OSB200_ERR_NCCL;  \
    } while (0)

}  // namespace


// ---- EXPERIMENT (OSB_SHARDED_MAPPED=1, never measured): no copy-engine operations inside a step ---------------------------
// The all-gathered histograms reach the host, and the plan reaches the device, through MAPPED pinned memory written / read by
// tiny kernels on the same stream; the host polls a flag word in that memory.  Written while chasing the N = 4 / 8 skew of
// round 2 (profiles/r02_sharded_host_wait.txt) under the hypothesis that a copy ordered behind a long-waiting compute kernel
// starts late; the skew turned out to come from the benchmark pinning every rank's main thread to core 0 (OMP_PROC_BIND, see
// bench.py), and the round's GPU budget was spent before this path could run once.  Kept for the next round to measure.
__global__ void __launch_bounds__(256)
publish_to_host_kernel(const unsigned long long* __restrict__ src, volatile unsigned long long* dst_mapped, int count,
                       volatile unsigned long long* flag_mapped, unsigned long long step)
{
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst_mapped[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { *flag_mapped = step; __threadfence_system(); }
}
__global__ void __launch_bounds__(256)
stage_from_host_kernel(const volatile unsigned long long* src_mapped, unsigned long long* __restrict__ dst, int count)
{
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src_mapped[i];
}

struct osb200_sharded_sorter {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    uint64_t max_n_local = 0, capacity = 0;  // capacity = receive-side keys (max_n_local + slack)
    osb200_handle exch = nullptr;            // kernels/state for the exchange pass over the local input
    osb200_handle local = nullptr;           // local OneSweep over the received keys
    uint32_t* send_buf = nullptr;            // staged mode only
    uint32_t* recv_buf = nullptr;
    unsigned long long* d_hist = nullptr;      // [256] local MSD histogram
    unsigned long long* d_hist_all = nullptr;  // [world][256]
    unsigned long long* d_out_base = nullptr;  // [256] virtual element indices (fused mode)
    unsigned long long* h_hist_all = nullptr;  // pinned
    unsigned long long* h_out_base = nullptr;  // pinned
    unsigned long long* h_coarse_hist = nullptr;  // pinned
    // device aliases of the three mapped pinned buffers above and of the flag word the host polls
    unsigned long long* dm_hist_all = nullptr;
    unsigned long long* dm_out_base = nullptr;
    unsigned long long* dm_coarse_hist = nullptr;
    unsigned long long* h_flag = nullptr;
    unsigned long long* dm_flag = nullptr;
    unsigned long long step = 0;
    bool copy_engine = true;   // cudaMemcpyAsync for the 8 KB readback and the 4 KB plan; OSB_SHARDED_MAPPED=1: the kernels below
    uint32_t* d_flag = nullptr;                // 1-element all-reduce used as a stream-ordered cross-GPU barrier
    void* peer_recv[kMaxWorld] = {};           // IPC-mapped receive buffers of all ranks (own = recv_buf)
    bool fused = true;
    bool force_fine = false;  // always use the 256-bucket plan (tests)
    int last_bins = 0;
    cudaEvent_t ev[4] = {};
    cudaEvent_t tev[3] = {};  // (OSB_SHARDED_TRACE) after the MSD histogram kernel, before / after the exchange kernel
    bool tev_valid = false;
    cudaEvent_t sync_ev = nullptr;  // host wait for the all-gathered histograms (busy-polled)
    // host rendezvous of the ranks (all on one node) in POSIX shared memory: keeps the ranks' steps in phase, see host_rendezvous
    std::atomic<unsigned long long>* shm_arrivals = nullptr;
    char shm_name[48] = {};
    float last_ms[4] = {0, 0, 0, 0};
};

extern "C" {

// Host-side plan shared by every rank (pure function of the all-gathered histograms; exported for CPU tests).
//   hist_all   [world][256]  digit counts per source rank
//   dest       [256]         owner rank of every bucket: contiguous, non-decreasing, balanced on global counts
//   recv_count [world]       keys each rank ends up with
//   recv_off   [world][256]  for THIS rank as a source (`rank`): element offset inside dest[d]'s receive buffer where
//                            its keys of bucket d go.  Layout at a destination: bucket-major, source-rank-minor,
//                            i.e. exactly the globally stable order of the MSD partition.
OSB200_API int osb200_sharded_plan(const uint64_t* hist_all, int world, int rank, int32_t* dest, uint64_t* recv_count,
                                   uint64_t* recv_off)
{
    if (!hist_all || !dest || !recv_count || !recv_off || world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return OSB200_ERR_INVALID_ARG;
    uint64_t bucket[kRadix], total = 0;
    for (int d = 0; d < kRadix; ++d) {
        bucket[d] = 0;
        for (int r = 0; r < world; ++r) bucket[d] += hist_all[r * kRadix + d];
        total += bucket[d];
    }
    // greedy contiguous split: bucket d goes to the rank whose ideal range contains the bucket's midpoint
    uint64_t before = 0;
    int prev = 0;
    for (int d = 0; d < kRadix; ++d) {
        int q = prev;
        if (total) {
            const long double mid = static_cast<long double>(before) + static_cast<long double>(bucket[d]) / 2;
            q = static_cast<int>(mid * world / static_cast<long double>(total));
            if (q >= world) q = world - 1;
            if (q < prev) q = prev;
        }
        dest[d] = q;
        prev = q;
        before += bucket[d];
    }
    for (int r = 0; r < world; ++r) recv_count[r] = 0;
    uint64_t fill[kMaxWorld] = {};  // running fill of every destination, bucket-major
    for (int d = 0; d < kRadix; ++d) {
        const int q = dest[d];
        uint64_t off = fill[q];
        for (int r = 0; r < world; ++r) {
            if (r == rank) recv_off[d] = off;
            off += hist_all[r * kRadix + d];
        }
        fill[q] = off;
    }
    for (int r = 0; r < world; ++r) recv_count[r] = fill[r];
    return OSB200_OK;
}

int osb200_sharded_unique_id(void* out_128_bytes)
{
    if (!out_128_bytes) return OSB200_ERR_INVALID_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
    ncclUniqueId id;
    OSB_NCCL(ncclGetUniqueId(&id));
    std::memcpy(out_128_bytes, &id, sizeof(id));
    return OSB200_OK;
}

int osb200_sharded_destroy(osb200_sharded_handle h)
{
    if (!h) return OSB200_ERR_INVALID_ARG;
    for (int r = 0; r < h->world; ++r)
        if (r != h->rank && h->peer_recv[r]) cudaIpcCloseMemHandle(h->peer_recv[r]);
    if (h->exch) osb200_destroy(h->exch);
    if (h->local) osb200_destroy(h->local);
    cudaFree(h->send_buf);
    cudaFree(h->recv_buf);
    cudaFree(h->d_hist);
    cudaFree(h->d_hist_all);
    cudaFree(h->d_out_base);
    cudaFree(h->d_flag);
    cudaFreeHost(h->h_hist_all);
    cudaFreeHost(h->h_out_base);
    cudaFreeHost(h->h_coarse_hist);
    cudaFreeHost(h->h_flag);
    for (cudaEvent_t e : h->ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->tev) if (e) cudaEventDestroy(e);
    if (h->sync_ev) cudaEventDestroy(h->sync_ev);
    if (h->shm_arrivals) munmap(static_cast<void*>(h->shm_arrivals), 4096);
    if (h->shm_name[0] && h->rank == 0) shm_unlink(h->shm_name);
    if (h->comm) ncclCommDestroy(h->comm);
    delete h;
    return OSB200_OK;
}

int osb200_sharded_create(osb200_sharded_handle* out, const void* unique_id_128_bytes, int rank, int world,
                          uint64_t max_n_local, int slack_percent)
{
    if (!out) return OSB200_ERR_INVALID_ARG;
    *out = nullptr;
    if (!unique_id_128_bytes || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || max_n_local == 0 ||
        slack_percent < 0 || slack_percent > 400)
        return OSB200_ERR_INVALID_ARG;
    osb200_sharded_sorter* s = new (std::nothrow) osb200_sharded_sorter();
    if (!s) return OSB200_ERR_ALLOC;
    s->rank = rank;
    s->world = world;
    s->max_n_local = max_n_local;
    s->capacity = max_n_local + max_n_local / 100 * slack_percent + 4096;

    ncclUniqueId id;
    std::memcpy(&id, unique_id_128_bytes, sizeof(id));
    if (ncclCommInitRank(&s->comm, world, id, rank) != ncclSuccess) { osb200_sharded_destroy(s); return OSB200_ERR_NCCL; }

    if (world > 1 && std::getenv("OSB_SHARDED_RENDEZVOUS")) {  // opt-in experiment (measured: it did not help, see host_rendezvous)
        unsigned long long tag = 0;
        std::memcpy(&tag, unique_id_128_bytes, sizeof(tag));
        unsigned long long tag2 = 0;
        std::memcpy(&tag2, static_cast<const char*>(unique_id_128_bytes) + 8, sizeof(tag2));
        std::snprintf(s->shm_name, sizeof(s->shm_name), "/osb200_%016llx%016llx", tag, tag2);
        const int fd = shm_open(s->shm_name, O_CREAT | O_RDWR, 0600);
        if (fd >= 0) {
            if (ftruncate(fd, 4096) == 0) {
                void* p = mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                if (p != MAP_FAILED) s->shm_arrivals = static_cast<std::atomic<unsigned long long>*>(p);  // zero-filled on creation
            }
            close(fd);
        }
    }
    int st = osb200_create(&s->exch, max_n_local, 4, 0);
    if (st == OSB200_OK) st = osb200_create(&s->local, s->capacity, 4, 0);
    if (st != OSB200_OK) { osb200_sharded_destroy(s); return st; }
    bool ok = cudaMalloc(&s->recv_buf, s->capacity * sizeof(uint32_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_hist, kRadix * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_hist_all, static_cast<size_t>(world) * kRadix * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_out_base, kRadix * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_flag, 64) == cudaSuccess;
    auto mapped = [&](unsigned long long** host, unsigned long long** dev, size_t count) {
        void* hp = nullptr;
        void* dp = nullptr;
        if (cudaHostAlloc(&hp, count * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess) return false;
        std::memset(hp, 0, count * sizeof(unsigned long long));
        *host = static_cast<unsigned long long*>(hp);
        if (cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess) return false;
        *dev = static_cast<unsigned long long*>(dp);
        return true;
    };
    ok = ok && mapped(&s->h_hist_all, &s->dm_hist_all, static_cast<size_t>(world) * kRadix);
    ok = ok && mapped(&s->h_out_base, &s->dm_out_base, kRadix);
    ok = ok && mapped(&s->h_coarse_hist, &s->dm_coarse_hist, kRadix);
    ok = ok && mapped(&s->h_flag, &s->dm_flag, 8);
    s->copy_engine = std::getenv("OSB_SHARDED_MAPPED") == nullptr;  // default: cudaMemcpyAsync (the measured path)
    for (auto& e : s->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s->sync_ev, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); osb200_sharded_destroy(s); return OSB200_ERR_ALLOC; }
    cudaMemset(s->d_flag, 0, 64);

    // Map every peer's receive buffer (CUDA IPC over NVLink/NVSwitch; all ranks are processes on one node).
    s->peer_recv[rank] = s->recv_buf;
    s->fused = world > 1;
    if (world > 1) {
        cudaIpcMemHandle_t mine;
        std::vector<cudaIpcMemHandle_t> all(world);
        cudaIpcMemHandle_t* d_handles = nullptr;
        bool ipc_ok = cudaIpcGetMemHandle(&mine, s->recv_buf) == cudaSuccess;
        ipc_ok = ipc_ok && cudaMalloc(&d_handles, sizeof(cudaIpcMemHandle_t) * world) == cudaSuccess;
        if (ipc_ok) {
            cudaMemcpy(d_handles + rank, &mine, sizeof(mine), cudaMemcpyHostToDevice);
            ncclResult_t r = ncclAllGather(d_handles + rank, d_handles, sizeof(mine), ncclChar, s->comm, nullptr);
            ipc_ok = r == ncclSuccess && cudaStreamSynchronize(nullptr) == cudaSuccess;
            if (ipc_ok) cudaMemcpy(all.data(), d_handles, sizeof(mine) * world, cudaMemcpyDeviceToHost);
        }
        // every rank must take the same decision: agree on success with an all-reduce(min)
        int* d_ok = reinterpret_cast<int*>(s->d_flag) + 8;
        int flag = ipc_ok ? 1 : 0;
        if (ipc_ok) {
            for (int r = 0; r < world && flag; ++r) {
                if (r == rank) continue;
                if (cudaIpcOpenMemHandle(&s->peer_recv[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    s->peer_recv[r] = nullptr;
                    flag = 0;
                }
            }
        }
        cudaMemcpy(d_ok, &flag, sizeof(int), cudaMemcpyHostToDevice);
        if (ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, s->comm, nullptr) != ncclSuccess ||
            cudaStreamSynchronize(nullptr) != cudaSuccess) {
            cudaFree(d_handles);
            osb200_sharded_destroy(s);
            return OSB200_ERR_NCCL;
        }
        cudaMemcpy(&flag, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
        s->fused = flag == 1;
        cudaFree(d_handles);
        cudaGetLastError();
    }
    if (!s->fused && world > 1) {
        if (cudaMalloc(&s->send_buf, max_n_local * sizeof(uint32_t)) != cudaSuccess) { osb200_sharded_destroy(s); return OSB200_ERR_ALLOC; }
    }
    *out = s;
    return OSB200_OK;
}

// 0 = staged (NCCL send/recv), 1 = fused NVLink scatter.  Must be called identically on every rank.
OSB200_API int osb200_sharded_set_fused(osb200_sharded_handle h, int fused)
{
    if (!h) return OSB200_ERR_INVALID_ARG;
    if (fused) {
        for (int r = 0; r < h->world; ++r) if (!h->peer_recv[r]) return OSB200_ERR_UNSUPPORTED;
        h->fused = true;
        return OSB200_OK;
    }
    if (!h->send_buf && cudaMalloc(&h->send_buf, h->max_n_local * sizeof(uint32_t)) != cudaSuccess) return OSB200_ERR_ALLOC;
    h->fused = false;
    return OSB200_OK;
}

// Testing hook: 1 = always use the 256-bucket greedy plan even when the 2^k equal-width split would fit.
OSB200_API int osb200_sharded_force_fine(osb200_sharded_handle h, int on)
{
    if (!h) return OSB200_ERR_INVALID_ARG;
    h->force_fine = on != 0;
    return OSB200_OK;
}

OSB200_API int osb200_sharded_local_handle(osb200_sharded_handle h, osb200_handle* exch, osb200_handle* local)
{
    if (!h) return OSB200_ERR_INVALID_ARG;
    if (exch) *exch = h->exch;
    if (local) *local = h->local;
    return OSB200_OK;
}

int osb200_sharded_sort_keys_u32(osb200_sharded_handle h, const uint32_t* d_keys_local, uint64_t n_local,
                                 uint32_t** d_out, uint64_t* n_out, void* stream)
{
    if (!h || !d_out || !n_out || (n_local && !d_keys_local)) return OSB200_ERR_INVALID_ARG;
    if (n_local > h->max_n_local) return OSB200_ERR_SIZE;
    cudaStream_t q = static_cast<cudaStream_t>(stream);
    const int R = h->world;

    // OSB_SHARDED_TRACE=1: host-side wall clock of this call's segments on stderr (diagnosis of straggling ranks)
    static const bool trace = std::getenv("OSB_SHARDED_TRACE") != nullptr;
    using clk = std::chrono::steady_clock;
    if (h->shm_arrivals) host_rendezvous(h->shm_arrivals, R);
    const clk::time_point t0 = clk::now();
    if (trace) {
        for (auto& e : h->tev) if (!e) OSB_TRY(cudaEventCreate(&e));
        if (h->tev_valid) {  // kernel-only durations of the PREVIOUS call (its events have long completed)
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventSynchronize(h->ev[3]);
            cudaEventElapsedTime(&a, h->ev[0], h->tev[0]);
            cudaEventElapsedTime(&b, h->tev[0], h->ev[1]);
            cudaEventElapsedTime(&c, h->ev[1], h->tev[1]);
            cudaEventElapsedTime(&d, h->tev[1], h->tev[2]);
            float e2 = 0;
            cudaEventElapsedTime(&e2, h->tev[2], h->ev[2]);
            std::fprintf(stderr, "[osb sharded rank %d] prev call: MSD hist kernel %.3f ms, allgather+D2H+sync+plan %.3f ms, barrier1+H2D %.3f ms, exchange kernel %.3f ms, barrier2 %.3f ms\n",
                         h->rank, a, b, c, d, e2);
        }
    }
    OSB_TRY(cudaEventRecord(h->ev[0], q));
    // 1-2. most-significant-digit histogram, all-gather
    int st = osb_internal_digit_histogram(h->exch, d_keys_local, n_local, 24, h->d_hist, q);
    if (st != OSB200_OK) return st;
    if (trace) OSB_TRY(cudaEventRecord(h->tev[0], q));
    OSB_NCCL(ncclAllGather(h->d_hist, h->d_hist_all, kRadix, ncclUint64, h->comm, q));
    const unsigned long long step = ++h->step;
    if (h->copy_engine)
        OSB_TRY(cudaMemcpyAsync(h->h_hist_all, h->d_hist_all, static_cast<size_t>(R) * kRadix * sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, q));
    else
    {
        publish_to_host_kernel<<<1, 256, 0, q>>>(h->d_hist_all, h->dm_hist_all, R * kRadix, h->dm_flag, step);
        OSB_TRY(cudaGetLastError());
    }
    const clk::time_point t1 = clk::now();
    // The plan (and the receive size) is needed on the host: the one host wait of a step.  It polls the flag word the
    // publishing kernel writes into mapped pinned memory (see publish_to_host_kernel for why no copy engine is involved).
    OSB_TRY(cudaEventRecord(h->sync_ev, q));
    if (h->copy_engine) {
        OSB_TRY(spin_until(h->sync_ev));
    } else {
        // the flag is written by the kernel after a system-scope fence; the event (recorded behind the kernel) is the safety net
        const volatile unsigned long long* flag = h->h_flag;
        unsigned polls = 0;
        while (*flag != step) {
            for (int i = 0; i < 16; ++i) __builtin_ia32_pause();
            if ((++polls & 0x3ffu) == 0) {
                const cudaError_t r = cudaEventQuery(h->sync_ev);
                if (r == cudaSuccess) break;  // the kernel has completed: its writes are visible
                if (r != cudaErrorNotReady) return cuda_status(r);
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    const clk::time_point t2 = clk::now();

    // 3. plan.  Preferred: R = 2^k ranks and the equal-width split of the key space fits the receive buffers ->
    //    exchange on the top k bits only (R bins instead of 256): runs of ~n/(tiles*R) keys (8 KB at R = 8) keep the
    //    NVLink stores full-sector (measured: 32-key misaligned runs reach ~55 % of the 128-B-aligned peer store rate,
    //    profiles/r01_p2p_store_ub.txt) and the ranking atomics nearly conflict-free.  Otherwise: 256 buckets, greedy.
    int32_t dest[kRadix];
    uint64_t recv_count[kMaxWorld], recv_off[kRadix];
    uint32_t xshift = 24;  // digit position of the exchange pass
    int bins = kRadix;
    int k = 0;
    while ((1 << k) < R) ++k;
    bool coarse = R > 1 && (1 << k) == R && !h->force_fine;
    if (coarse) {
        const int per = kRadix >> k;
        for (int q2 = 0; q2 < R; ++q2) recv_count[q2] = 0;
        for (int src = 0; src < R; ++src)
            for (int d = 0; d < kRadix; ++d) recv_count[d / per] += h->h_hist_all[static_cast<size_t>(src) * kRadix + d];
        for (int q2 = 0; q2 < R; ++q2) coarse = coarse && recv_count[q2] <= h->capacity;
    }
    if (coarse) {
        const int per = kRadix >> k;
        xshift = 32 - k;
        bins = R;
        for (int b = 0; b < kRadix; ++b) { dest[b] = b < R ? b : R - 1; recv_off[b] = 0; }
        for (int b = 0; b < R; ++b) {
            uint64_t off = 0;
            for (int src = 0; src < h->rank; ++src)
                for (int d = b * per; d < (b + 1) * per; ++d) off += h->h_hist_all[static_cast<size_t>(src) * kRadix + d];
            recv_off[b] = off;
        }
        // the coarse histogram of THIS rank's keys, in the layout the pass expects (bins beyond R are empty)
        for (int b = 0; b < kRadix; ++b) h->h_out_base[b] = 0;
        const unsigned long long* my_hist = h->h_hist_all + static_cast<size_t>(h->rank) * kRadix;
        for (int d = 0; d < kRadix; ++d) h->h_out_base[d / per] += my_hist[d];
        // (staged through its own pinned buffer: no host sync needed before h_out_base is filled again below)
        std::memcpy(h->h_coarse_hist, h->h_out_base, kRadix * sizeof(unsigned long long));
        if (h->copy_engine)
            OSB_TRY(cudaMemcpyAsync(h->d_hist, h->h_coarse_hist, kRadix * sizeof(unsigned long long), cudaMemcpyHostToDevice, q));
        else
        {
            stage_from_host_kernel<<<1, 256, 0, q>>>(h->dm_coarse_hist, h->d_hist, kRadix);
            OSB_TRY(cudaGetLastError());
        }
    } else {
        st = osb200_sharded_plan(reinterpret_cast<const uint64_t*>(h->h_hist_all), R, h->rank, dest, recv_count, recv_off);
        if (st != OSB200_OK) return st;
    }
    h->last_bins = bins;
    const uint64_t mine = recv_count[h->rank];
    // Bucket imbalance beyond the slack chosen at create.  Every rank holds the whole recv_count[] and the same capacity
    // (max_n_local and slack_percent must be identical on all ranks), so ALL ranks take this exit together, before any
    // collective or peer store of the exchange: nobody is left waiting in a barrier for a rank that bailed out.
    for (int r = 0; r < R; ++r)
        if (recv_count[r] > h->capacity) return OSB200_ERR_SIZE;
    OSB_TRY(cudaEventRecord(h->ev[1], q));

    // 4. exchange
    if (R == 1) {
        OSB_TRY(cudaMemcpyAsync(h->recv_buf, d_keys_local, n_local * sizeof(uint32_t), cudaMemcpyDeviceToDevice, q));
    } else if (h->fused) {
        // all ranks have finished reading their receive buffers from the previous call before anyone writes
        OSB_NCCL(ncclAllReduce(h->d_flag, h->d_flag, 1, ncclUint32, ncclSum, h->comm, q));
        for (int d = 0; d < kRadix; ++d) {
            const unsigned long long peer = reinterpret_cast<unsigned long long>(h->peer_recv[dest[d]]);
            h->h_out_base[d] = peer / sizeof(uint32_t) + recv_off[d];  // virtual element index relative to address 0
        }
        if (h->copy_engine)
            OSB_TRY(cudaMemcpyAsync(h->d_out_base, h->h_out_base, kRadix * sizeof(unsigned long long), cudaMemcpyHostToDevice, q));
        else
        {
            stage_from_host_kernel<<<1, 256, 0, q>>>(h->dm_out_base, h->d_out_base, kRadix);
            OSB_TRY(cudaGetLastError());
        }
        if (trace) OSB_TRY(cudaEventRecord(h->tev[1], q));
        st = osb_internal_binning_pass(h->exch, d_keys_local, nullptr, n_local, xshift, h->d_hist, h->d_out_base, q);
        if (st != OSB200_OK) return st;
        if (trace) { OSB_TRY(cudaEventRecord(h->tev[2], q)); h->tev_valid = true; }
        // every rank's scatter kernel has completed (and its NVLink stores are performed) before any local sort starts
        OSB_NCCL(ncclAllReduce(h->d_flag, h->d_flag, 1, ncclUint32, ncclSum, h->comm, q));
    } else {
        st = osb_internal_binning_pass(h->exch, d_keys_local, h->send_buf, n_local, xshift, h->d_hist, nullptr, q);
        if (st != OSB200_OK) return st;
        // send_buf is bin-major; the bins of destination p are contiguous.  Receive source-major.
        uint64_t send_off[kMaxWorld + 1] = {}, send_cnt[kMaxWorld] = {}, roff[kMaxWorld + 1] = {};
        const int per = coarse ? (kRadix >> k) : 1;
        const unsigned long long* my_hist = h->h_hist_all + static_cast<size_t>(h->rank) * kRadix;
        for (int d = 0; d < kRadix; ++d) send_cnt[dest[coarse ? d / per : d]] += my_hist[d];
        for (int p = 0; p < R; ++p) send_off[p + 1] = send_off[p] + send_cnt[p];
        for (int src = 0; src < R; ++src) {
            uint64_t c = 0;
            for (int d = 0; d < kRadix; ++d)
                if (dest[coarse ? d / per : d] == h->rank) c += h->h_hist_all[static_cast<size_t>(src) * kRadix + d];
            roff[src + 1] = roff[src] + c;
        }
        OSB_NCCL(ncclGroupStart());
        for (int p = 0; p < R; ++p) {
            if (send_cnt[p]) OSB_NCCL(ncclSend(h->send_buf + send_off[p], send_cnt[p], ncclUint32, p, h->comm, q));
            if (roff[p + 1] - roff[p]) OSB_NCCL(ncclRecv(h->recv_buf + roff[p], roff[p + 1] - roff[p], ncclUint32, p, h->comm, q));
        }
        OSB_NCCL(ncclGroupEnd());
    }
    OSB_TRY(cudaEventRecord(h->ev[2], q));

    // 5. local OneSweep over the whole key.  (After a coarse exchange the keys of a rank share their top log2(R) bits, and a
    //    sort on the bits below them alone -- osb200_sort_bits(0, 32 - log2 R) -- is correct too, but measured slower at 2^30
    //    keys: 12.6 ms against 11.1 ms, its masked histogram costs 1.35 instead of 0.73 ms and its 5-bit last digit takes
    //    the few-bins scatter, 3.44 instead of 2.59 ms; profiles/r02_sharded_local_sort_bits.txt.)
    st = osb200_sort_keys_u32(h->local, h->recv_buf, mine, q);
    if (st != OSB200_OK) return st;
    OSB_TRY(cudaEventRecord(h->ev[3], q));
    if (trace) {
        const clk::time_point t3 = clk::now();
        auto us = [](clk::time_point a, clk::time_point b) { return static_cast<long>(std::chrono::duration_cast<std::chrono::microseconds>(b - a).count()); };
        std::fprintf(stderr, "[osb sharded rank %d] enqueue hist+allgather %ld us, wait %ld us, plan+enqueue exchange+local sort %ld us\n",
                     h->rank, us(t0, t1), us(t1, t2), us(t2, t3));
    }
    *d_out = h->recv_buf;
    *n_out = mine;
    return OSB200_OK;
}

int osb200_sharded_last_timing(osb200_sharded_handle h, float* out_ms4)
{
    if (!h || !out_ms4) return OSB200_ERR_INVALID_ARG;
    OSB_TRY(spin_until(h->ev[3]));
    OSB_TRY(cudaEventElapsedTime(&out_ms4[0], h->ev[0], h->ev[1]));
    OSB_TRY(cudaEventElapsedTime(&out_ms4[1], h->ev[1], h->ev[2]));
    OSB_TRY(cudaEventElapsedTime(&out_ms4[2], h->ev[2], h->ev[3]));
    OSB_TRY(cudaEventElapsedTime(&out_ms4[3], h->ev[0], h->ev[3]));
    return OSB200_OK;
}

}  // extern "C"
