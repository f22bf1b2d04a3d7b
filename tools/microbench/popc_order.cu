// Does a shared-memory atomicAdd(&bin[d], 1) (SASS: ATOMS.POPC.INC) hand out its return values to the lanes of
// one warp instruction in ascending lane order?  If so it is a single-instruction, *stable* warp-level
// multisplit.  Compares every returned value with the ballot-derived stable rank.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int WARPS = 16, THREADS = WARPS * 32, K = 16;

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t match_ballot8(uint32_t d) {
    uint32_t mask = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) { const bool p = (d >> b) & 1; const uint32_t bal = __ballot_sync(0xffffffffu, p); mask &= p ? bal : ~bal; }
    return mask;
}

// mode: 0 uniform digits; 1..4 AND of (mode+1) draws (low entropy); 5: digit = lane/2 pattern; 6: all equal; 7: only 2 distinct
template <int partial, int VARIANT>
__global__ void __launch_bounds__(THREADS) check(unsigned long long* errs, unsigned long long* total, int iters, uint32_t seed, int mode)
{
    __shared__ uint32_t hist[WARPS * 256];
    __shared__ uint32_t ref[WARPS * 256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* wh = hist + warp * 256; uint32_t* wr = ref + warp * 256;
    uint32_t s = seed ^ ((blockIdx.x * THREADS + threadIdx.x) * 2654435761u);
    const uint32_t lt = lanemask_lt();
    unsigned long long bad = 0, cnt = 0;
    for (int it = 0; it < iters; ++it) {
        for (int i = lane; i < 256; i += 32) { wh[i] = 0; wr[i] = 0; }
        __syncwarp();
        for (int i = 0; i < K; ++i) {
            uint32_t d = 0xff;
            const int draws = (mode >= 1 && mode <= 4) ? mode + 1 : 1;
            for (int k = 0; k < draws; ++k) { s = s * 1664525u + 1013904223u; d &= (s >> 13); }
            d &= 255u;
            if (mode == 5) d = ((lane >> 1) + it + i) & 255u;
            if (mode == 6) d = (it * 7 + i) & 255u;
            if (mode == 7) d = ((s >> 9) & 1) ? 17u : 200u;
            // optionally only a data-dependent subset of lanes participates (exited/inactive lanes)
            const bool active = partial ? (((s >> 20) & 3) != 0) : true;
            const uint32_t amask = __ballot_sync(0xffffffffu, active);
            if (active) {
                const uint32_t got = VARIANT == 0 ? atomicAdd(&wh[d], 1u) : atomicAdd(&wh[d], (uint32_t)(threadIdx.x >= 0));
                // stable reference rank among the active lanes
                uint32_t m = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; ++b) { const bool p = (d >> b) & 1; const uint32_t bal = __ballot_sync(amask, p); m &= p ? bal : ~bal; }
                m &= amask;
                const uint32_t below = __popc(m & lt);
                uint32_t pre = 0;
                if (below == 0) { pre = wr[d]; wr[d] = pre + __popc(m); }
                pre = __shfl_sync(amask, pre, __ffs(m) - 1);
                bad += (got != pre + below);
                cnt += 1;
            }
            __syncwarp();
        }
        __syncwarp();
    }
    atomicAdd(errs, bad); atomicAdd(total, cnt);
}

int main()
{
    unsigned long long *errs, *total; CK(cudaMalloc(&errs, 8)); CK(cudaMalloc(&total, 8));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long gbad = 0;
    for (int cfg = 0; cfg < 4; ++cfg)
    for (int mode = 0; mode < 8; ++mode) {
        const int partial = cfg & 1;
        CK(cudaMemset(errs, 0, 8)); CK(cudaMemset(total, 0, 8));
        if (cfg == 0) check<0, 0><<<sms * 4, THREADS>>>(errs, total, 400, 12345u + mode, mode);
        if (cfg == 1) check<1, 0><<<sms * 4, THREADS>>>(errs, total, 400, 12345u + mode, mode);
        if (cfg == 2) check<0, 1><<<sms * 4, THREADS>>>(errs, total, 400, 12345u + mode, mode);
        if (cfg == 3) check<1, 1><<<sms * 4, THREADS>>>(errs, total, 400, 12345u + mode, mode);
        CK(cudaDeviceSynchronize());
        unsigned long long e, t; CK(cudaMemcpy(&e, errs, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&t, total, 8, cudaMemcpyDeviceToHost));
        printf("cfg %d mode %d partial %d: %llu mismatches / %llu atomics\n", cfg, mode, partial, e, t);
        gbad += e;
    }
    printf("TOTAL mismatches %llu => ATOMS.POPC.INC %s lane-ordered\n", gbad, gbad ? "IS NOT" : "is");
    return 0;
}
