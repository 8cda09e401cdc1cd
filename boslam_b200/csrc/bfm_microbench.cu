// bfm_microbench.cu - integer-pipe issue-rate probes for sm_100a (build step 0).
//
// The matcher's roofline is the POPC issue rate, which MEASURED_PEAKS.json does not carry.  Each
// probe runs 8 independent dependency chains per thread, 1024 threads per CTA, 2 CTAs per SM, and
// reports thread-level operations per clock per SM (from clock64 deltas inside the kernel) and
// per second (from CUDA events around it).
#include "../../include/bfm.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace {

constexpr int MB_THREADS = 1024;
constexpr int CHAINS = 8;
constexpr int UNROLL = 16;

// per-CTA record: [smid, start clock, end clock]; the host takes (max end - min start) per SM,
// because the two co-resident CTAs of an SM do not get equal issue priority.
__device__ __forceinline__ void record(long long *cycles, long long t0, long long t1) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cycles[3 * blockIdx.x + 0] = (long long)smid;
    cycles[3 * blockIdx.x + 1] = t0;
    cycles[3 * blockIdx.x + 2] = t1;
}

template <int TEST>
__device__ __forceinline__ void step(uint32_t (&x)[CHAINS], uint32_t a, uint32_t b) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        if (TEST == 0) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
        if (TEST == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]), "r"(b));
        if (TEST == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]));
        if (TEST == 3) {
            asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
        }
        if (TEST == 4) {
            asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x[c]) : "r"(b), "r"(a));
        }
        if (TEST == 5) asm volatile("redux.sync.min.u32 %0, %0, 0xffffffff;" : "+r"(x[c]));
        if (TEST == 6) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
        if (TEST == 7) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a ^ (uint32_t)c));
        if (TEST == 8) {
            asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
        }
    }
}

// ops issued per chain step, per thread
__host__ __device__ constexpr int ops_per_step(int test) {
    return (test == 3 || test == 8) ? 2 : (test == 4 ? 3 : 1);
}

template <int TEST>
__global__ void __launch_bounds__(MB_THREADS) probe_kernel(int iters, uint32_t a, uint32_t b, uint32_t *sink,
                                                           long long *cycles) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 2654435761u + c * 40503u + a;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) step<TEST>(x, a, b);
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c];
    if (acc == 0x12345678u) sink[0] = acc;  // keeps the chains alive
    __syncthreads();
    if (threadIdx.x == 0) record(cycles, t0, t1);
}

// The plain per-pair instruction mix of the matcher: 8 XOR + 8 POPC + adds + key + min, against
// register-resident operands (no memory), to see what the mix sustains in isolation.
__global__ void __launch_bounds__(MB_THREADS) probe_pair_kernel(int iters, uint32_t a, uint32_t b, uint32_t *sink,
                                                                long long *cycles) {
    uint32_t q[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = threadIdx.x * 2654435761u + i * 40503u + a; t[i] = b + i; }
    uint32_t best = 0xFFFFFFFFu;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            uint32_t d = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) d += __popc(q[w] ^ t[w]);
            best = min(best, (d << 22) + (uint32_t)(i * UNROLL + u));
#pragma unroll
            for (int w = 0; w < 8; ++w) t[w] += a;  // new "train" words without touching memory
        }
    }
    const long long t1 = clock64();
    if (best == 0x12345678u) sink[0] = best;
    __syncthreads();
    if (threadIdx.x == 0) record(cycles, t0, t1);
}

typedef void (*ProbeFn)(int, uint32_t, uint32_t, uint32_t *, long long *);

ProbeFn probe(int test) {
    switch (test) {
        case 0: return probe_kernel<0>;
        case 1: return probe_kernel<1>;
        case 2: return probe_kernel<2>;
        case 3: return probe_kernel<3>;
        case 4: return probe_kernel<4>;
        case 5: return probe_kernel<5>;
        case 6: return probe_kernel<6>;
        case 7: return probe_kernel<7>;
        case 8: return probe_kernel<8>;
        case 9: return probe_pair_kernel;
        default: return nullptr;
    }
}

}  // namespace

extern "C" int bfm_microbench(int device, int test, int iters, double *ops_per_clk_per_sm, double *ops_per_s,
                              double *sm_mhz) {
    ProbeFn fn = probe(test);
    if (!fn || iters <= 0) return BFM_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return BFM_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BFM_ERR_CUDA;
    const int ctas = prop.multiProcessorCount * 2;
    uint32_t *sink = nullptr;
    long long *cyc = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&sink, 64) != cudaSuccess || cudaMalloc(&cyc, sizeof(long long) * 3 * ctas) != cudaSuccess) return BFM_ERR_NOMEM;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fn<<<ctas, MB_THREADS>>>(std::max(iters / 8, 1), 3u, 5u, sink, cyc);  // warm-up
    cudaEventRecord(e0);
    fn<<<ctas, MB_THREADS>>>(iters, 3u, 5u, sink, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    int rc = BFM_OK;
    if (e != cudaSuccess) {
        rc = BFM_ERR_CUDA;
    } else {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(3 * (size_t)ctas);
        cudaMemcpy(h.data(), cyc, sizeof(long long) * 3 * ctas, cudaMemcpyDeviceToHost);
        // per SM: busy span = max(end) - min(start) over its CTAs, work = CTAs it ran
        std::vector<long long> lo(1024, -1), hi(1024, -1);
        std::vector<int> cnt(1024, 0);
        for (int i = 0; i < ctas; ++i) {
            const int sm = (int)(h[3 * i] & 1023);
            lo[sm] = lo[sm] < 0 ? h[3 * i + 1] : std::min(lo[sm], h[3 * i + 1]);
            hi[sm] = std::max(hi[sm], h[3 * i + 2]);
            ++cnt[sm];
        }
        double rate_sum = 0;  // mean over SMs of CTAs-per-cycle
        int sms = 0;
        for (int sm = 0; sm < 1024; ++sm)
            if (cnt[sm]) { rate_sum += (double)cnt[sm] / (double)(hi[sm] - lo[sm]); ++sms; }
        const double mean_cyc = 2.0 * sms / rate_sum;  // cycles an SM needs for 2 CTAs' work
        // per-thread ops: test 9 counts the 8 POPCs of a pair as the unit (POPC/clk/SM of the mix)
        const double per_thread = test == 9 ? (double)iters * UNROLL * 8.0
                                            : (double)iters * UNROLL * CHAINS * ops_per_step(test);
        const double total = per_thread * MB_THREADS * (double)ctas;
        // 2 CTAs share an SM: an SM executes 2 CTAs' work in ~mean_cyc cycles
        if (ops_per_clk_per_sm) *ops_per_clk_per_sm = per_thread * MB_THREADS * 2.0 / mean_cyc;
        if (ops_per_s) *ops_per_s = total / (ms * 1e-3);
        if (sm_mhz) *sm_mhz = (total / (ms * 1e-3)) / (per_thread * MB_THREADS * 2.0 / mean_cyc * prop.multiProcessorCount) / 1e6;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    return rc;
}
