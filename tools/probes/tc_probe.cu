// Probe A: one CTA, one 128 x 256 x 256 s8 GEMM on tcgen05 with hand-built descriptors; checks against the CPU.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (int i = 0; i < 4000000; ++i) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128) probe(const uint8_t *xa, const uint8_t *xb, int *d_out, int *status) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sa = smem, *sb = smem + 32768;
    __shared__ __align__(8) uint64_t bar_full, bar_mma;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        mbar_expect_tx(&bar_full, 32768 + 65536);
        bulk_g2s(sa, xa, 16384, &bar_full);
        bulk_g2s(sa + 16384, xa + 16384, 16384, &bar_full);
        bulk_g2s(sb, xb, 32768, &bar_full);
        bulk_g2s(sb + 32768, xb + 32768, 32768, &bar_full);
        if (!mbar_wait_bounded(&bar_full, 0)) { status[0] = 1; }
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (32u << 17) | (8u << 24);
        for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad = make_desc(smem_u32(sa) + (ks >> 2) * 16384 + (ks & 3) * 32);
            const uint64_t bd = make_desc(smem_u32(sb) + (ks >> 2) * 32768 + (ks & 3) * 32);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    __syncthreads();
    if (!mbar_wait_bounded(&bar_mma, 0)) { if (tid == 0) status[0] |= 2; }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int ch = 0; ch < 8; ++ch) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + ch * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                       "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                       "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) d_out[tid * 256 + ch * 32 + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// host: expanded planes.  plane ka (128 bits each), row r: 128 bytes, 16-byte chunk c stored at chunk c ^ (r % 8)
static void expand(const std::vector<uint8_t> &desc, int rows, std::vector<uint8_t> &x) {
    x.assign((size_t)2 * rows * 128, 0);
    for (int r = 0; r < rows; ++r)
        for (int bit = 0; bit < 256; ++bit) {
            const int b = (desc[r * 32 + bit / 8] >> (bit % 8)) & 1;
            const int ka = bit / 128, c = (bit % 128) / 16, e = bit % 16;
            x[(size_t)ka * rows * 128 + (size_t)r * 128 + ((c ^ (r % 8)) * 16) + e] = b ? 0xFF : 0x01;
        }
}

int main() {
    const int M = 128, N = 256;
    std::vector<uint8_t> da(M * 32), db(N * 32), xa, xb;
    srand(1);
    for (auto &v : da) v = rand();
    for (auto &v : db) v = rand();
    expand(da, M, xa);
    expand(db, N, xb);
    uint8_t *gxa, *gxb; int *gout, *gstat;
    cudaMalloc(&gxa, xa.size()); cudaMalloc(&gxb, xb.size()); cudaMalloc(&gout, M * N * 4); cudaMalloc(&gstat, 4);
    cudaMemcpy(gxa, xa.data(), xa.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(gxb, xb.data(), xb.size(), cudaMemcpyHostToDevice);
    cudaMemset(gstat, 0, 4); cudaMemset(gout, 0x7f, M * N * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024);
    probe<<<1, 128, 98304 + 1024>>>(gxa, gxb, gout, gstat);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> out(M * N); int stat = -1;
    cudaMemcpy(out.data(), gout, M * N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&stat, gstat, 4, cudaMemcpyDeviceToHost);
    printf("kernel: %s, status %d\n", cudaGetErrorString(e), stat);
    int bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int h = 0;
            for (int b = 0; b < 32; ++b) h += __builtin_popcount(da[i * 32 + b] ^ db[j * 32 + b]);
            const int want = 256 - 2 * h;
            if (out[i * N + j] != want) { if (bad < 10) printf("  D[%d][%d] = %d, want %d\n", i, j, out[i * N + j], want); ++bad; }
        }
    printf("mismatches: %d of %d\n", bad, M * N);
    return 0;
}
