// bfm_tensor.cuh - the matching kernel on the 5th-generation tensor cores (tcgen05, s8 operands, TMEM accumulators).
// Self-contained (own types, own PTX helpers): included by bfm_tensor.cu (the library) and by tools/tc_bench.cu (a
// stand-alone harness that checks it against a brute-force POPC kernel and times it).
//
// Hamming distance as a dot product: descriptor bit b becomes the s8 value 1 - 2b, so for two descriptors
// sum_k a_k * b_k = 256 - 2 * hamming(a, b), exactly, in the s32 accumulator.  One tcgen05.mma (M = 128 queries, N = 128
// train rows, K = 32 bits) replaces 16384 x 32 bit compares; eight of them give a 128 x 128 tile of complete distances.
// The POPC kernel (bfm_kernels.cuh) spends 24 instructions per pair and thread and is bound by the POPC pipe at 4 pairs
// per clock and SM; here the tensor pipe could deliver 30 and what binds is the epilogue that turns every distance
// into a key and keeps the two smallest per row (1 IMAD + 3 VIMNMX per pair: ~21 pairs per clock and SM).
//
// Layout.  expand_kernel writes every 32-byte descriptor as 256 bytes in two planes of 128 (bits 0..127 and 128..255);
// a plane is [row][128 bytes] with the eight 16-byte chunks of row r stored at chunk position c ^ (r % 8) - exactly the
// K-major SWIZZLE_128B shared-memory layout the MMA descriptors name, so a tile of rows is ONE contiguous range per
// plane and lands with a 1-D bulk copy (no tensor map).  Tiles start at multiples of 8 rows (1024-byte atoms), which
// the planner guarantees by giving every problem an 8-aligned base in the planes.
//
// One CTA per SM, persistent over the work items (a block of 256 query rows x a train range of the same problem):
//   warp 0     producer: the item's two A tiles (64 KB, double-buffered) and B tiles of 128 train rows (32 KB, 2 stages)
//   warp 1     MMA issuer: 2 x 8 tcgen05.mma per B tile into one of two 256-column TMEM buffers, tcgen05.commit
//   warps 2-9  epilogue: thread <-> query row (TMEM lane); tcgen05.ld 32 columns at a time, key = (-dot) * 2^21 + column
//              (one IMAD: the column is an immediate because the running keys are re-based by -128 per tile), best and
//              second in four interleaved min/max chains; at the end of the item the row-state atomics of the POPC kernel.
// Every mbarrier wait is bounded: a protocol error ends the kernel with *status = 1 instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bfm_tc {

struct Item {            // same fields as bfm::Segment
    int32_t q_row0;      // first row of the 256-row query block in the expanded query planes
    int32_t q_valid;     // rows of the block that exist
    int32_t q_local0;
    int32_t out_row0;
    int32_t t_row0;      // first row of the train range in the expanded train planes (multiple of 8)
    int32_t t_count;
    int32_t t_local0;
    int32_t problem;
};

struct Params {
    const uint8_t *xq, *xt;              // expanded planes: [2][rows][128] bytes
    unsigned long long xq_plane, xt_plane;   // bytes per plane
    const Item *items;
    int32_t n_items;
    unsigned long long *rowstate;        // [out rows] (best << 32) | second, all-ones when idle
    uint32_t *status;                    // set non-zero when a wait timed out
    int32_t mul;                         // -(1 << 21), passed as data so that the key stays ONE IMAD (FMA pipe)
};

constexpr int BQ = 256;          // query rows per item (two A tiles of 128)
constexpr int BT = 128;          // train rows per B tile
constexpr int NSTAGE = 2;        // B stages
constexpr int NTHREADS = 320;    // warp 0 producer, warp 1 MMA, warps 2..9 epilogue
constexpr uint32_t A_BYTES = BQ * 256, B_BYTES = BT * 256;
constexpr uint32_t SMEM_BYTES = 2 * A_BYTES + NSTAGE * B_BYTES + 1024;
constexpr int DIST_SHIFT_TC = 22;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
// bounded: a protocol error must end the kernel, not hang the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, volatile int *abort_flag) {
    for (int i = 0; i < 8000000; ++i) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
        if ((i & 1023) == 1023 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_s8(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}

// ---- expansion: one s8 per descriptor bit (+1 / -1), two planes of 128 bits, SWIZZLE_128B chunk order ---------
// The problem table is read as int32 words: q_begin, q_count, t_begin, t_count at words 0..3 of every entry; the
// problem's first row in the expanded planes at words off_xq0 / off_xt0 (both multiples of 8).
struct XProblem { int32_t q_begin, q_count, t_begin, t_count, xq0, xt0; };
__global__ void __launch_bounds__(256) expand_kernel(const uint16_t *__restrict__ q, const uint16_t *__restrict__ t, const int32_t *__restrict__ probs,
                                                     int stride, int off_xq0, int off_xt0, uint4 *__restrict__ xq, uint4 *__restrict__ xt,
                                                     unsigned long long xq_plane16, unsigned long long xt_plane16) {
    const int32_t *pr = probs + (size_t)blockIdx.y * stride;
    const bool train = blockIdx.z != 0;
    const int rows = train ? pr[3] : pr[1];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = idx >> 4, c16 = idx & 15;
    if (row >= rows) return;
    const uint16_t *src = train ? t : q;
    const size_t srow = (size_t)(train ? pr[2] : pr[0]) + row;
    const uint32_t bits = src[srow * 16 + c16];
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = (bits >> (4 * i)) & 0xFu;
        const uint32_t spread = (x * 0x00204081u) & 0x01010101u;   // bit e of x -> bit 0 of byte e
        w[i] = 0x01010101u | (spread * 0xFEu);                      // s8: bit 0 -> +1, bit 1 -> -1
    }
    const size_t xrow = (size_t)(train ? pr[off_xt0] : pr[off_xq0]) + row;
    const int ka = c16 >> 3, c = c16 & 7;
    uint4 *dst = (train ? xt : xq) + (size_t)ka * (train ? xt_plane16 : xq_plane16) + xrow * 8 + (c ^ (int)(xrow & 7));
    *dst = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- the scan ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) scan_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                       // [2 buffers][2 A tiles][2 planes][16 KB]
    uint8_t *sB = smem + 2 * A_BYTES;         // [NSTAGE][2 planes][16 KB]
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], b_full[NSTAGE], b_empty[NSTAGE], acc_full[2], acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        s_abort = 0;
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
    volatile int *abortp = &s_abort;

    if (warp == 0) {
        if (lane == 0) {
            // ---- producer ----
            uint32_t it = 0, T = 0;
            for (int item = blockIdx.x; item < p.n_items && !*abortp; item += gridDim.x, ++it) {
                const Item sg = p.items[item];
                const uint32_t ab = it & 1;
                if (!mbar_wait(&a_empty[ab], ((it >> 1) & 1) ^ 1, abortp)) break;
                mbar_expect_tx(&a_full[ab], A_BYTES);
#pragma unroll
                for (int at = 0; at < 2; ++at)
#pragma unroll
                    for (int ka = 0; ka < 2; ++ka)
                        bulk_g2s(sA + ab * A_BYTES + (at * 2 + ka) * 16384, p.xq + (size_t)ka * p.xq_plane + ((size_t)sg.q_row0 + at * 128) * 128, 16384, &a_full[ab]);
                const int ntiles = (sg.t_count + BT - 1) / BT;
                for (int tile = 0; tile < ntiles; ++tile, ++T) {
                    const uint32_t s = T % NSTAGE;
                    if (!mbar_wait(&b_empty[s], ((T / NSTAGE) & 1) ^ 1, abortp)) break;
                    mbar_expect_tx(&b_full[s], B_BYTES);
#pragma unroll
                    for (int ka = 0; ka < 2; ++ka)
                        bulk_g2s(sB + s * B_BYTES + ka * 16384, p.xt + (size_t)ka * p.xt_plane + ((size_t)sg.t_row0 + (size_t)tile * BT) * 128, 16384, &b_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer ----
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BT >> 3) << 17) | (8u << 24);   // s8 x s8 -> s32, K-major, N = 128, M = 128
            uint32_t it = 0, T = 0;
            for (int item = blockIdx.x; item < p.n_items && !*abortp; item += gridDim.x, ++it) {
                const int t_count = p.items[item].t_count;
                const uint32_t ab = it & 1;
                if (!mbar_wait(&a_full[ab], (it >> 1) & 1, abortp)) break;
                const int ntiles = (t_count + BT - 1) / BT;
                for (int tile = 0; tile < ntiles; ++tile, ++T) {
                    const uint32_t s = T % NSTAGE, cb = T & 1;
                    if (!mbar_wait(&b_full[s], (T / NSTAGE) & 1, abortp)) break;
                    if (!mbar_wait(&acc_empty[cb], ((T >> 1) & 1) ^ 1, abortp)) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int at = 0; at < 2; ++at) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint64_t ad = umma_desc(smem_u32(sA) + ab * A_BYTES + (at * 2 + (ks >> 2)) * 16384 + (ks & 3) * 32);
                            const uint64_t bd = umma_desc(smem_u32(sB) + s * B_BYTES + (ks >> 2) * 16384 + (ks & 3) * 32);
                            umma_s8(tmem + cb * 256 + at * 128, ad, bd, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&b_empty[s]);
                    umma_commit(&acc_full[cb]);
                }
                umma_commit(&a_empty[ab]);
            }
        }
    } else {
        // ---- epilogue: thread <-> one query row of the block ----
        const int e = warp - 2, quarter = warp & 3, at = e >> 2;
        const int lr = at * 128 + quarter * 32 + lane;
        const int MUL = p.mul;
        uint32_t T = 0;
        for (int item = blockIdx.x; item < p.n_items && !*abortp; item += gridDim.x) {
            const Item sg = p.items[item];
            int best[4], second[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { best[i] = 0x7FFFFFFF; second[i] = 0x7FFFFFFF; }
            const int ntiles = (sg.t_count + BT - 1) / BT;
            bool ok = true;
            for (int tile = 0; tile < ntiles; ++tile, ++T) {
                const uint32_t cb = T & 1;
                if (!mbar_wait(&acc_full[cb], (T >> 1) & 1, abortp)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int ncols = min(BT, sg.t_count - tile * BT);
                // keys are relative to the END of this tile: (-dot) * 2^21 + (column - 128); every tile shifts the running
                // keys down by 128, so one IMAD with an immediate column makes a key
#pragma unroll
                for (int i = 0; i < 4; ++i) { best[i] -= BT; second[i] -= BT; }
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    if (ch * 32 >= ncols) break;
                    uint32_t v[32];
                    tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + cb * 256 + at * 128 + ch * 32, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const int cbase = ch * 32 - BT;
                    if (ch * 32 + 32 <= ncols) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int key = (int)v[j] * MUL + (cbase + j);
                            second[j & 3] = min(second[j & 3], max(best[j & 3], key));
                            best[j & 3] = min(best[j & 3], key);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int key = (ch * 32 + j < ncols) ? (int)v[j] * MUL + (cbase + j) : 0x7FFFFFFF;
                            second[j & 3] = min(second[j & 3], max(best[j & 3], key));
                            best[j & 3] = min(best[j & 3], key);
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[cb]);
            }
            if (!ok) break;
            // merge the four interleaved chains
            int b1 = best[0], b2 = second[0];
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                b2 = min(max(b1, best[i]), min(b2, second[i]));
                b1 = min(b1, best[i]);
            }
            if (lr < sg.q_valid && b1 < 0x7FFFFFFF - (1 << 24) && ntiles > 0) {
                auto decode = [&](int key) -> uint32_t {
                    // key = nd * 2^21 + c, c in (-2^21, 0]: nd = -dot, c = column - ntiles * 128
                    const int nd = (key + (1 << 21) - 1) >> 21;
                    const int c = key - nd * (1 << 21);
                    const uint32_t dist = (uint32_t)(256 + nd) >> 1;
                    return (dist << DIST_SHIFT_TC) | (uint32_t)(sg.t_local0 + ntiles * BT + c);
                };
                const uint32_t key1 = decode(b1);
                const uint32_t key2 = b2 < 0x7FFFFFFF - (1 << 24) ? decode(b2) : 0xFFFFFFFFu;
                uint32_t *half = reinterpret_cast<uint32_t *>(p.rowstate + (size_t)(sg.out_row0 + lr));
                const uint32_t displaced = atomicMin(half + 1, key1);
                const uint32_t cand = min(max(displaced, key1), key2);
                if (cand != 0xFFFFFFFFu) atomicMin(half, cand);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    if (tid == 0 && s_abort && p.status) *p.status = 1u;
}

}  // namespace bfm_tc
