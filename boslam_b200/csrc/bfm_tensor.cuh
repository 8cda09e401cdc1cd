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
//   warp 0      producer: the item's two A tiles (64 KB, double-buffered) and B tiles of 128 train rows (32 KB, 3 stages)
//   warps 1, 10 MMA issuers, one per A tile: 8 tcgen05.mma per B tile into that tile's half of one of two 256-column TMEM
//               buffers, tcgen05.commit per half (the halves are handed to the epilogue separately)
//   warps 2-9, 11-18  epilogue, one group per TMEM buffer (every second B tile): four warps per scheduler in different
//               phases fill the issue slots that two leave idle (fixed-latency waits between dependent instructions);
//               thread <-> query row (TMEM lane); tcgen05.ld 32 columns at a time (the next load in flight while
//               these are folded, the first load of the next tile under the last fold), tile-local 16-bit keys two to a
//               register (fold_chunk), the tile's two smallest folded into 32-bit keys; at the end of the item the
//               row-state atomics of the POPC kernel.
// Every mbarrier wait is bounded: a protocol error ends the kernel with *status = 1 instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bfm_tc {

// timing probes of the stand-alone harness (tools/tc_bench.cu defines BFM_TC_HARNESS); constant false in the library
#ifdef BFM_TC_HARNESS
#define TC_DBG(p, bits) ((p).dbg & (bits))
#else
#define TC_DBG(p, bits) 0
#endif

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
    int32_t dbg;                         // harness only (TC_DBG), bits: 1 epilogue without folds (one load per tile), 2 no MMA, 4 no B loads, 8 no row-state commit; 0 in the library
    uint32_t mul_lo, mul_hi;             // -64 and -64 << 16 (two's complement), passed as data so that the keys stay IMADs (FMA pipe)
};

constexpr int BQ = 256;          // query rows per item (two A tiles of 128)
constexpr int BT = 128;          // train rows per B tile
constexpr int NTHREADS = 608;    // warp 0 producer, warps 1 and 10 MMA (one per A tile), warps 2..9 and 11..18 epilogue
constexpr uint32_t A_BYTES = BQ * 256, B_BYTES = BT * 256;
constexpr uint32_t smem_bytes(int nstage, int na = 2) { return (uint32_t)na * A_BYTES + (uint32_t)nstage * B_BYTES + 1024; }
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

// every tcgen05.ld of this thread has returned; the registers pass through the statement, so no use of them can be
// scheduled above it
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]),
                   "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                   "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// 32 accumulator columns (dot products of one query row with train rows CH * 32 .. + 31 of the tile) into the running
// tile-local keys.  key16 = distance << 7 | column (column < 128, distance = (256 - dot) / 2 <= 256) = 16384 - 64 * dot +
// column; columns c and c + 16 share a register, so ONE pair of IMADs (FMA pipe; the multipliers are kernel
// parameters so that they stay IMADs) makes two keys.  Two such registers are ordered (min, max) and merged into the
// running two smallest with a three-input minimum: 5 VIMNMX(3).U16x2 per four keys - 1.25 ALU instructions per
// distance instead of 3 with 32-bit keys; the ALU pipe is what the epilogue spends.
template <int CH, bool MASKED>
__device__ __forceinline__ void fold_chunk(const uint32_t (&v)[32], uint32_t (&p1)[2], uint32_t (&p2)[2], uint32_t m_lo, uint32_t m_hi, int ncols) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        uint32_t P[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t c_lo = CH * 32 + j + u, c_hi = c_lo + 16;
            const uint32_t cst = (16384u + c_lo) + ((16384u + c_hi) << 16);
            P[u] = v[j + u + 16] * m_hi + (v[j + u] * m_lo + cst);   // modulo 2^32; both halves are exact and non-negative
            if (MASKED) {
                if ((int)c_lo >= ncols) P[u] |= 0x0000FFFFu;
                if ((int)c_hi >= ncols) P[u] |= 0xFFFF0000u;
            }
        }
        const uint32_t lo = __vminu2(P[0], P[1]), hi = __vmaxu2(P[0], P[1]);
        const int c = (j >> 1) & 1;
        p2[c] = __vimin3_u16x2(p2[c], __vmaxu2(p1[c], lo), hi);
        p1[c] = __vminu2(p1[c], lo);
    }
}

// ---- expansion: one s8 per descriptor bit (+1 / -1), two planes of 128 bits, SWIZZLE_128B chunk order ---------
// The problem table is read as int32 words: q_begin, q_count, t_begin, t_count at words 0..3 of every entry; the
// problem's first row in the expanded planes at words off_xq0 / off_xt0 (both multiples of 8).
struct XProblem { int32_t q_begin, q_count, t_begin, t_count, xq0, xt0; };
constexpr int EXPAND_ROWS = 64;   // rows per CTA of expand_kernel: 16 threads per row, 4 rows per thread
__global__ void __launch_bounds__(256) expand_kernel(const uint16_t *__restrict__ q, const uint16_t *__restrict__ t, const int32_t *__restrict__ probs,
                                                     int stride, int off_xq0, int off_xt0, uint4 *__restrict__ xq, uint4 *__restrict__ xt,
                                                     unsigned long long xq_plane16, unsigned long long xt_plane16) {
    const int32_t *pr = probs + (size_t)blockIdx.y * stride;
    const bool train = blockIdx.z != 0;
    const int rows = train ? pr[3] : pr[1];
    const int row0 = blockIdx.x * EXPAND_ROWS + (threadIdx.x >> 4), c16 = threadIdx.x & 15;
    if (blockIdx.x * EXPAND_ROWS >= rows) return;
    const uint16_t *src = (train ? t : q) + (size_t)(train ? pr[2] : pr[0]) * 16 + c16;
    // the kernel is bound by the write stream (8 bytes out per byte in): four rows per thread keep four 16-byte stores
    // in flight, each instruction of a warp still writing whole 128-byte lines
    uint32_t bits[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int row = row0 + k * 16;
        bits[k] = row < rows ? src[(size_t)row * 16] : 0u;
    }
    const size_t xbase = (size_t)(train ? pr[off_xt0] : pr[off_xq0]);
    const int ka = c16 >> 3, c = c16 & 7;
    uint4 *plane = (train ? xt : xq) + (size_t)ka * (train ? xt_plane16 : xq_plane16);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int row = row0 + k * 16;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = (bits[k] >> (4 * i)) & 0xFu;
            const uint32_t spread = (x * 0x00204081u) & 0x01010101u;   // bit e of x -> bit 0 of byte e
            w[i] = 0x01010101u | (spread * 0xFEu);                      // s8: bit 0 -> +1, bit 1 -> -1
        }
        const size_t xrow = xbase + row;
        if (row < rows) plane[xrow * 8 + (c ^ (int)(xrow & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- the scan ------------------------------------------------------------------------------------------
// One tile of one epilogue thread: chunk 0 of the tile is already on its way into va.  The other three loads fly while
// the previous chunk is folded; the accumulator half goes back to the MMA warp as soon as the last load has returned,
// and - before the last fold - the first load of the NEXT tile is issued (its accumulator is normally complete by
// then: the wait costs nothing and the load's latency disappears under the fold).
// (the last tile of a train range: chunks that lie inside the range take the plain fold, the one the range ends in the
// masked fold - columns past the range hold garbage and become all-ones -, chunks past it nothing; ncols is the same for
// the whole warp, so the choice is a uniform branch)
template <int CH>
__device__ __forceinline__ void fold_any(const uint32_t (&v)[32], uint32_t (&p1)[2], uint32_t (&p2)[2], uint32_t m_lo, uint32_t m_hi, int ncols) {
    if (ncols >= CH * 32 + 32) fold_chunk<CH, false>(v, p1, p2, m_lo, m_hi, BT);
    else if (ncols > CH * 32) fold_chunk<CH, true>(v, p1, p2, m_lo, m_hi, ncols);
}
__device__ __forceinline__ void fold_tile(uint32_t taddr, uint32_t (&va)[32], uint32_t (&vb)[32], uint32_t (&p1)[2], uint32_t (&p2)[2], uint32_t m_lo, uint32_t m_hi,
                                          int ncols) {
    tmem_wait_ld(va);
    tmem_ld32(taddr + 32, vb);
    fold_any<0>(va, p1, p2, m_lo, m_hi, ncols);
    tmem_wait_ld(vb);
    tmem_ld32(taddr + 64, va);
    fold_any<1>(vb, p1, p2, m_lo, m_hi, ncols);
    tmem_wait_ld(va);
    tmem_ld32(taddr + 96, vb);
    fold_any<2>(va, p1, p2, m_lo, m_hi, ncols);
    tmem_wait_ld(vb);
}

// low word of a shared-memory matrix descriptor (K-major, SWIZZLE_128B): address / 16 in bits 0..13, LBO = 1 in bits
// 16..29; the high word is the same for every operand (SBO = 64: 8-row groups 1024 B apart, version 1, swizzle mode 2)
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);
template <int ACC>
__device__ __forceinline__ void umma_s8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile("{\n .reg .pred p;\n .reg .b64 da, db;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n setp.ne.b32 p, %5, 0;\n"
                 " tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "n"(ACC)
                 : "memory");
}

template <int NSTAGE, int NA = 2>   // B stages; A buffers (2: the next item's queries load while this item runs)
__global__ void __launch_bounds__(NTHREADS, 1) scan_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                       // [NA buffers][2 A tiles][2 planes][16 KB]
    uint8_t *sB = smem + NA * A_BYTES;        // [NSTAGE][2 planes][16 KB]
    // acc_*[buffer * 2 + A tile]: the two A tiles of a B tile are handed over separately, so the epilogue warps of tile 0
    // run half a tile ahead of those of tile 1 and the two warps of a scheduler are never both waiting
    __shared__ __align__(8) uint64_t a_full[NA], a_empty[NA], b_full[NSTAGE], b_empty[NSTAGE], acc_full[4], acc_empty[4];
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); }
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
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
                const uint32_t ab = it % NA;
                if (!mbar_wait(&a_empty[ab], ((it / NA) & 1) ^ 1, abortp)) break;
                mbar_expect_tx(&a_full[ab], A_BYTES);
#pragma unroll
                for (int at = 0; at < 2; ++at)
#pragma unroll
                    for (int ka = 0; ka < 2; ++ka)
                        bulk_g2s(sA + ab * A_BYTES + (at * 2 + ka) * 16384, p.xq + (size_t)ka * p.xq_plane + ((size_t)sg.q_row0 + at * 128) * 128, 16384, &a_full[ab]);
                const int ntiles = (sg.t_count + BT - 1) / BT;
                const uint8_t *src = p.xt + (size_t)sg.t_row0 * 128;
                for (int tile = 0; tile < ntiles; ++tile, ++T, src += BT * 128) {
                    const uint32_t s = T % NSTAGE;
                    if (!mbar_wait(&b_empty[s], ((T / NSTAGE) & 1) ^ 1, abortp)) break;
                    if (TC_DBG(p, 4)) { mbar_arrive(&b_full[s]); continue; }
                    mbar_expect_tx(&b_full[s], B_BYTES);
                    bulk_g2s(sB + s * B_BYTES, src, 16384, &b_full[s]);
                    bulk_g2s(sB + s * B_BYTES + 16384, src + p.xt_plane, 16384, &b_full[s]);
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ---- MMA issuers, one warp per A tile: a single thread needs ~80 clocks of descriptor arithmetic per
        //      instruction (it issues from a one-lane waterfall loop) and a 128 x 128 x 32 instruction occupies the tensor
        //      pipe for 64, so two threads share the 16 instructions of a B tile.  The whole warp walks the loop, lane 0
        //      issues and commits.  A descriptor is a base computed per tile plus a compile-time constant. ----
        const int at = warp == 1 ? 0 : 1;
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (8u << 24);   // s8 x s8 -> s32, K-major, M = 128; N is set per tile
        const uint32_t a_lo0 = desc_lo(smem_u32(sA)) + at * (32768 >> 4), b_lo0 = desc_lo(smem_u32(sB));
        uint32_t it = 0, T = 0;
        bool ok = true;
        int t_count_next = (int)blockIdx.x < p.n_items ? p.items[blockIdx.x].t_count : 0;
        for (int item = blockIdx.x; item < p.n_items && !*abortp && ok; item += gridDim.x, ++it) {
            const int t_count = t_count_next;
            if (item + (int)gridDim.x < p.n_items) t_count_next = p.items[item + gridDim.x].t_count;   // (a round trip to L2, off the issue path)
            const uint32_t ab = it % NA;
            if (!mbar_wait(&a_full[ab], (it / NA) & 1, abortp)) break;
            const uint32_t a_lo = a_lo0 + ab * (A_BYTES >> 4);
            const int ntiles = (t_count + BT - 1) / BT;
            for (int tile = 0; tile < ntiles; ++tile, ++T) {
                const uint32_t s = T % NSTAGE, cb = T & 1;
                if (!mbar_wait(&b_full[s], (T / NSTAGE) & 1, abortp)) { ok = false; break; }
                if (!mbar_wait(&acc_empty[cb * 2 + at], ((T >> 1) & 1) ^ 1, abortp)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t b_lo = b_lo0 + s * (B_BYTES >> 4);
                    const uint32_t d = tmem + cb * 256 + at * 128;
                    // the last tile of a train range multiplies only the columns that exist (N in steps of 16)
                    const uint32_t n_eff = (uint32_t)min(BT, (t_count - tile * BT + 15) & ~15);
                    const uint32_t idesc_t = idesc | ((n_eff >> 3) << 17);
                    if (!TC_DBG(p, 2)) {
                        umma_s8_lo<0>(d, a_lo, b_lo, idesc_t);
#pragma unroll
                        for (int ks = 1; ks < 8; ++ks) umma_s8_lo<1>(d, a_lo + (ks >> 2) * 1024 + (ks & 3) * 2, b_lo + (ks >> 2) * 1024 + (ks & 3) * 2, idesc_t);
                    }
                    umma_commit(&b_empty[s]);
                    umma_commit(&acc_full[cb * 2 + at]);
                }
                __syncwarp();
            }
            if (ok && lane == 0) umma_commit(&a_empty[ab]);
            __syncwarp();
        }
    } else {
        // ---- epilogue: thread <-> one query row of the block; group g (warps 2..9 / 11..18) owns accumulator buffer g, that
        //      is every second B tile of the CTA, so a scheduler always has four epilogue warps in different phases ----
        const int g = warp < 10 ? 0 : 1, e = warp < 10 ? warp - 2 : warp - 11;
        const int quarter = warp & 3, at = e >> 2;            // (a warp reads the TMEM lanes of its quarter: warp % 4)
        const int lr = at * 128 + quarter * 32 + lane;
        const uint32_t m_lo = p.mul_lo, m_hi = p.mul_hi;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + g * 256 + at * 128;
        uint64_t *full = &acc_full[g * 2 + at], *empty = &acc_empty[g * 2 + at];
        // T counts the B tiles of this CTA over all its items; Walk maps a T to (item, tile)
        struct Walk { int item, ntiles; uint32_t Tbase; int q_valid, out_row0, t_count, t_local0; };
        auto seek = [&](Walk &w, uint32_t T) -> bool {       // move w to the item that holds tile T; false: past the last item
            while (T >= w.Tbase + (uint32_t)w.ntiles) {
                w.Tbase += (uint32_t)w.ntiles;
                w.item += (int)gridDim.x;
                if (w.item >= p.n_items) return false;
                const Item sg = p.items[w.item];
                w.ntiles = (sg.t_count + BT - 1) / BT; w.q_valid = sg.q_valid; w.out_row0 = sg.out_row0; w.t_count = sg.t_count; w.t_local0 = sg.t_local0;
            }
            return true;
        };
        Walk cur;
        cur.item = (int)blockIdx.x - (int)gridDim.x; cur.ntiles = 0; cur.Tbase = 0; cur.q_valid = cur.out_row0 = cur.t_count = cur.t_local0 = 0;
        uint32_t T = (uint32_t)g;
        bool more = seek(cur, T);
        uint32_t best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;   // (distance << 22) | train index, as in the row state
        uint32_t va[32], vb[32];
        bool ok = true;
        if (more) {
            ok = mbar_wait(full, (T >> 1) & 1, abortp);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ok) tmem_ld32(taddr, va);
        }
        while (more && ok) {
            const int tile = (int)(T - cur.Tbase);
            const int ncols = min(BT, cur.t_count - tile * BT);
            // tile-local 16-bit keys, two columns per register (see fold_chunk); two interleaved chains
            uint32_t p1[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, p2[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
            if (TC_DBG(p, 1)) {
                tmem_wait_ld(va);
                p1[0] = va[0] ^ va[31];
            } else {
                fold_tile(taddr, va, vb, p1, p2, m_lo, m_hi, ncols);
            }
            // every load of this accumulator half has returned: hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(empty);
            // this group's next tile (of this item or of a later one): its first load goes under the last fold
            Walk nxt = cur;
            const uint32_t Tn = T + 2;
            more = seek(nxt, Tn);
            if (more) {
                ok = mbar_wait(full, (Tn >> 1) & 1, abortp);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (ok) tmem_ld32(taddr, va);
            }
            if (!TC_DBG(p, 1)) fold_any<3>(vb, p1, p2, m_lo, m_hi, ncols);
            // the tile's two smallest keys -> 32-bit keys with the train index inside the problem
            const uint32_t n1 = __vminu2(p1[0], p1[1]);
            const uint32_t n2 = __vminu2(__vmaxu2(p1[0], p1[1]), __vminu2(p2[0], p2[1]));
            const uint32_t a1 = n1 & 0xFFFFu, b1 = n1 >> 16, a2 = n2 & 0xFFFFu, b2 = n2 >> 16;
            const uint32_t t1 = min(a1, b1), t2 = min(max(a1, b1), min(a2, b2));
            const uint32_t tbase = (uint32_t)(cur.t_local0 + tile * BT);
            const uint32_t k1 = t1 == 0xFFFFu ? 0xFFFFFFFFu : (((t1 >> 7) << DIST_SHIFT_TC) | (tbase + (t1 & 127u)));
            const uint32_t k2 = t2 == 0xFFFFu ? 0xFFFFFFFFu : (((t2 >> 7) << DIST_SHIFT_TC) | (tbase + (t2 & 127u)));
            second = min(second, min(max(best, k1), k2));
            best = min(best, k1);
            if (!more || nxt.item != cur.item) {
                // this group's last tile of the item: the row-state protocol of the POPC kernel (bfm_kernels.cuh, "commit");
                // the other group commits its tiles of the same rows the same way
                if (ok && lr < cur.q_valid && best != 0xFFFFFFFFu && !TC_DBG(p, 8)) {
                    uint32_t *half = reinterpret_cast<uint32_t *>(p.rowstate + (size_t)(cur.out_row0 + lr));
                    const uint32_t displaced = atomicMin(half + 1, best);
                    const uint32_t cand = min(max(displaced, best), second);
                    if (cand != 0xFFFFFFFFu) atomicMin(half, cand);
                }
                best = second = 0xFFFFFFFFu;
            }
            cur = nxt;
            T = Tn;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    if (tid == 0 && s_abort && p.status) *p.status = 1u;
}

}  // namespace bfm_tc
