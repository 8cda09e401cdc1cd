// bfm_kernels.cuh - sm_100a device code of the brute-force Hamming matcher.
//
// What is computed (cv2.BFMatcher(NORM_HAMMING) semantics; the call sites
// this replaces are reference slam/tracking.py:56,121):
//   D[i][j] = popcount(q[i] XOR t[j]) over 256 bits
//   row state[i]  = the two smallest packed keys (D[i][j] << 22 | j) over the allowed j
//   col key[j]    = the smallest packed key (D[i][j] << 22 | i) over the allowed i   (cross-check)
// A single unsigned min over the packed key reproduces cv2's order "distance, then lowest index",
// so every reduction stage (registers, warp REDUX, shared memory, global atomics, multi-GPU) is
// the same associative/commutative min and the result is independent of how work is split.
//
// Mapping to the machine: this is an integer-pipe kernel (LOP3 + POPC + IADD3/IMAD + VIMNMX),
// not a tensor-core one.  Each thread keeps R query descriptors (8 x u32 each) in registers,
// the CTA streams a range of train descriptors through shared memory in 4 KB chunks that a
// single thread fetches with 1-D TMA bulk copies (cp.async.bulk + mbarrier, double buffered),
// and every lane reads the same train words (LDS.128 broadcast, conflict free).  HBM traffic is
// ~4e-2 bytes per pair; the bound is the POPC issue rate.
//
// One launch does everything: the CTA that completes a problem's last work item finalizes it
// (finalize_problem: knn table, cross-check / ratio / gate, ordered match list, workspace reset) and
// writes the results to up to eight destinations - device memory, pinned host memory, NVLink peers or an
// NVSwitch multicast address.  On the host path the first CTAs of the same grid stream the caller's pinned
// arrays into HBM (feed_rows) while the others wait at an input gate for the rows they read.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bfm {

constexpr int DIST_SHIFT = 22;
constexpr uint32_t IDX_MASK = (1u << DIST_SHIFT) - 1u;
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;      // empty slot
constexpr uint32_t DIST_MASKED = 511u;          // distance given to a masked pair (real distances are 0..256)
constexpr uint32_t KEY_DEAD = DIST_MASKED << DIST_SHIFT;  // any key >= this is "no candidate"
constexpr uint32_t KEY16_DEAD = DIST_MASKED << 7;          // same, for the chunk-local 16-bit keys (d << 7 | j)
constexpr int TT = 128;                         // train rows per shared-memory chunk (4 KB)

// One work item: a block of BQ = NT*R query rows against a contiguous range of train rows of
// the same problem.  Built on the host by the planner (bfm_api.cu: plan_segments).  Every problem
// has at least one segment (an empty problem gets one with t_count == 0) because the CTA that
// completes a problem's last segment also finalizes it.
struct __align__(16) Segment {
    int32_t q_row0;    // first query row (global row in the query array)
    int32_t q_valid;   // rows of this block that exist (0..BQ)
    int32_t q_local0;  // index of q_row0 inside its problem (cv2 queryIdx of the first row)
    int32_t out_row0;  // output row of the first query of the block
    int32_t t_row0;    // first train row (global row in the train array)
    int32_t t_count;   // train rows in this segment (0 only for an empty problem)
    int32_t t_local0;  // index of t_row0 inside its problem (cv2 trainIdx)
    int32_t problem;   // index into the problem table
};

// One finalize tile of the persistent form: FT_TILE_ROWS query rows of one problem.
struct __align__(16) FinTile {
    int32_t problem;
    int32_t row0;        // first row of the tile inside its problem
    int32_t index;       // tile number inside its problem (0 .. n_tiles - 1)
    int32_t n_tiles;     // tiles of the problem
    int32_t slot0;       // fin_count slot of the problem's tile 0
    int32_t pad[3];
};

struct __align__(16) Problem {   // device view of bfm_problem_t
    int32_t q_begin, q_count, t_begin, t_count, out_begin;
    int32_t col0;      // first column-key slot of this problem (cross-check)
    int32_t n_segs;    // segments of this problem (>= 1)
    int32_t pad;
};

constexpr int MAX_DEST = 8;  // result replicas (this GPU + NVLink peers)
constexpr int FEED_MAX = 32; // feeder CTAs of the SM-fed upload (their progress words share one 128-byte line)
constexpr int FEED_HEAD = 8; // the first round of the SM-fed upload is delivered as this many short rounds

// Everything one launch needs.  The kernel is scan + finalize in one: there is no second launch.
struct ScanParams {
    const uint4 *q;                  // [rows][2] 16-byte halves of the 32-byte descriptors
    const uint4 *t;
    const Segment *segs;
    const Problem *problems;
    unsigned long long *rowstate;    // [out rows] (best key << 32) | second key
    uint32_t *colkeys;               // [sum of problem train rows] (cross-check), problem base = Problem::col0
    uint32_t *done;                  // [problems] completed-segment counters, all-ones when idle
    // input gating (pipelined host path): a CTA starts once the copy engine has landed the rows it
    // reads; ready[0] / ready[1] = ready_base + query / train rows uploaded so far.  NULL = resident.
    const unsigned long long *ready;
    unsigned long long ready_base;
    uint32_t *status;                // set non-zero if the gate timed out (host reports the error)
    // SM-fed upload (host path with pinned inputs): the first n_feed CTAs of the grid stream the caller's
    // arrays from pinned host memory into HBM in `feed_rounds` rounds and publish their progress; every
    // other CTA waits until all feeders have passed the round that covers its rows.  See feed_rows().
    int32_t n_feed;
    int32_t feed_rounds;
    int32_t feed_q_rows, feed_t_rows;        // rows of the query / train arrays delivered per round
    const uint4 *feed_src[4];                // host-mapped: q desc, t desc, q_xy, t_xy (NULL = absent)
    uint4 *feed_dst[4];                      // device copies
    unsigned long long feed_bytes[4];        // total bytes of each array
    uint32_t *feed_prog;                     // [FEED_MAX] (epoch << 16 | rounds completed) per feeder CTA
    uint32_t feed_epoch;                     // this call's epoch (1..65535): words of earlier calls compare lower,
                                             // so the progress words need no reset between calls
    int32_t feed_stall;                      // test hook: feeders deliver nothing, so the gate's time-out path runs
    const uint32_t *feed_host_ready;         // pinned host word: rounds staged by the host's worker threads so far
                                             // (pageable caller arrays); NULL = the sources are complete
    // Persistent form (resident inputs): the grid is at most one wave; CTA c starts with work item c and draws
    // further items from the ticket counter `queue`; when the queue is empty it finalizes tiles
    // [cta_tiles[c].x, +.y) of fin_tiles.  A tile waits for its problem's items (all drawn by then, each by a
    // running CTA) and for the tiles before it, which belong to CTAs with a lower or equal index - dispatched
    // no later than this one.  queue == NULL: one work item per CTA, the CTA that completes a problem finalizes
    // it (the gated host path).
    uint32_t *queue, *queue_other;   // this launch's ticket counter (zero on entry) and the next launch's (zeroed here)
    int32_t n_items, n_ctas;
    const int2 *cta_tiles;
    const struct FinTile *fin_tiles;
    uint32_t *fin_done;              // [problems] finalized-tile counters, all-ones when idle
    uint32_t *fin_count;             // [tiles] (epoch << 12 | rows kept by the tile): look-back of the ordered compaction
    uint32_t fin_epoch;              // this call's epoch (1 .. 2^20 - 1); words of earlier calls never match
    const int32_t *t_limit;          // optional device scalar: train rows that exist (single problem), else NULL
    int32_t limit_segs;              // with t_limit: work items per query block (the rows that exist are re-cut over them)
    const uint8_t *mask;             // dense mask (single problem): [q_local][mask_stride]
    long long mask_stride;
    const float2 *q_xy;              // window: pixel coordinates per query / train row
    const float2 *t_xy;
    float radius;
    // multipliers for key formation, passed as data so that ptxas keeps them IMADs (FMA pipe)
    // instead of strength-reducing to LEA/SHF on the ALU pipe, which is the binding one
    uint32_t mul_d32;   // 1 << 22
    uint32_t mul_lo16;  // 1 << 7
    uint32_t mul_hi16;  // 1 << 23
    uint32_t mul_one, mul_two, mul_four;  // weights of the carry-save popcount sum, same reason
    // ---- k > 2: one pass per two neighbours.  Pass n only admits keys strictly greater than the last
    //      key pass n-1 found for that row (keys are unique per row, so this is exact exclusion).
    const uint32_t *lower;           // [out rows] or NULL (first pass)
    uint32_t *lower_out;             // [out rows] or NULL: finalize stores the row's second key of this pass
    int32_t knn_col0, knn_cols;      // this pass fills columns [knn_col0, knn_col0 + knn_cols) of the knn table
    int32_t defer_finalize;          // static form, 1: the scan only reduces; bfm_tiles_kernel finalizes tile by tile
    // ---- finalize (run by the CTA that completes a problem's last segment) ----------------------
    int32_t k;             // columns (row stride) of the knn table
    int32_t cross_check;
    int32_t max_distance;  // < 0 off
    int32_t use_ratio;
    double ratio;
    // Result destinations.  dest[0] is the caller's buffers (device memory, or pinned host memory
    // written over PCIe on the host path); dest[1..] are the same buffers of NVLink peers
    // (multi-GPU gather fused into the epilogue).  Each pointer may be NULL.
    // debug timeline (bfm_debug_timeline): CTA b of the launch stores %globaltimer at up to 8 points of its life into
    // trace[(trace_base + b) * 8 ..]; NULL (the normal case) costs one uniform predicate per point
    unsigned long long *trace;
    int32_t trace_base, trace_cap;
    int32_t n_dest;
    uint32_t dest_multicast;   // bit d set: dest[d] holds NVSwitch multicast addresses (one multimem.st reaches every GPU)
    struct Dest {
        int32_t *knn_idx, *knn_dist;                 // [out rows][k]
        int32_t *m_query, *m_train, *m_dist;         // [out rows]
        int32_t *m_count;                            // [problems]
    } dest[MAX_DEST];
    // (new fields go to the END of this struct: even the offsets of the fields above change the register assignment
    // ptxas gives the static kernel's inner loop, tests/test_hotloop_fingerprint.py)
};

// ---- PTX helpers: mbarrier + 1-D TMA bulk copy ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// debug timeline: one thread per CTA stamps point `slot` (0..7); see ScanParams::trace
__device__ __forceinline__ void trace_mark(const ScanParams &p, int slot) {
#ifdef BFM_AB_NOTRACE
    (void)p; (void)slot; return;
#endif
    if (p.trace != nullptr && threadIdx.x == 0 && (int)blockIdx.x + p.trace_base < p.trace_cap)
        p.trace[(size_t)(p.trace_base + (int)blockIdx.x) * 8 + slot] = global_timer_ns();
}

// ---- 256-bit Hamming distance ------------------------------------------------------------------
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t max_u16x2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// PM selects how the 256-bit popcount is evaluated (all variants are exact):
//   8        plain: 8 XOR + 8 POPC
//   6, 5, 4  carry-save adder (Harley-Seal) tree on LOP3 over the 8 XOR words, 6/5/4 POPCs
//   50, 40   the same trees over a TRANSFORMED descriptor.  Both sides are stored as
//              (w0, w1, w3, w4, w7, w0^w1^w2, w3^w4^w5, w0^...^w6)
//            which is linear over XOR, so the sum word of every first/second level adder is ONE
//            XOR of precomputed words, and the carry is a LOP3 of (x_a, x_b, sum) because the
//            third input is recoverable: maj(a, b, a^b^s) = LUT 0xD4.  11 (PM 50) or 13 (PM 40)
//            LOP3 per pair instead of 14 / 16, and w2, w5, w6 are never formed.
// POPC issues at 16 lanes/clk/SM and LOP3 at 64 (measured, profiles/popc_peak_r01.json), so the
// trade is worth it until the ALU pipe becomes the limiter.
__device__ __forceinline__ uint32_t carry_from_sum(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r;  // maj(a, b, a ^ b ^ s)
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}

__host__ __device__ constexpr bool pm_transformed(int pm) { return pm >= 40; }

// in: the 8 words of a descriptor; out: the transformed descriptor (see above)
__device__ __forceinline__ void transform_desc(uint32_t (&w)[8]) {
    const uint32_t s0 = xor3(w[0], w[1], w[2]);
    const uint32_t s1 = xor3(w[3], w[4], w[5]);
    const uint32_t s2 = xor3(s0, s1, w[6]);
    const uint32_t w3 = w[3], w4 = w[4], w7 = w[7];
    w[2] = w3; w[3] = w4; w[4] = w7; w[5] = s0; w[6] = s1; w[7] = s2;
}

template <int PM>
__device__ __forceinline__ uint32_t hamming256(const uint32_t (&q)[8], const uint32_t (&t)[8], const ScanParams &p) {
    if constexpr (pm_transformed(PM)) {
        // layout: [0]=w0 [1]=w1 [2]=w3 [3]=w4 [4]=w7 [5]=S012 [6]=S345 [7]=S0..6
        const uint32_t x0 = q[0] ^ t[0], x1 = q[1] ^ t[1], s0 = q[5] ^ t[5];
        const uint32_t c0 = carry_from_sum(x0, x1, s0);
        const uint32_t x3 = q[2] ^ t[2], x4 = q[3] ^ t[3], s1 = q[6] ^ t[6];
        const uint32_t c1 = carry_from_sum(x3, x4, s1);
        const uint32_t s2 = q[7] ^ t[7];
        const uint32_t c2 = carry_from_sum(s0, s1, s2);   // maj(s0, s1, x6)
        const uint32_t x7 = q[4] ^ t[4];
        if constexpr (PM == 50) {
            return (__popc(s2) + __popc(x7)) + 2u * (__popc(c0) + __popc(c1) + __popc(c2));
        } else {  // PM == 40: the weighted sum as three IMADs (FMA pipe; the ALU pipe is the binding one)
            const uint32_t s3 = xor3(c0, c1, c2), c3 = maj3(c0, c1, c2);
            return __popc(c3) * p.mul_four + (__popc(s3) * p.mul_two + (__popc(s2) * p.mul_one + __popc(x7)));
        }
    } else {
        uint32_t x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = q[i] ^ t[i];
        if constexpr (PM == 8) {
            return (__popc(x[0]) + __popc(x[1]) + __popc(x[2])) + (__popc(x[3]) + __popc(x[4]) + __popc(x[5])) +
                   (__popc(x[6]) + __popc(x[7]));
        } else {
            const uint32_t s0 = xor3(x[0], x[1], x[2]), c0 = maj3(x[0], x[1], x[2]);
            const uint32_t s1 = xor3(x[3], x[4], x[5]), c1 = maj3(x[3], x[4], x[5]);
            if constexpr (PM == 6) {
                return (__popc(s0) + __popc(s1) + __popc(x[6])) + __popc(x[7]) + 2u * (__popc(c0) + __popc(c1));
            } else {
                const uint32_t s2 = xor3(s0, s1, x[6]), c2 = maj3(s0, s1, x[6]);
                if constexpr (PM == 5) {
                    return (__popc(s2) + __popc(x[7])) + 2u * (__popc(c0) + __popc(c1) + __popc(c2));
                } else {  // PM == 4
                    const uint32_t s3 = xor3(c0, c1, c2), c3 = maj3(c0, c1, c2);
                    return (__popc(s2) + __popc(x[7])) + 2u * __popc(s3) + 4u * __popc(c3);
                }
            }
        }
    }
}

// A result store.  Ordinary destinations take a plain store; a multicast destination (an address of an
// NVSwitch multicast object spanning the symmetric buffers of all ranks) takes multimem.st, which the switch
// replicates into every GPU's copy: one store instead of one per peer.
__device__ __forceinline__ void put_i32(int32_t *p, int v, bool multicast) {
    if (multicast) asm volatile("multimem.st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else *p = v;
}

// ---- finalize: decode keys, cross-check / ratio / distance gate, ordered compaction -----------------
// Run by ONE CTA per problem: the one whose segment completed the problem (last-arriver pattern,
// see the kernel tail).  Decodes the packed keys into the dense cv2 knnMatch table, applies
// cross-check (colkey[trainIdx] == (d, queryIdx)), the ratio test in fp64 exactly as Python
// evaluates `m.distance < ratio * n.distance`, the distance gate, and compacts the surviving
// matches in ascending queryIdx order.  Every workspace slot it reads is reset to all-ones, so the
// workspace is self-cleaning and a steady-state call needs no memset.  Results go to every
// destination in p.dest (plain stores: device memory, pinned host memory or NVLink peer memory).
constexpr int FIN_RPT = 8;   // rows per thread and tile: the CTA that completes a problem walks it in tiles of NT * 8 rows
constexpr int FT_RPT = 4;    // persistent form: tiles of NT * 4 rows, each finalized by the CTA the plan names

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t *p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Rows [base, base + NT * RPT) of problem `pi`: decode, knn table, keep decisions, ordered compaction, row-state reset.
// `running` = matches kept by the rows before `base`.  LOOKBACK: the rows before `base` belong to other tiles (other
// CTAs): this tile publishes its own count in fin_count[slot0 + index] and sums the counts of the tiles before it
// (they are finalized by CTAs dispatched no later than this one, so the wait cannot deadlock).  Returns the tile's count.
template <int NT, int RPT, bool LOOKBACK>
__device__ __forceinline__ int finalize_rows(const ScanParams &p, const Problem &pr, const int pi, const int base, int running,
                                             int (*s_cnt)[NT / 32], const FinTile *ft, uint32_t *arrivals = nullptr) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int k = p.k;
    int idx1[RPT], d1[RPT];
    bool keep[RPT];
    uint32_t bal[RPT];
    // phase 1: all row-state loads of the tile in flight together, then the resets
    unsigned long long st[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int i = base + j * NT + tid;
        st[j] = i < pr.q_count ? __ldcg(p.rowstate + (size_t)pr.out_begin + i) : ~0ull;
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int i = base + j * NT + tid;
        if (i < pr.q_count) p.rowstate[(size_t)pr.out_begin + i] = ~0ull;
    }
    // phase 2: all column-key loads (cross-check) in flight together
    uint32_t ck[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const uint32_t k1 = (uint32_t)(st[j] >> 32);
        ck[j] = KEY_NONE;
        if (p.cross_check && k1 < KEY_DEAD) ck[j] = __ldcg(p.colkeys + (size_t)pr.col0 + (k1 & IDX_MASK));
    }
    // phase 3: decode, knn table, keep decisions
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const int i = base + j * NT + tid;
        const bool in = i < pr.q_count;
        const uint32_t k1 = (uint32_t)(st[j] >> 32), k2 = (uint32_t)st[j];
        const bool has1 = k1 < KEY_DEAD, has2 = k2 < KEY_DEAD;
        idx1[j] = has1 ? (int)(k1 & IDX_MASK) : -1;
        d1[j] = has1 ? (int)(k1 >> DIST_SHIFT) : -1;
        const int idx2 = has2 ? (int)(k2 & IDX_MASK) : -1, d2 = has2 ? (int)(k2 >> DIST_SHIFT) : -1;
        if (in) {
            const size_t o = ((size_t)pr.out_begin + i) * (size_t)k + (size_t)p.knn_col0;
            for (int d = 0; d < p.n_dest; ++d) {
                int32_t *ki = p.dest[d].knn_idx, *kd = p.dest[d].knn_dist;
                if (!ki) continue;
                const bool mc = (p.dest_multicast >> d) & 1u;
                if (k == 2 && !mc) {   // 8-byte aligned: o is even
                    *reinterpret_cast<int2 *>(ki + o) = make_int2(idx1[j], idx2);
                    *reinterpret_cast<int2 *>(kd + o) = make_int2(d1[j], d2);
                } else {
                    put_i32(ki + o, idx1[j], mc);
                    put_i32(kd + o, d1[j], mc);
                    if (p.knn_cols > 1) { put_i32(ki + o + 1, idx2, mc); put_i32(kd + o + 1, d2, mc); }
                }
            }
            if (p.lower_out) p.lower_out[(size_t)pr.out_begin + i] = has2 ? k2 : KEY_NONE;
        }
        bool kp = in && has1;
        if (kp && p.cross_check) kp = ck[j] == (((uint32_t)d1[j] << DIST_SHIFT) | (uint32_t)i);
        if (kp && p.use_ratio) kp = has2 && ((double)d1[j] < p.ratio * (double)d2);
        if (kp && p.max_distance >= 0) kp = d1[j] <= p.max_distance;
        keep[j] = kp;
        bal[j] = __ballot_sync(0xffffffffu, kp);
        if (lane == 0) s_cnt[j][warp] = __popc(bal[j]);
    }
    __syncthreads();
    // ordered compaction: rows ascend with (j, warp, lane)
    int before = 0, tile_total = 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = s_cnt[j][w];
            tile_total += c;
            if (j == 0 && w < warp) before += c;
        }
    }
    if (LOOKBACK) {
        // publish this tile's count, then add up the tiles before it (every thread sums them all: index <= a few dozen
        // for any real frame; the words sit in L2)
        if (tid == 0) {
            *(volatile uint32_t *)(p.fin_count + ft->slot0 + ft->index) = (p.fin_epoch << 12) | (uint32_t)tile_total;
            // Every read this tile makes of the problem's column keys and row states has returned (the keep
            // decisions behind the barrier above consumed them): count the tile as done reading NOW, so the round
            // trip of the atomic runs under the look-back and the match stores instead of after them.
            *arrivals = atomicAdd(p.fin_done + pi, 1u);   // idles at all-ones
        }
        const uint32_t tag = p.fin_epoch << 12;
        for (int t = lane; t < ft->index; t += 32) {
            uint32_t v;
            unsigned long long t0 = 0;
            while (((v = *(volatile const uint32_t *)(p.fin_count + ft->slot0 + t)) & 0xFFFFF000u) != tag) {
                __nanosleep(64);
                if (t0 == 0) t0 = global_timer_ns();
                else if (global_timer_ns() - t0 > 2000000000ull) { v = 0; break; }   // never hang the GPU
            }
            running += (int)(v & 0xFFFu);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) running += __shfl_xor_sync(0xffffffffu, running, o);
    }
    int pos_j = running + before;   // rank of this warp's first kept row of slot j = 0
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        if (j > 0) {
            // rows of slot j come after every row of slot j-1: add the rest of slot j-1 and the
            // warps before this one in slot j
            int add = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if (w >= warp) add += s_cnt[j - 1][w];
                if (w < warp) add += s_cnt[j][w];
            }
            pos_j += add;
        }
        if (keep[j]) {
            const size_t o = (size_t)pr.out_begin + pos_j + __popc(bal[j] & lt);
            const int i = base + j * NT + tid;
            for (int d = 0; d < p.n_dest; ++d) {
                if (!p.dest[d].m_count) continue;
                const bool mc = (p.dest_multicast >> d) & 1u;
                put_i32(p.dest[d].m_query + o, i, mc);
                put_i32(p.dest[d].m_train + o, idx1[j], mc);
                put_i32(p.dest[d].m_dist + o, d1[j], mc);
            }
        }
    }
    if (LOOKBACK && ft->index == ft->n_tiles - 1 && tid == 0)
        for (int d = 0; d < p.n_dest; ++d)
            if (p.dest[d].m_count) put_i32(p.dest[d].m_count + pi, running + tile_total, (p.dest_multicast >> d) & 1u);
    (void)pi;
    __syncthreads();   // s_cnt is rewritten by the next tile
    return tile_total;
}

template <int NT>
__device__ __noinline__ void finalize_problem(const ScanParams &p, const int pi, int (*s_cnt)[NT / 32]) {
    const Problem pr = p.problems[pi];
    const int tid = threadIdx.x;
    int running = 0;
    for (int base = 0; base < pr.q_count; base += NT * FIN_RPT)
        running += finalize_rows<NT, FIN_RPT, false>(p, pr, pi, base, running, s_cnt, nullptr);
    if (tid == 0)
        for (int d = 0; d < p.n_dest; ++d)
            if (p.dest[d].m_count) put_i32(p.dest[d].m_count + pi, running, (p.dest_multicast >> d) & 1u);
    if (p.cross_check) {
        // every column-key read of this problem happened above, in this CTA
        for (int j = tid; j < pr.t_count; j += NT) p.colkeys[(size_t)pr.col0 + j] = KEY_NONE;
    }
}

// Persistent form: one tile of a problem.  Waits until every work item of the problem has been committed (the done
// counter idles at all-ones, so n arrivals read n - 1), finalizes the tile's rows, and the tile that finishes last
// restores the problem's column keys and counters - the workspace stays self-cleaning.
template <int NT>
__device__ __noinline__ void finalize_tile(const ScanParams &p, const FinTile *ftp, int (*s_cnt)[NT / 32], int *s_flag) {
    const int tid = threadIdx.x;
    const int pi = ftp->problem;
    const Problem pr = p.problems[pi];
    if (tid == 0) {
        const uint32_t want = (uint32_t)pr.n_segs - 1u;
        unsigned long long t0 = 0;
        while (ld_acquire_gpu(p.done + pi) != want) {
            __nanosleep(100);
            if (t0 == 0) t0 = global_timer_ns();
            else if (global_timer_ns() - t0 > 2000000000ull) break;   // never hang the GPU
        }
    }
    __syncthreads();
    uint32_t old = 0;
    finalize_rows<NT, FT_RPT, true>(p, pr, pi, ftp->row0, 0, s_cnt, ftp, &old);
    // the tile that arrives last restores what all tiles of the problem have read (no fence: see finalize_rows)
    if (tid == 0) *s_flag = (old == (uint32_t)ftp->n_tiles - 2u) ? 1 : 0;
    __syncthreads();
    if (*s_flag) {
        if (tid == 0) {
            p.done[pi] = 0xFFFFFFFFu;
            p.fin_done[pi] = 0xFFFFFFFFu;
        }
        if (p.cross_check) {
            uint32_t *ck = p.colkeys + (size_t)pr.col0;
            int j = tid;
            // 16-byte stores over the aligned middle
            const int head = (int)(((16u - ((uint32_t)(uintptr_t)ck & 15u)) & 15u) >> 2);
            for (; j < min(head, pr.t_count); j += NT) ck[j] = KEY_NONE;
            const int n4 = (pr.t_count - min(head, pr.t_count)) >> 2;
            uint4 *ck4 = reinterpret_cast<uint4 *>(ck + min(head, pr.t_count));
            for (int q4 = tid; q4 < n4; q4 += NT) ck4[q4] = make_uint4(KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE);
            for (int r = min(head, pr.t_count) + 4 * n4 + tid; r < pr.t_count; r += NT) ck[r] = KEY_NONE;
        }
    }
    __syncthreads();   // s_flag is rewritten by the next tile
}

// ---- SM-fed upload ---------------------------------------------------------------------------------------
// A feeder CTA copies, round after round, its share of the next slice of every input array from pinned
// host memory (zero-copy loads over PCIe, four 16-byte loads in flight per thread) into HBM, then
// publishes "round r done".  No copy-engine operation, event or host involvement per slice, so the
// slices can be as fine as a keyframe pair and the matching CTAs start microseconds after the launch.
template <int NT>
__device__ __noinline__ void feed_rows(const ScanParams &p) {
    const int feeder = blockIdx.x, tid = threadIdx.x;
    const unsigned long long stride = (unsigned long long)p.n_feed * NT;       // 16-byte words per sweep
    const unsigned long long me = (unsigned long long)feeder * NT + tid;
    __shared__ int s_go;
    uint32_t staged = 0;   // rounds the host is known to have staged
    for (int r = 0; r < p.feed_rounds; ++r) {
        if (p.feed_host_ready != nullptr && (uint32_t)r >= staged) {
            // the host is still staging: wait for round r.  A poll is a PCIe read that queues behind the bulk
            // loads of every feeder, so the value is remembered and polled again only when it runs out.
            if (tid == 0) {
                const unsigned long long t0 = global_timer_ns();
                uint32_t v;
                while ((v = *(volatile const uint32_t *)p.feed_host_ready) < (uint32_t)(r + 1)) {
                    __nanosleep(500);
                    if (global_timer_ns() - t0 > 4000000000ull) { v = 0; break; }
                }
                s_go = (int)v;
            }
            __syncthreads();
            staged = (uint32_t)s_go;
            __syncthreads();
            if (staged == 0) return;   // timed out: the waiting CTAs time out on their own and report it
        }
        // pairs of arrays (descriptors, then pixel coordinates): the loads of both are in flight together
#pragma unroll
        for (int pair = 0; pair < 2; ++pair) {
            unsigned long long w0[2], w1[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int a = 2 * pair + s;
                const unsigned long long rows = s ? (unsigned long long)p.feed_t_rows : (unsigned long long)p.feed_q_rows;
                // rows per round is a multiple of 128; the first FEED_HEAD rounds are eighths of a round, so the
                // matching CTAs of the first problems start after a few microseconds of upload
                const unsigned long long per_round = rows * (pair == 0 ? 32ull : 8ull);
                const unsigned long long b0 = r < FEED_HEAD ? (unsigned long long)r * (per_round / FEED_HEAD)
                                                            : (unsigned long long)(r - FEED_HEAD + 1) * per_round;
                const unsigned long long b1 = r < FEED_HEAD ? b0 + per_round / FEED_HEAD : b0 + per_round;
                const unsigned long long lo = min(b0, p.feed_bytes[a]);
                const unsigned long long hi = min(b1, p.feed_bytes[a]);
                w0[s] = lo >> 4;
                w1[s] = p.feed_src[a] ? (hi >> 4) : w0[s];
                if (p.feed_src[a] && hi == p.feed_bytes[a] && (hi & 15) && hi > lo && me == 0) {
                    // odd number of coordinate rows: the last 8 bytes (never read past the caller's array)
                    const uint2 v = *reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(p.feed_src[a]) + (hi & ~15ull));
                    *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(p.feed_dst[a]) + (hi & ~15ull)) = v;
                }
            }
            const unsigned long long span = max(w1[0] - w0[0], w1[1] - w0[1]);
            for (unsigned long long o = me; o < span; o += 4 * stride) {
                uint4 v[2][4];
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (w0[s] + o + k * stride < w1[s]) v[s][k] = __ldcs(p.feed_src[2 * pair + s] + w0[s] + o + k * stride);
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (w0[s] + o + k * stride < w1[s]) p.feed_dst[2 * pair + s][w0[s] + o + k * stride] = v[s][k];
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            volatile uint32_t *prog = p.feed_prog + feeder;
            *prog = (p.feed_epoch << 16) | (uint32_t)(r + 1);
        }
    }
}

// ---- the distance-scan kernel --------------------------------------------------------------------
// R     queries per thread (register tile)
// K     1 or 2 neighbours tracked per query
// CROSS also reduce the per-train column key (cross-check); K must be 1
// MASK  0 none, 1 dense uint8 mask, 2 projection window
// PM    popcount evaluation (see hamming256)
// NT    threads per CTA (== TT: one thread transforms one staged train row)
// BOUND only admit keys above a per-row lower bound (passes 2.. of knnMatch with k > 2; R == 1)
// DYN   persistent form (resident inputs) / static form (the gated host path), see below
//
// Static form: CTA b takes work item b (after the feeders): chunk c+2 of its train range is fetched by TMA while chunk
// c is scanned, one __syncthreads per chunk; the CTA that completes a problem finalizes it.  CTAs [0, n_feed) are
// feeders (host path with pinned inputs).
//
// Persistent form: the grid is at most one wave.  A CTA starts with item `cta` and draws its next items from a
// device-wide ticket counter (the ticket is drawn one item ahead, so its latency hides behind the scan), which keeps
// every SM supplied until the queue is empty whatever pace its CTAs run at; consecutive items of one query block keep
// the queries and the running keys in registers (one commit per run); when the queue is empty the CTA finalizes the
// tiles the plan gave it.  Measured against two alternatives (profiles/r02_kernel_forms.md): equal static shares per
// CTA (stream-K) lose 10 % because the warp schedulers serve the CTAs of an SM at rates up to 3x apart and the slow
// ones finish alone, and a fully overlapped chunk stream (tickets, descriptors and TMA several items ahead) loses to
// this form because CTAs that never stall starve their neighbours - the per-item stall is what shares the pipe.
// Shared memory: 2 x 4 KB train rows, 2 x 1 KB pixel coords (window), 2 x 2 KB column keys (cross-check).
// (the round-1 finalize of the static form, kept with it: see the note below)
template <int NT>
__device__ __noinline__ void finalize_problem_v1(const ScanParams &p, const int pi, int (*s_cnt)[NT / 32]) {
    constexpr int NW = NT / 32;
    const Problem pr = p.problems[pi];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int k = p.k;
    int running = 0;
    for (int base = 0; base < pr.q_count; base += NT * FIN_RPT) {
        int idx1[FIN_RPT], d1[FIN_RPT];
        bool keep[FIN_RPT];
        uint32_t bal[FIN_RPT];
        // phase 1: all row-state loads of the tile in flight together, then the resets
        unsigned long long st[FIN_RPT];
#pragma unroll
        for (int j = 0; j < FIN_RPT; ++j) {
            const int i = base + j * NT + tid;
            st[j] = i < pr.q_count ? __ldcg(p.rowstate + (size_t)pr.out_begin + i) : ~0ull;
        }
#pragma unroll
        for (int j = 0; j < FIN_RPT; ++j) {
            const int i = base + j * NT + tid;
            if (i < pr.q_count) p.rowstate[(size_t)pr.out_begin + i] = ~0ull;
        }
        // phase 2: all column-key loads (cross-check) in flight together
        uint32_t ck[FIN_RPT];
#pragma unroll
        for (int j = 0; j < FIN_RPT; ++j) {
            const uint32_t k1 = (uint32_t)(st[j] >> 32);
            ck[j] = KEY_NONE;
            if (p.cross_check && k1 < KEY_DEAD) ck[j] = __ldcg(p.colkeys + (size_t)pr.col0 + (k1 & IDX_MASK));
        }
        // phase 3: decode, knn table, keep decisions
#pragma unroll
        for (int j = 0; j < FIN_RPT; ++j) {
            const int i = base + j * NT + tid;
            const bool in = i < pr.q_count;
            const uint32_t k1 = (uint32_t)(st[j] >> 32), k2 = (uint32_t)st[j];
            const bool has1 = k1 < KEY_DEAD, has2 = k2 < KEY_DEAD;
            idx1[j] = has1 ? (int)(k1 & IDX_MASK) : -1;
            d1[j] = has1 ? (int)(k1 >> DIST_SHIFT) : -1;
            const int idx2 = has2 ? (int)(k2 & IDX_MASK) : -1, d2 = has2 ? (int)(k2 >> DIST_SHIFT) : -1;
            if (in) {
                const size_t o = ((size_t)pr.out_begin + i) * (size_t)k + (size_t)p.knn_col0;
                for (int d = 0; d < p.n_dest; ++d) {
                    int32_t *ki = p.dest[d].knn_idx, *kd = p.dest[d].knn_dist;
                    if (!ki) continue;
                    const bool mc = (p.dest_multicast >> d) & 1u;
                    if (k == 2 && !mc) {   // 8-byte aligned: o is even
                        *reinterpret_cast<int2 *>(ki + o) = make_int2(idx1[j], idx2);
                        *reinterpret_cast<int2 *>(kd + o) = make_int2(d1[j], d2);
                    } else {
                        put_i32(ki + o, idx1[j], mc);
                        put_i32(kd + o, d1[j], mc);
                        if (p.knn_cols > 1) { put_i32(ki + o + 1, idx2, mc); put_i32(kd + o + 1, d2, mc); }
                    }
                }
                if (p.lower_out) p.lower_out[(size_t)pr.out_begin + i] = has2 ? k2 : KEY_NONE;
            }
            bool kp = in && has1;
            if (kp && p.cross_check) kp = ck[j] == (((uint32_t)d1[j] << DIST_SHIFT) | (uint32_t)i);
            if (kp && p.use_ratio) kp = has2 && ((double)d1[j] < p.ratio * (double)d2);
            if (kp && p.max_distance >= 0) kp = d1[j] <= p.max_distance;
            keep[j] = kp;
            bal[j] = __ballot_sync(0xffffffffu, kp);
            if (lane == 0) s_cnt[j][warp] = __popc(bal[j]);
        }
        __syncthreads();
        // ordered compaction: rows ascend with (j, warp, lane)
        int before = running, tile_total = 0;
#pragma unroll
        for (int j = 0; j < FIN_RPT; ++j) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int c = s_cnt[j][w];
                tile_total += c;
                if (j == 0 && w < warp) before += c;
            }
        }
        int pos_j = before;   // rank of this warp's first kept row of slot j = 0
#pragma unroll
        for (int j = 0; j < FIN_RPT; ++j) {
            if (j > 0) {
                // rows of slot j come after every row of slot j-1: add the rest of slot j-1 and the
                // warps before this one in slot j
                int add = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    if (w >= warp) add += s_cnt[j - 1][w];
                    if (w < warp) add += s_cnt[j][w];
                }
                pos_j += add;
            }
            if (keep[j]) {
                const size_t o = (size_t)pr.out_begin + pos_j + __popc(bal[j] & lt);
                const int i = base + j * NT + tid;
                for (int d = 0; d < p.n_dest; ++d) {
                    if (!p.dest[d].m_count) continue;
                    const bool mc = (p.dest_multicast >> d) & 1u;
                    put_i32(p.dest[d].m_query + o, i, mc);
                    put_i32(p.dest[d].m_train + o, idx1[j], mc);
                    put_i32(p.dest[d].m_dist + o, d1[j], mc);
                }
            }
        }
        running += tile_total;
        __syncthreads();   // s_cnt is rewritten by the next tile
    }
    if (tid == 0)
        for (int d = 0; d < p.n_dest; ++d)
            if (p.dest[d].m_count) put_i32(p.dest[d].m_count + pi, running, (p.dest_multicast >> d) & 1u);
    if (p.cross_check) {
        // every column-key read of this problem happened above, in this CTA
        for (int j = tid; j < pr.t_count; j += NT) p.colkeys[(size_t)pr.col0 + j] = KEY_NONE;
    }
}


// ---- static form ---------------------------------------------------------------------------------
// (kept exactly as measured in round 1: the same source inside a larger kernel gives the inner loop a different
// instruction schedule - same 194 instructions - that is 1.5 % slower on the headline batch, tools/ab_old_new.py)
template <int R, int K, bool CROSS, int MASK, int PM, int NT, bool BOUND>
__global__ void __launch_bounds__(NT, MASK == 0 ? 8 : 6) bfm_scan_static_kernel(const __grid_constant__ ScanParams p) {
    static_assert(!BOUND || R == 1, "the lower-bound variant (k > 2 passes) uses the plain 32-bit key path");
    constexpr int NW = NT / 32;
    constexpr bool XF = pm_transformed(PM);
    static_assert(NT == TT, "one thread per staged train row");
    __shared__ __align__(128) uint4 s_t[2][TT * 2];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float2 s_xy[MASK == 2 ? 2 : 1][MASK == 2 ? TT : 1];
    __shared__ uint32_t s_col[CROSS ? 2 : 1][CROSS ? NW : 1][CROSS ? TT : 1];
    __shared__ int s_cnt[FIN_RPT][NW];
    __shared__ int s_flag;

    if ((int)blockIdx.x < p.n_feed) {   // the first CTAs of the grid feed the others (SM-fed upload)
        if (!p.feed_stall) feed_rows<NT>(p);
        return;
    }
    const Segment sg = p.segs[blockIdx.x - p.n_feed];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    int col0 = 0;
    if (CROSS) col0 = p.problems[sg.problem].col0;

    // -- input gate, SM-fed variant: wait until every feeder has finished the round that covers our rows
    if (p.n_feed > 0) {
        if (warp == 0) {
            const int q_need = sg.q_row0 + sg.q_valid, t_need = sg.t_row0 + sg.t_count;
            auto round_of = [](int rows, int per_round) {   // rounds that must be complete for rows [0, rows)
                if (rows <= 0) return 0;
                if (rows <= per_round) return (rows + per_round / FEED_HEAD - 1) / (per_round / FEED_HEAD);
                return FEED_HEAD - 1 + (rows + per_round - 1) / per_round;
            };
            const uint32_t need = (p.feed_epoch << 16) |
                                  (uint32_t)min(p.feed_rounds, max(round_of(q_need, p.feed_q_rows), round_of(t_need, p.feed_t_rows)));
            const unsigned long long t0 = global_timer_ns();
            int ok = 1;
            uint32_t sleep_ns = 250u;
            while (true) {
                uint32_t v = 0xFFFFFFFFu;
                if (lane < p.n_feed) v = *(volatile const uint32_t *)(p.feed_prog + lane);
                v = __reduce_min_sync(0xffffffffu, v);
                if (v >= need) break;
                __nanosleep(sleep_ns);                      // back off: a thousand CTAs poll one cache line
                sleep_ns = min(sleep_ns * 2u, 4000u);
                if (global_timer_ns() - t0 > 4000000000ull) {
                    ok = 0;
                    if (lane == 0) *(volatile uint32_t *)p.status = 1u   /* pinned host word: a plain store, no PCIe atomic needed */;
                    break;
                }
            }
            // feeders: data stores, __threadfence, progress store; here: progress load, fence, data loads
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            if (lane == 0) s_flag = ok;
        }
        __syncthreads();
        if (!s_flag) return;
        __syncthreads();   // s_flag is reused by the kernel tail
    }

    // -- input gate (pipelined host path): wait until the copy engine has landed this segment's rows
    if (p.ready != nullptr) {
        if (tid == 0) {
            const unsigned long long need_q = p.ready_base + (unsigned long long)(sg.q_row0 + sg.q_valid);
            const unsigned long long need_t = p.ready_base + (unsigned long long)(sg.t_row0 + sg.t_count);
            const unsigned long long t0 = global_timer_ns();
            int ok = 1;
            while (ld_relaxed_sys(p.ready) < need_q || ld_relaxed_sys(p.ready + 1) < need_t) {
                __nanosleep(200);
                if (global_timer_ns() - t0 > 4000000000ull) {   // 4 s: the copies were never queued
                    ok = 0;
                    *(volatile uint32_t *)p.status = 1u   /* pinned host word: a plain store, no PCIe atomic needed */;
                    break;
                }
            }
            // the rows were written by the copy engine before the flag: order our reads (generic and
            // async proxy) after the flag read
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
            s_flag = ok;
        }
        __syncthreads();
        if (!s_flag) return;
        __syncthreads();   // s_flag is reused by the kernel tail
    }

    // train rows of this segment; with a device-side limit (a train set whose size was decided by an
    // earlier kernel on the same stream, e.g. the visible local-map points) the range is clamped here
    // (single problem).  The rows that exist are then re-cut evenly over this query block's `limit_segs` work
    // items, so every CTA gets the same share whatever the host guessed when it planned.
    int t_count = sg.t_count, t_row0 = sg.t_row0, t_local0 = sg.t_local0;
    if (p.t_limit != nullptr) {
        const int rows = max(0, __ldg(p.t_limit));
        const int s_idx = ((int)blockIdx.x - p.n_feed) % p.limit_segs;
        const int per = (rows + p.limit_segs - 1) / p.limit_segs;
        t_row0 = sg.t_row0 - sg.t_local0 + s_idx * per;
        t_local0 = s_idx * per;
        t_count = max(0, min(per, rows - t_local0));
    }

    // -- start the train stream first: the TMA of chunks 0 and 1 flies while the queries are loaded ----
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
        const int n0 = min(TT, t_count), n1 = min(TT, t_count - TT);
        if (n0 > 0) {
            mbar_expect_tx(&s_bar[0], (uint32_t)n0 * 32u);
            bulk_g2s(&s_t[0][0], p.t + 2 * (size_t)t_row0, (uint32_t)n0 * 32u, &s_bar[0]);
        }
        if (n1 > 0) {
            mbar_expect_tx(&s_bar[1], (uint32_t)n1 * 32u);
            bulk_g2s(&s_t[1][0], p.t + 2 * (size_t)(t_row0 + TT), (uint32_t)n1 * 32u, &s_bar[1]);
        }
    }

    // -- this thread's R query descriptors (two coalesced 16-byte loads each) ---------------------
    uint32_t qw[R][8];
    uint32_t ibias[R];          // CROSS: low bits of the column key (query index), dead bit if row absent
    bool valid[R];
    float qx[R], qy[R];
    const uint8_t *mrow[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int lr = r * NT + tid;
        valid[r] = lr < sg.q_valid;
        const int row = sg.q_row0 + (valid[r] ? lr : 0);
        // .cg (L2) loads: on the SM-fed host path these arrays are written by feeder CTAs of this very launch, and
        // ld.global.nc is only defined for memory that is read-only for the kernel's lifetime
        const uint4 a = __ldcg(p.q + 2 * (size_t)row);
        const uint4 b = __ldcg(p.q + 2 * (size_t)row + 1);
        qw[r][0] = a.x; qw[r][1] = a.y; qw[r][2] = a.z; qw[r][3] = a.w;
        qw[r][4] = b.x; qw[r][5] = b.y; qw[r][6] = b.z; qw[r][7] = b.w;
        if (XF) transform_desc(qw[r]);
        ibias[r] = (uint32_t)(sg.q_local0 + lr);
        if (MASK == 0 && !valid[r]) ibias[r] = KEY_DEAD;
        if (MASK == 2) {
            const float2 xy = __ldcg(p.q_xy + row);
            // an absent row gets NaN coordinates: every window compare is false
            qx[r] = valid[r] ? xy.x : __int_as_float(0x7fc00000);
            qy[r] = xy.y;
        }
        if (MASK == 1) mrow[r] = p.mask + (size_t)(sg.q_local0 + (valid[r] ? lr : 0)) * (size_t)p.mask_stride;
    }

    uint32_t b1[R], b2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { b1[r] = KEY_NONE; b2[r] = KEY_NONE; }
    uint32_t lb[R];
#pragma unroll
    for (int r = 0; r < R; ++r) lb[r] = (BOUND && valid[r]) ? __ldg(p.lower + sg.out_row0 + r * NT + tid) : 0u;

    const int nchunks = (t_count + TT - 1) / TT;
    auto chunk_rows = [&](int c) { return min(TT, t_count - c * TT); };
    auto fetch = [&](int c) {   // one thread: TMA bulk copy of chunk c into buffer c&1
        const uint32_t bytes = (uint32_t)chunk_rows(c) * 32u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes (transform) before async writes
        mbar_expect_tx(&s_bar[c & 1], bytes);
        bulk_g2s(&s_t[c & 1][0], p.t + 2 * (size_t)(t_row0 + c * TT), bytes, &s_bar[c & 1]);
    };
    auto stage_xy = [&](int c) {
        if (MASK == 2 && tid < chunk_rows(c)) s_xy[c & 1][tid] = __ldcg(p.t_xy + t_row0 + c * TT + tid);
    };
    auto land = [&](int c) {    // wait for chunk c, then (XF) rewrite its rows in place, one per thread
        mbar_wait(&s_bar[c & 1], (uint32_t)((c >> 1) & 1));
        if (XF && tid < chunk_rows(c)) {
            const uint4 a = s_t[c & 1][2 * tid], b = s_t[c & 1][2 * tid + 1];
            uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            transform_desc(w);
            s_t[c & 1][2 * tid] = make_uint4(w[0], w[1], w[2], w[3]);
            s_t[c & 1][2 * tid + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };

    __syncthreads();   // barrier init (thread 0, above) visible to every waiter
    if (nchunks > 0) {
        stage_xy(0);
        if (nchunks > 1) stage_xy(1);
        land(0);
    }
    __syncthreads();

    for (int c = 0; c < nchunks; ++c) {
        const int b = c & 1;
        const int n = chunk_rows(c);
        const uint32_t jbase = (uint32_t)(t_local0 + c * TT);
        if constexpr (R >= 2) {
            // ---- packed path: two queries share one register of chunk-local 16-bit keys ----------
            // key16 = d << 7 | j (j < 128), so one VIMNMX.U16x2 updates two queries at once; the
            // chunk's winners are folded into the 32-bit global keys once per 128 train rows.
            uint32_t p1[R / 2], p2[R / 2];
#pragma unroll
            for (int h = 0; h < R / 2; ++h) { p1[h] = 0xFFFFFFFFu; p2[h] = 0xFFFFFFFFu; }
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const uint4 ta = s_t[b][2 * j];
                const uint4 tb = s_t[b][2 * j + 1];
                const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
                float2 txy;
                if (MASK == 2) txy = s_xy[b][j];
                const uint32_t jpack = (uint32_t)j * 0x10001u;
                uint32_t ck = KEY_NONE;
#pragma unroll
                for (int h = 0; h < R / 2; ++h) {
                    uint32_t d[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int r = 2 * h + e;
                        d[e] = hamming256<PM>(qw[r], tw, p);
                        if (MASK == 1) {
                            const bool ok = valid[r] && (__ldg(mrow[r] + jbase + j) != 0);
                            d[e] = ok ? d[e] : DIST_MASKED;
                        }
                        if (MASK == 2) {
                            const bool ok = (fabsf(qx[r] - txy.x) < p.radius) && (fabsf(qy[r] - txy.y) < p.radius);
                            d[e] = ok ? d[e] : DIST_MASKED;
                        }
                        if (CROSS) ck = min(ck, d[e] * p.mul_d32 + ibias[r]);
                    }
                    const uint32_t packed = d[1] * p.mul_hi16 + (d[0] * p.mul_lo16 + jpack);
                    if (K == 2) p2[h] = min_u16x2(p2[h], max_u16x2(p1[h], packed));
                    p1[h] = min_u16x2(p1[h], packed);
                }
                if (CROSS) {
                    ck = __reduce_min_sync(0xffffffffu, ck);
                    if (lane == 0) s_col[b][warp][j] = ck;
                }
            }
            // fold the chunk winners into the global 32-bit keys (sorted-pair merge)
#pragma unroll
            for (int h = 0; h < R / 2; ++h) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = 2 * h + e;
                    const uint32_t k1 = e ? (p1[h] >> 16) : (p1[h] & 0xFFFFu);
                    const uint32_t k2 = e ? (p2[h] >> 16) : (p2[h] & 0xFFFFu);
                    const uint32_t c1 = k1 >= KEY16_DEAD ? KEY_NONE : ((k1 >> 7) << DIST_SHIFT) + jbase + (k1 & 127u);
                    const uint32_t c2 = (K == 1 || k2 >= KEY16_DEAD) ? KEY_NONE : ((k2 >> 7) << DIST_SHIFT) + jbase + (k2 & 127u);
                    if (K == 2) b2[r] = min(max(b1[r], c1), min(b2[r], c2));
                    b1[r] = min(b1[r], c1);
                }
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < n; ++j) {
                const uint4 ta = s_t[b][2 * j];
                const uint4 tb = s_t[b][2 * j + 1];
                const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
                float2 txy;
                if (MASK == 2) txy = s_xy[b][j];
                const uint32_t jj = jbase + (uint32_t)j;
                uint32_t ck = KEY_NONE;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    uint32_t d = hamming256<PM>(qw[r], tw, p);
                    if (MASK == 1) {
                        const bool ok = valid[r] && (__ldg(mrow[r] + jj) != 0);
                        d = ok ? d : DIST_MASKED;
                    }
                    if (MASK == 2) {
                        const bool ok = (fabsf(qx[r] - txy.x) < p.radius) && (fabsf(qy[r] - txy.y) < p.radius);
                        d = ok ? d : DIST_MASKED;
                    }
                    uint32_t key = d * p.mul_d32 + jj;
                    if (BOUND) key = key > lb[r] ? key : KEY_NONE;
                    if (K == 2) b2[r] = min(b2[r], max(b1[r], key));
                    b1[r] = min(b1[r], key);
                    if (CROSS) ck = min(ck, d * p.mul_d32 + ibias[r]);
                }
                if (CROSS) {
                    ck = __reduce_min_sync(0xffffffffu, ck);
                    if (lane == 0) s_col[b][warp][j] = ck;
                }
            }
        }
        if (c + 1 < nchunks) land(c + 1);
        __syncthreads();   // chunk c fully consumed: s_t[b] / s_xy[b] free, s_col[b] complete, chunk c+1 ready
        if (c + 2 < nchunks) {
            if (tid == 0) fetch(c + 2);
            stage_xy(c + 2);
        }
        if (CROSS) {
            // s_col[b] is next written in iteration c+2, i.e. after the barrier of iteration c+1
            for (int j = tid; j < n; j += NT) {
                uint32_t m = s_col[b][0][j];
#pragma unroll
                for (int w = 1; w < NW; ++w) m = min(m, s_col[b][w][j]);
                if (m < KEY_DEAD) atomicMin(p.colkeys + (size_t)col0 + t_local0 + c * TT + j, m);
            }
        }
    }

    // -- commit: associative min-merge into the global row state ---------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!valid[r] || b1[r] >= KEY_DEAD) continue;
        // row state = (best << 32) | second, updated through its two 32-bit halves (little endian):
        // one atomicMin on `best` returns the displaced key; whatever lost there, or this segment's
        // own runner-up, competes for `second` with a fire-and-forget atomic.  Every key except the
        // final best is offered to `second` exactly when it stops being (or fails to become) the
        // best, so `second` ends as the true runner-up for any arrival order.
        uint32_t *half = reinterpret_cast<uint32_t *>(p.rowstate + (size_t)(sg.out_row0 + r * NT + tid));
        if (K == 1) {
            atomicMin(half + 1, b1[r]);
        } else {
            const uint32_t n2 = b2[r] >= KEY_DEAD ? KEY_NONE : b2[r];
            const uint32_t displaced = atomicMin(half + 1, b1[r]);
            const uint32_t cand = min(max(displaced, b1[r]), n2);
            if (cand != KEY_NONE) atomicMin(half, cand);
        }
    }

    // -- problem completion: the CTA whose segment is the last of its problem finalizes it -------------
    // (threadfence + counter: every CTA's state updates are visible before its count is)
    if (p.defer_finalize) return;   // one very large problem: finalized by the tile-parallel kernels below
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t n_segs = (uint32_t)p.problems[sg.problem].n_segs;
        const uint32_t old = atomicAdd(p.done + sg.problem, 1u);   // idles at 0xFFFFFFFF: first arrival wraps to 0
        s_flag = (old == n_segs - 2u) ? 1 : 0;
        if (s_flag) p.done[sg.problem] = 0xFFFFFFFFu;              // self-cleaning, like the rest of the workspace
    }
    __syncthreads();
    if (s_flag) {
        __threadfence();
        finalize_problem_v1<NT>(p, sg.problem, s_cnt);
    }
}


// The n train rows staged in buffer b (train indices jbase ..) against the R queries of every thread.  A macro, not a
// lambda: expanded in place the loop gets the instruction schedule of the round-1 kernel (1023 us on the headline batch);
// as an inlined lambda ptxas interleaves the same 194 instructions differently and the launch takes 1039 us
// (tools/ab_old_new.py).
#define BFM_SCAN_CHUNK(b_arg, n_arg, jbase_arg)                                                                      \
    {                                                                                                                \
        const int b = (b_arg);                                                                                       \
        const int n = (n_arg);                                                                                       \
        const uint32_t jbase = (jbase_arg);                                                                          \
                                                                                                                     \
        if constexpr (R >= 2) {                                                                                      \
            /* ---- packed path: two queries share one register of chunk-local 16-bit keys ---------- */             \
            /* key16 = d << 7 | j (j < 128), so one VIMNMX.U16x2 updates two queries at once; the */                 \
            /* chunk's winners are folded into the 32-bit global keys once per 128 train rows. */                    \
            uint32_t p1[R / 2], p2[R / 2];                                                                           \
_Pragma("unroll")                                                                                                    \
            for (int h = 0; h < R / 2; ++h) { p1[h] = 0xFFFFFFFFu; p2[h] = 0xFFFFFFFFu; }                            \
_Pragma("unroll 2")                                                                                                  \
            for (int j = 0; j < n; ++j) {                                                                            \
                const uint4 ta = s_t[b][2 * j];                                                                      \
                const uint4 tb = s_t[b][2 * j + 1];                                                                  \
                const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};                             \
                float2 txy;                                                                                          \
                if (MASK == 2) txy = s_xy[b][j];                                                                     \
                const uint32_t jpack = (uint32_t)j * 0x10001u;                                                       \
                uint32_t ck = KEY_NONE;                                                                              \
_Pragma("unroll")                                                                                                    \
                for (int h = 0; h < R / 2; ++h) {                                                                    \
                    uint32_t d[2];                                                                                   \
_Pragma("unroll")                                                                                                    \
                    for (int e = 0; e < 2; ++e) {                                                                    \
                        const int r = 2 * h + e;                                                                     \
                        d[e] = hamming256<PM>(qw[r], tw, p);                                                         \
                        if (MASK == 1) {                                                                             \
                            const bool ok = valid[r] && (__ldg(mrow[r] + jbase + j) != 0);                           \
                            d[e] = ok ? d[e] : DIST_MASKED;                                                          \
                        }                                                                                            \
                        if (MASK == 2) {                                                                             \
                            const bool ok = (fabsf(qx[r] - txy.x) < p.radius) && (fabsf(qy[r] - txy.y) < p.radius);  \
                            d[e] = ok ? d[e] : DIST_MASKED;                                                          \
                        }                                                                                            \
                        if (CROSS) ck = min(ck, d[e] * p.mul_d32 + ibias[r]);                                        \
                    }                                                                                                \
                    const uint32_t packed = d[1] * p.mul_hi16 + (d[0] * p.mul_lo16 + jpack);                         \
                    if (K == 2) p2[h] = min_u16x2(p2[h], max_u16x2(p1[h], packed));                                  \
                    p1[h] = min_u16x2(p1[h], packed);                                                                \
                }                                                                                                    \
                if (CROSS) {                                                                                         \
                    ck = __reduce_min_sync(0xffffffffu, ck);                                                         \
                    if (lane == 0) s_col[b][warp][j] = ck;                                                           \
                }                                                                                                    \
            }                                                                                                        \
            /* fold the chunk winners into the global 32-bit keys (sorted-pair merge) */                             \
_Pragma("unroll")                                                                                                    \
            for (int h = 0; h < R / 2; ++h) {                                                                        \
_Pragma("unroll")                                                                                                    \
                for (int e = 0; e < 2; ++e) {                                                                        \
                    const int r = 2 * h + e;                                                                         \
                    const uint32_t k1 = e ? (p1[h] >> 16) : (p1[h] & 0xFFFFu);                                       \
                    const uint32_t k2 = e ? (p2[h] >> 16) : (p2[h] & 0xFFFFu);                                       \
                    const uint32_t c1 = k1 >= KEY16_DEAD ? KEY_NONE : ((k1 >> 7) << DIST_SHIFT) + jbase + (k1 & 127u); \
                    const uint32_t c2 = (K == 1 || k2 >= KEY16_DEAD) ? KEY_NONE : ((k2 >> 7) << DIST_SHIFT) + jbase + (k2 & 127u); \
                    if (K == 2) b2[r] = min(max(b1[r], c1), min(b2[r], c2));                                         \
                    b1[r] = min(b1[r], c1);                                                                          \
                }                                                                                                    \
            }                                                                                                        \
        } else {                                                                                                     \
_Pragma("unroll 2")                                                                                                  \
            for (int j = 0; j < n; ++j) {                                                                            \
                const uint4 ta = s_t[b][2 * j];                                                                      \
                const uint4 tb = s_t[b][2 * j + 1];                                                                  \
                const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};                             \
                float2 txy;                                                                                          \
                if (MASK == 2) txy = s_xy[b][j];                                                                     \
                const uint32_t jj = jbase + (uint32_t)j;                                                             \
                uint32_t ck = KEY_NONE;                                                                              \
_Pragma("unroll")                                                                                                    \
                for (int r = 0; r < R; ++r) {                                                                        \
                    uint32_t d = hamming256<PM>(qw[r], tw, p);                                                       \
                    if (MASK == 1) {                                                                                 \
                        const bool ok = valid[r] && (__ldg(mrow[r] + jj) != 0);                                      \
                        d = ok ? d : DIST_MASKED;                                                                    \
                    }                                                                                                \
                    if (MASK == 2) {                                                                                 \
                        const bool ok = (fabsf(qx[r] - txy.x) < p.radius) && (fabsf(qy[r] - txy.y) < p.radius);      \
                        d = ok ? d : DIST_MASKED;                                                                    \
                    }                                                                                                \
                    uint32_t key = d * p.mul_d32 + jj;                                                               \
                    if (BOUND) key = key > lb[r] ? key : KEY_NONE;                                                   \
                    if (K == 2) b2[r] = min(b2[r], max(b1[r], key));                                                 \
                    b1[r] = min(b1[r], key);                                                                         \
                    if (CROSS) ck = min(ck, d * p.mul_d32 + ibias[r]);                                               \
                }                                                                                                    \
                if (CROSS) {                                                                                         \
                    ck = __reduce_min_sync(0xffffffffu, ck);                                                         \
                    if (lane == 0) s_col[b][warp][j] = ck;                                                           \
                }                                                                                                    \
            }                                                                                                        \
        }                                                                                                            \
                                                                                                                     \
    }

template <int R, int K, bool CROSS, int MASK, int PM, int NT, bool BOUND>
__global__ void __launch_bounds__(NT, MASK == 0 ? 8 : 6) bfm_scan_persistent_kernel(const __grid_constant__ ScanParams p) {
    constexpr bool DYN = true;
    static_assert(!BOUND || R == 1, "the lower-bound variant (k > 2 passes) uses the plain 32-bit key path");
    constexpr int NW = NT / 32;
    constexpr bool XF = pm_transformed(PM);
    static_assert(NT == TT, "one thread per staged train row");
    __shared__ __align__(128) uint4 s_t[2][TT * 2];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ float2 s_xy[MASK == 2 ? 2 : 1][MASK == 2 ? TT : 1];
    __shared__ uint32_t s_col[CROSS ? 2 : 1][CROSS ? NW : 1][CROSS ? TT : 1];
    __shared__ int s_cnt[FIN_RPT][NW];
    __shared__ int s_flag;
    __shared__ int s_run[DYN ? 8 : 1];   // persistent form: CTA-uniform state (see below)

    trace_mark(p, 0);   // CTA entry
#ifndef BFM_AB_NOTRACE
    if (p.trace != nullptr && threadIdx.x == 0 && (int)blockIdx.x + p.trace_base < p.trace_cap) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[(size_t)(p.trace_base + (int)blockIdx.x) * 8 + 7] = smid;
    }
#endif
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int cta = (int)blockIdx.x;

    // the query block held in registers and its running keys
    uint32_t qw[R][8];
    uint32_t ibias[R];          // CROSS: low bits of the column key (query index), dead bit if row absent
#ifdef BFM_AB_VALID
    bool valid[R];
#else
    bool valid[MASK == 1 ? R : 1];
#endif
    float qx[R], qy[R];
    const uint8_t *mrow[R];
    uint32_t b1[R], b2[R], lb[R];
    int col0 = 0;

    // this thread's R query descriptors of a block (two coalesced 16-byte loads each); the running keys start empty
    auto load_queries = [&](const int q_row0, const int q_valid, const int q_local0, const int out_row0, const int problem) {
        if (CROSS) col0 = p.problems[problem].col0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int lr = r * NT + tid;
            const bool have = lr < q_valid;
#ifdef BFM_AB_VALID
            valid[r] = have;
#else
            if (MASK == 1) valid[r] = have;
#endif
            const int row = q_row0 + (have ? lr : 0);
            // .cg (L2) loads: on the SM-fed host path these arrays are written by feeder CTAs of this very launch, and
            // ld.global.nc is only defined for memory that is read-only for the kernel's lifetime
            const uint4 a = __ldcg(p.q + 2 * (size_t)row);
            const uint4 b = __ldcg(p.q + 2 * (size_t)row + 1);
            qw[r][0] = a.x; qw[r][1] = a.y; qw[r][2] = a.z; qw[r][3] = a.w;
            qw[r][4] = b.x; qw[r][5] = b.y; qw[r][6] = b.z; qw[r][7] = b.w;
            if (XF) transform_desc(qw[r]);
            ibias[r] = (uint32_t)(q_local0 + lr);
            if (MASK == 0 && !have) ibias[r] = KEY_DEAD;
            if (MASK == 2) {
                const float2 xy = __ldcg(p.q_xy + row);
                // an absent row gets NaN coordinates: every window compare is false
                qx[r] = have ? xy.x : __int_as_float(0x7fc00000);
                qy[r] = xy.y;
            }
            if (MASK == 1) mrow[r] = p.mask + (size_t)(q_local0 + (have ? lr : 0)) * (size_t)p.mask_stride;
            b1[r] = KEY_NONE;
            b2[r] = KEY_NONE;
            lb[r] = (BOUND && have) ? __ldg(p.lower + out_row0 + r * NT + tid) : 0u;
        }
    };

    // commit: associative min-merge of the running keys into the global row state
    // row state = (best << 32) | second, updated through its two 32-bit halves (little endian):
    // one atomicMin on `best` returns the displaced key; whatever lost there, or this run's
    // own runner-up, competes for `second` with a fire-and-forget atomic.  Every key except the
    // final best is offered to `second` exactly when it stops being (or fails to become) the
    // best, so `second` ends as the true runner-up for any arrival order.
    auto commit = [&](const int out0, const int q_valid) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
#ifdef BFM_AB_VALID
            (void)q_valid;
            if (!valid[r] || b1[r] >= KEY_DEAD) continue;
#else
            if (r * NT + tid >= q_valid || b1[r] >= KEY_DEAD) continue;
#endif
            uint32_t *half = reinterpret_cast<uint32_t *>(p.rowstate + (size_t)(out0 + r * NT + tid));
            if (K == 1) {
                atomicMin(half + 1, b1[r]);
            } else {
                const uint32_t n2 = b2[r] >= KEY_DEAD ? KEY_NONE : b2[r];
                const uint32_t displaced = atomicMin(half + 1, b1[r]);
                const uint32_t cand = min(max(displaced, b1[r]), n2);
                if (cand != KEY_NONE) atomicMin(half, cand);
            }
        }
    };

    // cross-check: the per-warp column minima of a scanned chunk (complete after a barrier) into the global column keys
    auto flush_cols = [&](const int b, const int n, const int col_first) {
        for (int j = tid; j < n; j += NT) {
            uint32_t m = s_col[b][0][j];
#pragma unroll
            for (int w = 1; w < NW; ++w) m = min(m, s_col[b][w][j]);
            if (m < KEY_DEAD) atomicMin(p.colkeys + (size_t)col0 + col_first + j, m);
        }
    };
    // a landed chunk is rewritten in place (XF), one row per thread
    auto transform_rows = [&](const int b, const int n) {
        if (XF && tid < n) {
            const uint4 a = s_t[b][2 * tid], bb = s_t[b][2 * tid + 1];
            uint32_t w[8] = {a.x, a.y, a.z, a.w, bb.x, bb.y, bb.z, bb.w};
            transform_desc(w);
            s_t[b][2 * tid] = make_uint4(w[0], w[1], w[2], w[3]);
            s_t[b][2 * tid + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }

    {
        // CTA-uniform state, kept in shared memory so that it costs the inner loop no registers: s_run[0] current item,
        // [1] problem / [2] first query row / [3] first output row / [5] valid rows of the query block held in
        // registers, [4] items scanned for it since the last commit, [6] train chunks streamed so far
        uint32_t nxt = 0;   // (thread 0) the next ticket, in flight while the current item is scanned
        if (tid == 0) {
            s_run[0] = cta;
            s_run[1] = -1;
            s_run[2] = -1;
            s_run[3] = 0;
            s_run[4] = 0;
            s_run[5] = 0;
            s_run[6] = 0;
            if (cta == 0) *p.queue_other = 0u;   // the counter of the NEXT launch (idle: the previous launch has drained)
            nxt = (uint32_t)p.n_ctas + atomicAdd(p.queue, 1u);
        }
        __syncthreads();
        for (;;) {
            // (every value of s_run read here was written before the last barrier)
            const int item = s_run[0];
            if (item >= p.n_items) break;
            const bool first_item = item == cta;
            const uint32_t g0 = (uint32_t)s_run[6];   // chunk g lives in buffer g & 1, barrier phase (g >> 1) & 1
            const int acc_problem = s_run[1], acc_q0 = s_run[2], acc_out0 = s_run[3], acc_items = s_run[4], acc_qvalid = s_run[5];
            const Segment sg = p.segs[item];
            int t_count = sg.t_count, t_row0 = sg.t_row0, t_local0 = sg.t_local0;
            if (p.t_limit != nullptr) {   // a device-side train count: the rows that exist, re-cut over the planned items
                const int rows = max(0, __ldg(p.t_limit));
                const int s_idx = item % p.limit_segs;
                const int per = (rows + p.limit_segs - 1) / p.limit_segs;
                t_row0 = sg.t_row0 - sg.t_local0 + s_idx * per;
                t_local0 = s_idx * per;
                t_count = max(0, min(per, rows - t_local0));
            }
            // start the train stream first (the closing barrier of the previous item freed both buffers)
            if (tid == 0) {
                const int n0 = min(TT, t_count), n1 = min(TT, t_count - TT);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes of the previous item (transform) before async writes
                if (n0 > 0) {
                    mbar_expect_tx(&s_bar[g0 & 1], (uint32_t)n0 * 32u);
                    bulk_g2s(&s_t[g0 & 1][0], p.t + 2 * (size_t)t_row0, (uint32_t)n0 * 32u, &s_bar[g0 & 1]);
                }
                if (n1 > 0) {
                    mbar_expect_tx(&s_bar[(g0 + 1) & 1], (uint32_t)n1 * 32u);
                    bulk_g2s(&s_t[(g0 + 1) & 1][0], p.t + 2 * (size_t)(t_row0 + TT), (uint32_t)n1 * 32u, &s_bar[(g0 + 1) & 1]);
                }
            }
            // a new query block: commit the run that ends here, then load the block
            const bool new_block = sg.problem != acc_problem || sg.q_local0 != acc_q0;
            if (new_block && acc_items > 0) commit(acc_out0, acc_qvalid);
            __syncthreads();   // s_run has been read by every thread; release: the atomics above are visible to whoever acquires the counter
            if (tid == 0) {
                if (new_block) {
                    if (acc_items > 0) red_release_gpu_add(p.done + acc_problem, (uint32_t)acc_items);
                    s_run[1] = sg.problem;
                    s_run[2] = sg.q_local0;
                    s_run[3] = sg.out_row0;
                    s_run[5] = sg.q_valid;
                }
                s_run[4] = new_block ? 1 : acc_items + 1;
            }
            if (new_block) load_queries(sg.q_row0, sg.q_valid, sg.q_local0, sg.out_row0, sg.problem);

            const int nchunks = (t_count + TT - 1) / TT;
            auto chunk_rows = [&](int c) { return min(TT, t_count - c * TT); };
            auto fetch = [&](int c) {   // one thread: TMA bulk copy of chunk c into buffer (g0 + c) & 1
                const uint32_t bytes = (uint32_t)chunk_rows(c) * 32u;
                const uint32_t gb = (g0 + (uint32_t)c) & 1u;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes (transform) before async writes
                mbar_expect_tx(&s_bar[gb], bytes);
                bulk_g2s(&s_t[gb][0], p.t + 2 * (size_t)(t_row0 + c * TT), bytes, &s_bar[gb]);
            };
            auto stage_xy = [&](int c) {
                if (MASK == 2 && tid < chunk_rows(c)) s_xy[(g0 + (uint32_t)c) & 1u][tid] = __ldcg(p.t_xy + t_row0 + c * TT + tid);
            };
            auto land = [&](int c) {    // wait for chunk c, then (XF) rewrite its rows in place, one per thread
                const uint32_t g = g0 + (uint32_t)c;
                mbar_wait(&s_bar[g & 1u], (g >> 1) & 1u);
                transform_rows((int)(g & 1u), chunk_rows(c));
            };
            if (first_item) trace_mark(p, 1);   // TMA issued, queries loaded
            if (nchunks > 0) {
                stage_xy(0);
                if (nchunks > 1) stage_xy(1);
                land(0);
            }
            __syncthreads();
            if (first_item) trace_mark(p, 2);   // first chunk landed
            for (int c = 0; c < nchunks; ++c) {
                const int buf = (int)((g0 + (uint32_t)c) & 1u);
                const int rows = chunk_rows(c);
                BFM_SCAN_CHUNK(buf, rows, (uint32_t)(t_local0 + c * TT));
                if (c + 1 < nchunks) land(c + 1);
                __syncthreads();   // chunk c fully consumed: s_t[b] / s_xy[b] free, s_col[b] complete, chunk c+1 ready
                if (c + 2 < nchunks) {
                    if (tid == 0) fetch(c + 2);
                    stage_xy(c + 2);
                }
                if (CROSS) flush_cols(buf, rows, t_local0 + c * TT);   // s_col[b] is next written two chunks later, after the next barrier
            }
            if (first_item) trace_mark(p, 3);   // scan of the first item done
            // the next ticket (drawn while this item was scanned); draw the one after it right away
            if (tid == 0) {
                s_run[6] = (int)(g0 + (uint32_t)nchunks);
                s_run[0] = (int)min(nxt, 0x7FFFFFFFu);
                if (nxt < (uint32_t)p.n_items) nxt = (uint32_t)p.n_ctas + atomicAdd(p.queue, 1u);
            }
            __syncthreads();
        }
        // -- the queue is empty: commit the last run, then the finalize tiles the plan gave this CTA -----------------
        if (s_run[4] > 0) {
            commit(s_run[3], s_run[5]);
            __syncthreads();
            if (tid == 0) red_release_gpu_add(p.done + s_run[1], (uint32_t)s_run[4]);
        }
        trace_mark(p, 5);   // all work items of this CTA committed
        const int2 w = __ldg(p.cta_tiles + cta);
        if (w.y > 0) {
            for (int f = w.x; f < w.x + w.y; ++f) finalize_tile<NT>(p, p.fin_tiles + f, s_cnt, &s_flag);
            trace_mark(p, 6);   // its finalize tiles done
        }
    }
}

#ifndef BFM_SCAN_INST_ONLY   // defined once, in bfm_api.cu
// Static form with a few large problems (a brute-force sweep point): a second launch finalizes them tile by tile, one
// tile per CTA, instead of one CTA walking 65536 rows (64 tiles x ~5 us).  The scan has completed (stream order), so
// the tiles' wait on the done counter is satisfied at once (the table says 0 work items).
__global__ void __launch_bounds__(128) bfm_tiles_kernel(const __grid_constant__ ScanParams p) {
    __shared__ int s_cnt[FIN_RPT][4];
    __shared__ int s_flag;
    finalize_tile<128>(p, p.fin_tiles + blockIdx.x, s_cnt, &s_flag);
}
#endif

typedef void (*ScanFn)(const ScanParams);
// defined in bfm_scan_inst.cu (compiled once per register tile R and mode: 0 k = 1, 1 cross-check, 2 k = 2);
// dyn: the persistent form (resident inputs) instead of the static one
ScanFn pick_scan_r1_m0(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r1_m1(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r1_m2(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r2_m0(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r2_m1(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r2_m2(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r4_m0(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r4_m1(int mask, int pm, bool bound, bool dyn);
ScanFn pick_scan_r4_m2(int mask, int pm, bool bound, bool dyn);

// defined in bfm_tensor.cu: the matching kernel on the tensor cores (tcgen05; resident inputs, k <= 2, no mask, no
// cross-check).  Work items are Segments with 256-row query blocks whose q_row0 / t_row0 are rows of the EXPANDED planes.
constexpr int TC_BQ = 256;               // query rows per work item
constexpr int TC_BT = 128;               // train rows per tile: segment lengths are multiples of it
constexpr int TC_THREADS = 608;          // threads of a CTA of the tensor scan (bfm_tensor.cuh)
constexpr int TC_SLACK_ROWS = 512;       // rows past the end of a plane that a tile may read (they only have to exist)
struct TensorLaunch {
    const void *q, *t;                   // packed descriptors (32 bytes per row)
    const Problem *probs;                // device table: pad = first expanded query row, col0 = first expanded train row
    int n_problems, max_rows;            // max_rows: the longest query / train set (grid of the expansion); 0 = planes are current
    void *xq, *xt;                       // expanded planes [2][rows + slack][128]
    unsigned long long xq_plane, xt_plane;   // bytes per plane
    const Segment *items;
    int n_items, grid;
    unsigned long long *rowstate;
    uint32_t *status;                    // pinned host word or NULL
    cudaEvent_t ev_scan[2];              // timing knob: recorded around the scan kernel alone (NULL otherwise)
};
int tensor_init();
int tensor_launch(const TensorLaunch &L, cudaStream_t st);

}  // namespace bfm
