// bfm_window.cuh - projection-window search over a spatially binned train set.
// (included by bfm_api.cu inside its anonymous namespace, after bfm_kernels.cuh)
//
// The tracking shape of the north star (a frame's descriptors against the local map's projected
// points, each query only allowed to match points within `radius` pixels of it, in both axes) admits
// ~0.3 % of the Q x T pairs at 2000 x 20000 / 15 px.  The brute-force kernel evaluates all of them
// and masks; this path evaluates only the admissible ones:
//   bin      counting sort of the train rows by grid cell (cell edge >= 2 * radius, 64 x 64 cells, indices
//            wrapped so ANY coordinate range maps somewhere: aliasing only adds candidates): one pass
//            ranks every row inside its cell with an atomic and its last CTA scans the counters, a
//            second pass moves the rows
//   search   one warp per query walks the <= 3 x 3 cells its window touches, applies the exact fp32
//            predicate of the brute-force kernel, XOR+POPC on admitted rows only, per-lane top-2 of the
//            packed keys (distance << 22 | ORIGINAL train index, so ties still break to the lowest
//            trainIdx), warp merge, cross-check column keys by atomicMin; the last CTA finalizes.
// Results are bit-identical to the brute-force window path (same predicate, same keys, same min).

constexpr int WB_G = 64;                 // grid is WB_G x WB_G cells (wrapped)
constexpr int WB_CELLS = WB_G * WB_G;
constexpr int WB_MAX_ROWS = 1 << 20;     // above this the brute-force window kernel is used

struct BinView {
    uint4 *desc;        // [T][2] train descriptors, cell order
    float2 *xy;         // [T]
    int32_t *orig;      // [T] original train index
    int32_t *cell_start;  // [WB_CELLS + 1]
};

// the same monotone map is used for train points and for window bounds, which is what guarantees
// that every admissible row lies in a visited cell (see the header comment of bfm_api.cu's caller)
__device__ __forceinline__ long long wb_cell_coord(float x, double inv_cell) {
    const double c = floor((double)x * inv_cell);
    return (long long)fmin(fmax(c, -4.0e15), 4.0e15);
}
__device__ __forceinline__ int wb_wrap(long long c) { return (int)(((c % WB_G) + WB_G) % WB_G); }

// pass 1 (many CTAs): cell of every train row, rank inside its cell (the atomic's return value);
// the last CTA to finish turns the counters into cell offsets and zeroes them for the next call.
// (a batch: blockIdx.y = problem; every problem has its own counters, ticket and cell table, and its rows sit at
// offset Problem::col0 - the sum of the train rows of the problems before it - in the binned arrays)
__global__ void __launch_bounds__(256) wb_count_kernel(const float2 *t_xy_all, const Problem *problems, const int32_t *t_limit, double inv_cell,
                                                       int32_t *cnt_all, uint32_t *ticket_all, int32_t *cell_of_all, int32_t *rank_of_all,
                                                       int32_t *cell_start_all) {
    __shared__ int s_warp[32];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const Problem pr = problems[blockIdx.y];
    const float2 *t_xy = t_xy_all + pr.t_begin;
    int32_t *cnt = cnt_all + (size_t)blockIdx.y * WB_CELLS, *cell_start = cell_start_all + (size_t)blockIdx.y * (WB_CELLS + 1);
    int32_t *cell_of = cell_of_all + pr.col0, *rank_of = rank_of_all + pr.col0;
    uint32_t *ticket = ticket_all + blockIdx.y;
    const int n_rows = pr.t_count;
    const int n = t_limit ? max(0, min(n_rows, *t_limit)) : n_rows;
    const int i = blockIdx.x * 256 + tid;
    if (i < n) {
        const float2 p = t_xy[i];
        const int cell = wb_wrap(wb_cell_coord(p.y, inv_cell)) * WB_G + wb_wrap(wb_cell_coord(p.x, inv_cell));
        cell_of[i] = cell;
        rank_of[i] = atomicAdd(cnt + cell, 1);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) *ticket = 0;
    // exclusive scan of the WB_CELLS counters by this CTA: 16 per thread
    constexpr int PER = WB_CELLS / 256;
    const int lane = tid & 31, warp = tid >> 5;
    int v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = __ldcg(cnt + tid * PER + k); sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < 8 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        if (lane < 8) s_warp[lane] = w;
    }
    __syncthreads();
    int run = inc - sum + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        cell_start[tid * PER + k] = run;
        cnt[tid * PER + k] = 0;      // self-cleaning
        run += v[k];
    }
    if (tid == 255) cell_start[WB_CELLS] = run;
}

// pass 2 (many CTAs): rows to their place in cell order
__global__ void __launch_bounds__(256) wb_scatter_kernel(const uint4 *t_desc_all, const float2 *t_xy_all, const Problem *problems,
                                                         const int32_t *t_limit, const int32_t *cell_of_all, const int32_t *rank_of_all,
                                                         BinView out_all) {
    const Problem pr = problems[blockIdx.y];
    const uint4 *t_desc = t_desc_all + 2 * (size_t)pr.t_begin;
    const float2 *t_xy = t_xy_all + pr.t_begin;
    const int32_t *cell_of = cell_of_all + pr.col0, *rank_of = rank_of_all + pr.col0;
    BinView out{out_all.desc + 2 * (size_t)pr.col0, out_all.xy + pr.col0, out_all.orig + pr.col0,
                out_all.cell_start + (size_t)blockIdx.y * (WB_CELLS + 1)};
    const int n_rows = pr.t_count;
    const int n = t_limit ? max(0, min(n_rows, *t_limit)) : n_rows;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int pos = out.cell_start[cell_of[i]] + rank_of[i];
    out.desc[2 * (size_t)pos] = t_desc[2 * (size_t)i];
    out.desc[2 * (size_t)pos + 1] = t_desc[2 * (size_t)i + 1];
    out.xy[pos] = t_xy[i];
    out.orig[pos] = i;
}

constexpr int WS_NT = 256;   // 8 queries per CTA, one warp each

template <int K, bool CROSS>
__global__ void __launch_bounds__(WS_NT) wb_search_kernel(const __grid_constant__ ScanParams p, const BinView bins_all,
                                                          const double inv_cell) {
    __shared__ int s_cnt[FIN_RPT][WS_NT / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pi = blockIdx.y;
    const Problem pr = p.problems[pi];
    const BinView bins{bins_all.desc + 2 * (size_t)pr.col0, bins_all.xy + pr.col0, bins_all.orig + pr.col0,
                       bins_all.cell_start + (size_t)pi * (WB_CELLS + 1)};
    const int n_query = pr.q_count;
    const int qi = blockIdx.x * (WS_NT / 32) + warp;
    if (qi < n_query) {
        const size_t qrow = (size_t)pr.q_begin + qi;
        const uint4 a = __ldg(p.q + 2 * qrow), b = __ldg(p.q + 2 * qrow + 1);
        const uint32_t qw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float2 qxy = __ldg(p.q_xy + qrow);
        uint32_t b1 = KEY_NONE, b2 = KEY_NONE;
        if (!(isnan(qxy.x) || isnan(qxy.y))) {
            const long long cx0 = wb_cell_coord(qxy.x - p.radius, inv_cell), cx1 = min(wb_cell_coord(qxy.x + p.radius, inv_cell), cx0 + 2);
            const long long cy0 = wb_cell_coord(qxy.y - p.radius, inv_cell), cy1 = min(wb_cell_coord(qxy.y + p.radius, inv_cell), cy0 + 2);
            for (long long cy = cy0; cy <= cy1; ++cy) {
                for (long long cx = cx0; cx <= cx1; ++cx) {
                    const int cell = wb_wrap(cy) * WB_G + wb_wrap(cx);
                    const int r0 = bins.cell_start[cell], r1 = bins.cell_start[cell + 1];
                    for (int r = r0 + lane; r < r1; r += 32) {
                        const float2 txy = bins.xy[r];
                        if (!((fabsf(qxy.x - txy.x) < p.radius) && (fabsf(qxy.y - txy.y) < p.radius))) continue;
                        const uint4 ta = bins.desc[2 * (size_t)r], tb = bins.desc[2 * (size_t)r + 1];
                        const uint32_t d = __popc(qw[0] ^ ta.x) + __popc(qw[1] ^ ta.y) + __popc(qw[2] ^ ta.z) + __popc(qw[3] ^ ta.w) +
                                           __popc(qw[4] ^ tb.x) + __popc(qw[5] ^ tb.y) + __popc(qw[6] ^ tb.z) + __popc(qw[7] ^ tb.w);
                        const uint32_t orig = (uint32_t)bins.orig[r];
                        const uint32_t key = (d << DIST_SHIFT) | orig;
                        if (K == 2) b2 = min(b2, max(b1, key));
                        b1 = min(b1, key);
                        if (CROSS) atomicMin(p.colkeys + (size_t)pr.col0 + orig, (d << DIST_SHIFT) | (uint32_t)qi);
                    }
                }
            }
        }
        // merge the 32 sorted pairs (a wrapped cell may have been visited twice: equal keys collapse)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, b1, o), o2 = __shfl_xor_sync(0xffffffffu, b2, o);
            const uint32_t m1 = min(b1, o1);
            uint32_t m2 = min(max(b1, o1), min(b2, o2));
            if (b1 == o1) m2 = min(b2, o2);   // the same row seen by both: not its own runner-up
            b1 = m1;
            b2 = m2;
        }
        if (lane == 0) p.rowstate[(size_t)pr.out_begin + qi] = ((unsigned long long)b1 << 32) | (K == 2 ? b2 : KEY_NONE);
    }
    // the last CTA of a problem finalizes it (every problem counts gridDim.x arrivals: CTAs past its queries just arrive)
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t old = atomicAdd(p.done + pi, 1u);
        s_flag = (old == (uint32_t)gridDim.x - 2u) ? 1 : 0;
        if (s_flag) p.done[pi] = 0xFFFFFFFFu;
    }
    __syncthreads();
    if (s_flag) {
        __threadfence();
        finalize_problem<WS_NT>(p, pi, s_cnt);
    }
}
