// bfm_api.cu - host side of libbfm_b200.so: the C ABI declared in include/bfm.h.
//
// Replaces the cv2.BFMatcher object held at reference slam/tracking.py:45 (and
// slam/local_mapping.py:21, slam/covisibility_graph.py:34) for the calls at
// slam/tracking.py:56,121.  Everything here is plumbing around two kernels
// (bfm_kernels.cuh): plan the (query block x train range) segments, reset the packed-key state,
// launch the scan, launch the finalize.  No CPU compute path exists: without a CUDA device
// bfm_create fails and says so.
#include "bfm_kernels.cuh"
#include "../../include/bfm.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

using bfm::Problem;
using bfm::ScanParams;
using bfm::Segment;

constexpr int NT = 128;          // threads per scan CTA
constexpr int FIN_NT = 256;      // threads per finalize CTA
constexpr int MIN_SEG_ROWS = 32; // smallest train range worth a CTA
constexpr int N_TABLE_SLOTS = 4;

std::string g_create_error;
std::mutex g_create_mutex;

typedef void (*ScanFn)(const ScanParams);

template <int R, int MODE, int MASK, int PM>
ScanFn scan_fn() {
    // MODE 0: k=1, 1: k=1 + cross-check, 2: k=2
    return bfm::bfm_scan_kernel<R, (MODE == 2 ? 2 : 1), (MODE == 1), MASK, PM, NT>;
}
template <int R, int MODE, int MASK>
ScanFn pick_pm(int pm) {
    switch (pm) {
        case 4: return scan_fn<R, MODE, MASK, 4>();
        case 5: return scan_fn<R, MODE, MASK, 5>();
        case 6: return scan_fn<R, MODE, MASK, 6>();
        case 40: return scan_fn<R, MODE, MASK, 40>();
        case 50: return scan_fn<R, MODE, MASK, 50>();
        default: return scan_fn<R, MODE, MASK, 8>();
    }
}
template <int R, int MODE>
ScanFn pick_mask(int mask, int pm) {
    switch (mask) {
        case 1: return pick_pm<R, MODE, 1>(pm);
        case 2: return pick_pm<R, MODE, 2>(pm);
        default: return pick_pm<R, MODE, 0>(pm);
    }
}
template <int R>
ScanFn pick_mode(int mode, int mask, int pm) {
    switch (mode) {
        case 1: return pick_mask<R, 1>(mask, pm);
        case 2: return pick_mask<R, 2>(mask, pm);
        default: return pick_mask<R, 0>(mask, pm);
    }
}
ScanFn pick_scan(int r, int mode, int mask, int pm) {
    switch (r) {
        case 1: return pick_mode<1>(mode, mask, pm);
        case 2: return pick_mode<2>(mode, mask, pm);
        default: return pick_mode<4>(mode, mask, pm);
    }
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    unsigned generation = 0;  // bumped on every (re)allocation; the address alone may be reused by cudaMalloc
};

}  // namespace

struct bfm_handle_s {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    DevBuf state;    // rowstate (u64 per out row) followed by colkeys (u32 per problem-train row)
    DevBuf tables;   // device copy of [problems | segments]
    void *h_tables[N_TABLE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};  // pinned staging ring
    size_t h_tables_cap[N_TABLE_SLOTS] = {0, 0, 0, 0};
    cudaEvent_t table_ev[N_TABLE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int table_slot = 0;

    // host-mode staging (BFM_MEM_HOST)
    DevBuf d_in, d_out;
    void *h_out = nullptr;
    size_t h_out_cap = 0;

    // tuning knobs
    int popc_mode = 0, qpt = 0, timing = 0, segment_rows = 0, waves = 0, pipeline_chunks = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};

    bfm_launch_info_t info{};
    int64_t launches = 0;
    std::vector<Segment> segs_host;
    std::vector<int> seg_begin;  // first segment of every problem (+ end sentinel)
    std::vector<Problem> probs_host;
    // plan cache + workspace hygiene
    bool plan_valid = false, state_clean = false;
    int plan_sig[6] = {0, 0, 0, 0, 0, 0};
    int plan_seg_rows = 0;
    std::vector<bfm_problem_t> plan_problems;
    // pipelined host path: copy-in / copy-out streams and per-chunk events
    cudaStream_t in_stream = nullptr, out_stream = nullptr;
    cudaEvent_t chunk_ev[2 * 16] = {};
    int occ_cache[3][3][3][6];  // [R idx][mode][mask][pm idx] -> CTAs per SM (0 = unknown)
};

namespace {

int fail(bfm_handle_t h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

#define CU_TRY(h, call)                                                                           \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(h, BFM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

int ensure(bfm_handle_t h, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return BFM_OK;
    if (b.p) {
        // the buffer may still be in use by work queued on a caller stream
        CU_TRY(h, cudaDeviceSynchronize());
        CU_TRY(h, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = std::max(bytes, (size_t)1 << 16);
    want += want / 4;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) return fail(h, BFM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    b.cap = want;
    ++b.generation;
    return BFM_OK;
}

int pm_index(int pm) { return pm == 4 ? 0 : pm == 5 ? 1 : pm == 6 ? 2 : pm == 40 ? 4 : pm == 50 ? 5 : 3; }
int r_index(int r) { return r == 1 ? 0 : r == 2 ? 1 : 2; }

int occupancy(bfm_handle_t h, int r, int mode, int mask, int pm, int *out) {
    int &c = h->occ_cache[r_index(r)][mode][mask][pm_index(pm)];
    if (c == 0) {
        int n = 0;
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pick_scan(r, mode, mask, pm), NT, 0));
        c = std::max(n, 1);
    }
    *out = c;
    return BFM_OK;
}

// Cut every problem into (query block, train range) segments of near-equal cost so that the grid
// is a few balanced waves over all SMs, whatever the batch shape.
void plan_segments(bfm_handle_t h, const bfm_problem_t *problems, int n_problems, int r, int slots, int groups,
                   std::vector<Segment> &segs, std::vector<int> &seg_begin, int *seg_rows_out) {
    const int bq = NT * r;
    long long steps = 0;  // sum over query blocks of their train rows
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        if (pr.q_count <= 0 || pr.t_count <= 0) continue;
        steps += (long long)((pr.q_count + bq - 1) / bq) * pr.t_count;
    }
    // `groups` launches share this plan (pipelined host path): each of them should still be a few waves
    const int ng = std::max(groups, 1);
    int L;
    if (h->segment_rows > 0) {
        L = h->segment_rows;
    } else if (h->waves > 0) {
        const long long target = (long long)slots * h->waves * ng;
        L = (int)std::max<long long>(MIN_SEG_ROWS, (steps + target - 1) / target);
    } else {
        // All CTAs of a launch cost about the same, so a grid of n CTAs over `slots` resident ones runs
        // about ceil(n / slots) waves; avoid a nearly empty last wave.  (Measured effect on the 256-pair
        // batch is small - 7168 CTAs = 6.05 waves: 1.042 ms, 8192 = 6.92 waves: 1.039 ms - because CTA
        // start times drift apart, but it costs nothing.)  Search the segment length between ~3 and ~12
        // waves per launch for the best wave efficiency, discounted by the per-CTA fixed cost (~4 rows).
        const long long lo = std::max<long long>(MIN_SEG_ROWS, steps / ((long long)slots * 12 * ng));
        const long long hi = std::max<long long>(lo, steps / ((long long)slots * 3 * ng) + 1);
        double best = -1.0;
        L = (int)lo;
        const long long stride = std::max<long long>(1, (hi - lo) / 256);
        for (long long cand = hi; cand >= lo; cand -= stride) {
            long long n = 0;
            for (int p = 0; p < n_problems; ++p) {
                const bfm_problem_t &pr = problems[p];
                if (pr.q_count <= 0 || pr.t_count <= 0) continue;
                n += (long long)((pr.q_count + bq - 1) / bq) * ((pr.t_count + cand - 1) / cand);
            }
            const double per_launch = (double)n / ng;
            const double waves = per_launch / slots;
            const double eff = waves / std::ceil(waves - 1e-9);
            const double score = eff * (double)cand / ((double)cand + 4.0);
            if (score > best + 1e-9) {
                best = score;
                L = (int)cand;
            }
        }
    }
    *seg_rows_out = L;
    segs.clear();
    seg_begin.assign((size_t)n_problems + 1, 0);
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        seg_begin[p] = (int)segs.size();
        seg_begin[p + 1] = (int)segs.size();
        if (pr.q_count <= 0 || pr.t_count <= 0) continue;
        const int nsp = (pr.t_count + L - 1) / L;
        const int base = pr.t_count / nsp, rem = pr.t_count % nsp;
        for (int qb = 0; qb * bq < pr.q_count; ++qb) {
            int t0 = 0;
            for (int s = 0; s < nsp; ++s) {
                const int cnt = base + (s < rem ? 1 : 0);
                Segment sg;
                sg.q_row0 = pr.q_begin + qb * bq;
                sg.q_valid = std::min(bq, pr.q_count - qb * bq);
                sg.q_local0 = qb * bq;
                sg.out_row0 = pr.out_begin + qb * bq;
                sg.t_row0 = pr.t_begin + t0;
                sg.t_count = cnt;
                sg.t_local0 = t0;
                sg.col0 = 0;  // filled with the problem's column-key base by the caller
                segs.push_back(sg);
                t0 += cnt;
            }
        }
        seg_begin[p + 1] = (int)segs.size();
    }
}

// Optional launch grouping (pipelined host path): the batch is planned and its tables uploaded once,
// then problems [bounds[g], bounds[g+1]) are scanned + finalized as launch group g, with the hooks
// called around each group to tie it to the copy streams.
struct GroupHooks {
    int n_groups;
    const int *bounds;
    void *ctx;
    int (*before)(void *ctx, int g);
    int (*after)(void *ctx, int g);
};

int check_opts(bfm_handle_t h, const bfm_options_t *o, int n_problems) {
    if (!o) return fail(h, BFM_ERR_INVALID, "options is NULL");
    if (o->k < 1) return fail(h, BFM_ERR_INVALID, "k must be >= 1");
    if (o->k > 2) return fail(h, BFM_ERR_UNSUPPORTED, "k > 2 is not supported by this build");
    if (o->cross_check && o->k != 1) return fail(h, BFM_ERR_INVALID, "cross_check requires k == 1 (cv2 asserts the same)");
    if (o->cross_check && o->ratio >= 0) return fail(h, BFM_ERR_INVALID, "cross_check and ratio are exclusive");
    if (o->mask_kind < 0 || o->mask_kind > 2) return fail(h, BFM_ERR_INVALID, "bad mask_kind");
    if (o->mask_kind == BFM_MASK_DENSE) {
        if (n_problems != 1) return fail(h, BFM_ERR_INVALID, "a dense mask is only accepted for a single problem");
        if (!o->mask) return fail(h, BFM_ERR_INVALID, "mask_kind is DENSE but mask is NULL");
    }
    if (o->mask_kind == BFM_MASK_WINDOW && (!o->q_xy || !o->t_xy))
        return fail(h, BFM_ERR_INVALID, "mask_kind is WINDOW but q_xy / t_xy is NULL");
    return BFM_OK;
}

// The device path: every data pointer is a device pointer, work is queued on `st`.
int run_device(bfm_handle_t h, const uint8_t *q, int32_t nq_rows, const uint8_t *t, int32_t nt_rows,
               const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
               const bfm_options_t *o, int32_t *knn_idx, int32_t *knn_dist, int32_t *m_query,
               int32_t *m_train, int32_t *m_dist, int32_t *m_count, cudaStream_t st,
               const GroupHooks *hooks = nullptr) {
    h->info = bfm_launch_info_t{};
    if (n_problems <= 0 || n_out_rows <= 0) return BFM_OK;
    const int one_group[2] = {0, n_problems};
    const int n_groups = hooks ? hooks->n_groups : 1;
    const int *gbounds = hooks ? hooks->bounds : one_group;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(t)) & 15)
        return fail(h, BFM_ERR_INVALID, "descriptor arrays must be 16-byte aligned");
    const bool want_matches = m_count != nullptr;
    if (want_matches && (!m_query || !m_train || !m_dist))
        return fail(h, BFM_ERR_INVALID, "m_query/m_train/m_dist/m_count must be given together");
    if ((knn_idx == nullptr) != (knn_dist == nullptr))
        return fail(h, BFM_ERR_INVALID, "knn_idx and knn_dist must be given together");

    // -- validate problems, lay out column keys ------------------------------------------------
    h->probs_host.resize(n_problems);
    long long col_rows = 0;
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        if (pr.q_count < 0 || pr.t_count < 0 || pr.q_begin < 0 || pr.t_begin < 0 || pr.out_begin < 0 ||
            (long long)pr.q_begin + pr.q_count > nq_rows || (long long)pr.t_begin + pr.t_count > nt_rows ||
            (long long)pr.out_begin + pr.q_count > n_out_rows)
            return fail(h, BFM_ERR_INVALID, "problem " + std::to_string(p) + " is out of range");
        if (pr.t_count >= BFM_MAX_TRAIN_ROWS || pr.q_count >= BFM_MAX_QUERY_ROWS)
            return fail(h, BFM_ERR_INVALID, "problem " + std::to_string(p) + " exceeds 2^22 rows");
        Problem d;
        d.q_begin = pr.q_begin; d.q_count = pr.q_count; d.t_begin = pr.t_begin; d.t_count = pr.t_count;
        d.out_begin = pr.out_begin;
        d.col0 = (int32_t)col_rows;  // column-key base of this problem
        h->probs_host[p] = d;
        if (o->cross_check) col_rows += pr.t_count;
    }
    if (col_rows >= (1ll << 31)) return fail(h, BFM_ERR_INVALID, "batch too large for cross-check");

    // -- choose the kernel variant ----------------------------------------------------------------
    const int mode = o->cross_check ? 1 : ((o->k >= 2 || o->ratio >= 0) ? 2 : 0);
    const int mask = o->mask_kind;
    const int pm = h->popc_mode ? h->popc_mode : 40;  // measured best for every mode: profiles/sweep_r01.md
    int r = h->qpt;
    int slots = 0, seg_rows = 0;
    if (r != 1 && r != 2 && r != 4) {
        // largest register tile that still leaves >= 2 work items per CTA slot
        for (int cand : {4, 2, 1}) {
            int occ = 0;
            int rc = occupancy(h, cand, mode, mask, pm, &occ);
            if (rc) return rc;
            long long units = 0;
            for (int p = 0; p < n_problems; ++p)
                if (problems[p].q_count > 0 && problems[p].t_count > 0)
                    units += (long long)((problems[p].q_count + NT * cand - 1) / (NT * cand)) *
                             ((problems[p].t_count + 2 * MIN_SEG_ROWS - 1) / (2 * MIN_SEG_ROWS));
            r = cand;
            if (units >= 2ll * occ * h->sm_count) break;
        }
    }
    {
        int occ = 0;
        int rc = occupancy(h, r, mode, mask, pm, &occ);
        if (rc) return rc;
        slots = occ * h->sm_count;
    }
    // -- plan cache: same problems + same variant as the previous call -> the device tables are
    //    already in place (steady state of a tracking loop with fixed shapes, bench loops) ----------
    const int plan_sig[6] = {n_problems, r, mode * 64 + n_groups, h->segment_rows, h->waves, slots};
    const bool plan_hit = h->plan_valid && std::memcmp(plan_sig, h->plan_sig, sizeof(plan_sig)) == 0 &&
                          h->plan_problems.size() == (size_t)n_problems &&
                          std::memcmp(h->plan_problems.data(), problems, sizeof(bfm_problem_t) * (size_t)n_problems) == 0;
    if (!plan_hit) {
    plan_segments(h, problems, n_problems, r, slots, n_groups, h->segs_host, h->seg_begin, &seg_rows);
    h->plan_seg_rows = seg_rows;
        if (o->cross_check) {
        // segment order follows problem order: recover each segment's problem by walking
        size_t si = 0;
        for (int p = 0; p < n_problems; ++p) {
            const bfm_problem_t &pr = problems[p];
            if (pr.q_count <= 0 || pr.t_count <= 0) continue;
            const int nsp = (pr.t_count + seg_rows - 1) / seg_rows;
            const int nqb = (pr.q_count + NT * r - 1) / (NT * r);
            for (int i = 0; i < nsp * nqb; ++i, ++si) h->segs_host[si].col0 = h->probs_host[p].col0;
        }
    }
    }  // !plan_hit
    seg_rows = h->plan_seg_rows;
    const size_t n_segs = h->segs_host.size();

    // -- workspace: self-cleaning (the finalize kernel restores every slot it read to all-ones), so
    //    a memset is only queued after (re)allocation or after a call that failed half-way -----------
    const size_t state_bytes = (size_t)n_out_rows * 8;
    const size_t col_bytes = (size_t)col_rows * 4;
    const unsigned state_gen = h->state.generation;
    int rc = ensure(h, h->state, state_bytes + col_bytes);
    if (rc) return rc;
    if (h->state.generation != state_gen) h->state_clean = false;
    unsigned long long *rowstate = static_cast<unsigned long long *>(h->state.p);
    uint32_t *colkeys = reinterpret_cast<uint32_t *>(static_cast<char *>(h->state.p) + state_bytes);

    const size_t prob_bytes = ((size_t)n_problems * sizeof(Problem) + 15) & ~(size_t)15;
    const size_t table_bytes = prob_bytes + n_segs * sizeof(Segment);
    const unsigned tables_gen = h->tables.generation;
    rc = ensure(h, h->tables, table_bytes);
    if (rc) return rc;
    if (!plan_hit || h->tables.generation != tables_gen) {
    h->plan_valid = false;
    const int slot = h->table_slot;
    h->table_slot = (slot + 1) % N_TABLE_SLOTS;
    if (h->h_tables_cap[slot] < table_bytes) {
        if (h->h_tables[slot]) {
            CU_TRY(h, cudaEventSynchronize(h->table_ev[slot]));
            CU_TRY(h, cudaFreeHost(h->h_tables[slot]));
            h->h_tables[slot] = nullptr;
        }
        const size_t want = table_bytes + table_bytes / 2 + 4096;
        CU_TRY(h, cudaMallocHost(&h->h_tables[slot], want));
        h->h_tables_cap[slot] = want;
    } else {
        CU_TRY(h, cudaEventSynchronize(h->table_ev[slot]));  // previous upload from this slot is done
    }
    std::memcpy(h->h_tables[slot], h->probs_host.data(), (size_t)n_problems * sizeof(Problem));
    if (n_segs) std::memcpy(static_cast<char *>(h->h_tables[slot]) + prob_bytes, h->segs_host.data(), n_segs * sizeof(Segment));
    // NOTE: the device table is shared by consecutive calls on one handle; stream order keeps the
    // upload of call n+1 behind the kernels of call n when both use the same stream (the contract).
    CU_TRY(h, cudaMemcpyAsync(h->tables.p, h->h_tables[slot], table_bytes, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaEventRecord(h->table_ev[slot], st));
    std::memcpy(h->plan_sig, plan_sig, sizeof(plan_sig));
    h->plan_problems.assign(problems, problems + n_problems);
    h->plan_valid = true;
    }
    const Problem *d_probs = static_cast<const Problem *>(h->tables.p);
    const Segment *d_segs = reinterpret_cast<const Segment *>(static_cast<char *>(h->tables.p) + prob_bytes);

    if (h->timing) CU_TRY(h, cudaEventRecord(h->ev[0], st));
    if (!h->state_clean) CU_TRY(h, cudaMemsetAsync(h->state.p, 0xFF, h->state.cap, st));
    h->state_clean = false;  // set again once the finalize kernel (which restores the state) is queued

    int kernels = 0;
    ScanParams sp;
    sp.q = reinterpret_cast<const uint4 *>(q);
    sp.t = reinterpret_cast<const uint4 *>(t);
    sp.rowstate = rowstate;
    sp.colkeys = colkeys;
    sp.mask = o->mask;
    sp.mask_stride = o->mask_row_stride;
    sp.q_xy = reinterpret_cast<const float2 *>(o->q_xy);
    sp.t_xy = reinterpret_cast<const float2 *>(o->t_xy);
    sp.radius = o->window_radius;
    sp.mul_d32 = 1u << bfm::DIST_SHIFT;
    sp.mul_lo16 = 1u << 7;
    sp.mul_hi16 = 1u << 23;
    sp.mul_one = 1u;
    sp.mul_two = 2u;
    sp.mul_four = 4u;
    bfm::FinalizeParams fp;
    fp.rowstate = rowstate;
    fp.colkeys = colkeys;
    fp.k = o->k;
    fp.cross_check = o->cross_check;
    fp.max_distance = o->max_distance;
    fp.use_ratio = o->ratio >= 0;
    fp.ratio = o->ratio;
    fp.knn_idx = knn_idx;
    fp.knn_dist = knn_dist;
    fp.m_query = m_query;
    fp.m_train = m_train;
    fp.m_dist = m_dist;
    const ScanFn fn = pick_scan(r, mode, mask, pm);
    for (int g = 0; g < n_groups; ++g) {
        const int p0 = gbounds[g], p1 = gbounds[g + 1];
        if (p1 <= p0) continue;
        if (hooks && hooks->before) {
            rc = hooks->before(hooks->ctx, g);
            if (rc) return rc;
        }
        const int s0 = h->seg_begin[p0], s1 = h->seg_begin[p1];
        if (s1 > s0) {
            sp.segs = d_segs + s0;
            if (h->timing && g == 0) CU_TRY(h, cudaEventRecord(h->ev[1], st));
            fn<<<(unsigned)(s1 - s0), NT, 0, st>>>(sp);
            CU_TRY(h, cudaGetLastError());
            if (h->timing && g == n_groups - 1) CU_TRY(h, cudaEventRecord(h->ev[2], st));
            ++kernels;
        }
        fp.problems = d_probs + p0;
        fp.m_count = m_count ? m_count + p0 : nullptr;
        bfm::bfm_finalize_kernel<FIN_NT><<<(unsigned)(p1 - p0), FIN_NT, 0, st>>>(fp);
        CU_TRY(h, cudaGetLastError());
        ++kernels;
        if (hooks && hooks->after) {
            rc = hooks->after(hooks->ctx, g);
            if (rc) return rc;
        }
    }
    h->state_clean = true;  // every slot touched above is restored by its group's finalize kernel
    if (h->timing) CU_TRY(h, cudaEventRecord(h->ev[3], st));

    h->launches += kernels;
    h->info.kernels_launched = kernels;
    h->info.scan_grid = (int32_t)n_segs;
    h->info.scan_block = NT;
    h->info.queries_per_thread = r;
    h->info.popc_mode = pm;
    h->info.segments = (int32_t)n_segs;
    h->info.train_rows_per_segment = seg_rows;
    if (h->timing) {
        CU_TRY(h, cudaEventSynchronize(h->ev[3]));
        if (n_segs) CU_TRY(h, cudaEventElapsedTime(&h->info.scan_ms, h->ev[1], h->ev[2]));
        CU_TRY(h, cudaEventElapsedTime(&h->info.total_ms, h->ev[0], h->ev[3]));
    }
    if (std::getenv("BFM_CHECK_CLEAN")) {  // debugging aid: the workspace must be all-ones after every call
        CU_TRY(h, cudaDeviceSynchronize());
        std::vector<unsigned char> host(h->state.cap);
        CU_TRY(h, cudaMemcpy(host.data(), h->state.p, h->state.cap, cudaMemcpyDeviceToHost));
        size_t bad = 0, first = 0;
        for (size_t i = 0; i < host.size(); ++i)
            if (host[i] != 0xFF) { if (!bad) first = i; ++bad; }
        if (bad)
            std::fprintf(stderr, "[bfm check] workspace dirty after call: %zu bytes, first at %zu (rows=%d col_rows=%lld state_bytes=%zu cap=%zu mode=%d P=%d groups=%d)\n",
                         bad, first, n_out_rows, col_rows, state_bytes, h->state.cap, mode, n_problems, n_groups);
    }
    return BFM_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// The host path: stage in, run the device path on the handle's stream, stage out, wait.
int run_host(bfm_handle_t h, const uint8_t *q, int32_t nq_rows, const uint8_t *t, int32_t nt_rows,
             const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
             const bfm_options_t *o, int32_t *knn_idx, int32_t *knn_dist, int32_t *m_query,
             int32_t *m_train, int32_t *m_dist, int32_t *m_count) {
    if (n_problems <= 0 || n_out_rows <= 0) return BFM_OK;
    cudaStream_t st = h->stream;
    const size_t qb = (size_t)nq_rows * 32, tb = (size_t)nt_rows * 32;
    size_t mask_b = 0, qxy_b = 0, txy_b = 0;
    if (o->mask_kind == BFM_MASK_DENSE && o->mask) mask_b = problems[0].q_count > 0 ? (size_t)(problems[0].q_count - 1) * (size_t)o->mask_row_stride + (size_t)problems[0].t_count : 0;
    if (o->mask_kind == BFM_MASK_WINDOW) { qxy_b = (size_t)nq_rows * 8; txy_b = (size_t)nt_rows * 8; }
    const size_t o_q = 0, o_t = align256(o_q + qb), o_m = align256(o_t + tb), o_qxy = align256(o_m + mask_b),
                 o_txy = align256(o_qxy + qxy_b), in_total = align256(o_txy + txy_b);
    int rc = ensure(h, h->d_in, in_total);
    if (rc) return rc;
    char *din = static_cast<char *>(h->d_in.p);
    if (qb) CU_TRY(h, cudaMemcpyAsync(din + o_q, q, qb, cudaMemcpyHostToDevice, st));
    if (tb) CU_TRY(h, cudaMemcpyAsync(din + o_t, t, tb, cudaMemcpyHostToDevice, st));
    bfm_options_t od = *o;
    if (mask_b) {
        CU_TRY(h, cudaMemcpyAsync(din + o_m, o->mask, mask_b, cudaMemcpyHostToDevice, st));
        od.mask = reinterpret_cast<const uint8_t *>(din + o_m);
    }
    if (qxy_b) {
        CU_TRY(h, cudaMemcpyAsync(din + o_qxy, o->q_xy, qxy_b, cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaMemcpyAsync(din + o_txy, o->t_xy, txy_b, cudaMemcpyHostToDevice, st));
        od.q_xy = reinterpret_cast<const float *>(din + o_qxy);
        od.t_xy = reinterpret_cast<const float *>(din + o_txy);
    }
    // outputs: [knn_idx | knn_dist | m_query | m_train | m_dist | m_count] in one block
    const size_t knn_b = knn_idx ? (size_t)n_out_rows * o->k * 4 : 0;
    const size_t m_b = m_count ? (size_t)n_out_rows * 4 : 0;
    const size_t cnt_b = m_count ? (size_t)n_problems * 4 : 0;
    const size_t out_total = 2 * knn_b + 3 * m_b + cnt_b;
    rc = ensure(h, h->d_out, std::max<size_t>(out_total, 16));
    if (rc) return rc;
    if (h->h_out_cap < out_total) {
        if (h->h_out) CU_TRY(h, cudaFreeHost(h->h_out));
        h->h_out = nullptr;
        const size_t want = out_total + out_total / 4 + 4096;
        CU_TRY(h, cudaMallocHost(&h->h_out, want));
        h->h_out_cap = want;
    }
    char *dout = static_cast<char *>(h->d_out.p);
    int32_t *d_ki = knn_idx ? reinterpret_cast<int32_t *>(dout) : nullptr;
    int32_t *d_kd = knn_idx ? reinterpret_cast<int32_t *>(dout + knn_b) : nullptr;
    int32_t *d_mq = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b) : nullptr;
    int32_t *d_mt = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b + m_b) : nullptr;
    int32_t *d_md = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b + 2 * m_b) : nullptr;
    int32_t *d_mc = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b + 3 * m_b) : nullptr;
    rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows,
                    reinterpret_cast<const uint8_t *>(din + o_t), nt_rows, problems, n_problems, n_out_rows,
                    &od, d_ki, d_kd, d_mq, d_mt, d_md, d_mc, st);
    if (rc) return rc;
    if (out_total) CU_TRY(h, cudaMemcpyAsync(h->h_out, dout, out_total, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    const char *ho = static_cast<const char *>(h->h_out);
    if (knn_idx) {
        std::memcpy(knn_idx, ho, knn_b);
        std::memcpy(knn_dist, ho + knn_b, knn_b);
    }
    if (m_count) {
        std::memcpy(m_count, ho + 2 * knn_b + 3 * m_b, cnt_b);
        // only the filled prefix of every problem's slice is meaningful; copy exactly that
        const int32_t *hq = reinterpret_cast<const int32_t *>(ho + 2 * knn_b);
        const int32_t *ht = reinterpret_cast<const int32_t *>(ho + 2 * knn_b + m_b);
        const int32_t *hd = reinterpret_cast<const int32_t *>(ho + 2 * knn_b + 2 * m_b);
        for (int p = 0; p < n_problems; ++p) {
            const size_t b = (size_t)problems[p].out_begin, n = (size_t)m_count[p];
            std::memcpy(m_query + b, hq + b, n * 4);
            std::memcpy(m_train + b, ht + b, n * 4);
            std::memcpy(m_dist + b, hd + b, n * 4);
        }
    }
    return BFM_OK;
}

#include "bfm_pipeline.cuh"

}  // namespace

extern "C" {

int bfm_abi_version(void) { return BFM_ABI_VERSION; }

int bfm_create(int device, bfm_handle_t *out) {
    if (!out) return BFM_ERR_INVALID;
    *out = nullptr;
    std::lock_guard<std::mutex> lock(g_create_mutex);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this engine has no CPU path)";
        return BFM_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        g_create_error = "device index out of range";
        return BFM_ERR_INVALID;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        return BFM_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_error = "libbfm_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return BFM_ERR_UNSUPPORTED;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        return BFM_ERR_CUDA;
    }
    bfm_handle_t h = new bfm_handle_s();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    std::memset(h->occ_cache, 0, sizeof(h->occ_cache));
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 32; ++i) ok = cudaEventCreate(&h->chunk_ev[i]) == cudaSuccess;
    for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    for (int i = 0; ok && i < N_TABLE_SLOTS; ++i) {
        ok = cudaEventCreateWithFlags(&h->table_ev[i], cudaEventDisableTiming) == cudaSuccess;
        if (ok) ok = cudaEventRecord(h->table_ev[i], h->stream) == cudaSuccess;
    }
    if (!ok) {
        g_create_error = std::string("stream/event creation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete h;
        return BFM_ERR_CUDA;
    }
    *out = h;
    return BFM_OK;
}

int bfm_destroy(bfm_handle_t h) {
    if (!h) return BFM_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&h->state, &h->tables, &h->d_in, &h->d_out})
        if (b->p) cudaFree(b->p);
    for (int i = 0; i < N_TABLE_SLOTS; ++i) {
        if (h->h_tables[i]) cudaFreeHost(h->h_tables[i]);
        if (h->table_ev[i]) cudaEventDestroy(h->table_ev[i]);
    }
    if (h->h_out) cudaFreeHost(h->h_out);
    for (int i = 0; i < 4; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 32; ++i)
        if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
    if (h->in_stream) cudaStreamDestroy(h->in_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return BFM_OK;
}

const char *bfm_last_error(bfm_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int bfm_match_batched(bfm_handle_t h, int mem, const uint8_t *q, int32_t n_query_rows, const uint8_t *t,
                      int32_t n_train_rows, const bfm_problem_t *problems, int32_t n_problems,
                      int32_t n_out_rows, const bfm_options_t *opts, int32_t *knn_idx, int32_t *knn_dist,
                      int32_t *m_query, int32_t *m_train, int32_t *m_dist, int32_t *m_count, void *stream) {
    if (!h) return BFM_ERR_INVALID;
    h->err.clear();
    if (n_problems < 0 || n_query_rows < 0 || n_train_rows < 0 || n_out_rows < 0)
        return fail(h, BFM_ERR_INVALID, "negative size");
    if (n_problems > 0 && !problems) return fail(h, BFM_ERR_INVALID, "problems is NULL");
    int rc = check_opts(h, opts, n_problems);
    if (rc) return rc;
    if ((n_query_rows > 0 && !q) || (n_train_rows > 0 && !t)) return fail(h, BFM_ERR_INVALID, "descriptor pointer is NULL");
    CU_TRY(h, cudaSetDevice(h->device));
    if (mem == BFM_MEM_DEVICE)
        return run_device(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, knn_idx,
                          knn_dist, m_query, m_train, m_dist, m_count, stream == BFM_STREAM_OWN ? h->stream : static_cast<cudaStream_t>(stream));
    if (mem == BFM_MEM_HOST && pipeline_eligible(h, n_query_rows, n_train_rows, problems, n_problems, opts))
        return run_host_pipelined(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, knn_idx,
                                  knn_dist, m_query, m_train, m_dist, m_count);
    if (mem == BFM_MEM_HOST)
        return run_host(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, knn_idx,
                        knn_dist, m_query, m_train, m_dist, m_count);
    return fail(h, BFM_ERR_INVALID, "mem must be BFM_MEM_HOST or BFM_MEM_DEVICE");
}

int bfm_knn(bfm_handle_t h, int mem, const uint8_t *q, int32_t nq, const uint8_t *t, int32_t nt,
            const bfm_options_t *opts, int32_t *knn_idx, int32_t *knn_dist, void *stream) {
    bfm_problem_t pr = {0, nq, 0, nt, 0, 0};
    return bfm_match_batched(h, mem, q, nq, t, nt, &pr, 1, nq, opts, knn_idx, knn_dist, nullptr, nullptr, nullptr,
                             nullptr, stream);
}

int bfm_match(bfm_handle_t h, int mem, const uint8_t *q, int32_t nq, const uint8_t *t, int32_t nt,
              const bfm_options_t *opts, int32_t *m_query, int32_t *m_train, int32_t *m_dist, int32_t *m_count,
              void *stream) {
    bfm_problem_t pr = {0, nq, 0, nt, 0, 0};
    return bfm_match_batched(h, mem, q, nq, t, nt, &pr, 1, nq, opts, nullptr, nullptr, m_query, m_train, m_dist,
                             m_count, stream);
}

int bfm_get_launch_info(bfm_handle_t h, bfm_launch_info_t *out) {
    if (!h || !out) return BFM_ERR_INVALID;
    *out = h->info;
    return BFM_OK;
}

int bfm_set_tuning(bfm_handle_t h, const char *knob, int32_t value) {
    if (!h || !knob) return BFM_ERR_INVALID;
    const std::string k(knob);
    if (k == "popc_mode") {
        if (value != 0 && value != 4 && value != 5 && value != 6 && value != 8 && value != 40 && value != 50)
            return fail(h, BFM_ERR_INVALID, "popc_mode must be 0,4,5,6,8,40,50");
        h->popc_mode = value;
    } else if (k == "queries_per_thread") {
        if (value != 0 && value != 1 && value != 2 && value != 4) return fail(h, BFM_ERR_INVALID, "queries_per_thread must be 0,1,2,4");
        h->qpt = value;
    } else if (k == "timing") {
        h->timing = value != 0;
    } else if (k == "segment_rows") {
        if (value < 0) return fail(h, BFM_ERR_INVALID, "segment_rows must be >= 0");
        h->segment_rows = value;
    } else if (k == "pipeline_chunks") {
        if (value < 0 || value > 16) return fail(h, BFM_ERR_INVALID, "pipeline_chunks must be 0 (auto), 1 (off) .. 16");
        h->pipeline_chunks = value;
    } else if (k == "waves") {
        if (value < 0) return fail(h, BFM_ERR_INVALID, "waves must be >= 0");
        h->waves = value;
    } else {
        return fail(h, BFM_ERR_INVALID, "unknown tuning knob: " + k);
    }
    return BFM_OK;
}

int64_t bfm_kernel_launch_count(bfm_handle_t h) { return h ? h->launches : 0; }

int bfm_host_alloc(uint64_t bytes, void **out) {
    if (!out) return BFM_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_create_mutex);
        g_create_error = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
        return BFM_ERR_NOMEM;
    }
    return BFM_OK;
}

int bfm_host_free(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) return BFM_ERR_CUDA;
    return BFM_OK;
}

int bfm_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz, char *name, int name_len) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BFM_ERR_CUDA;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *clock_khz = khz;
    }
    if (name && name_len > 0) {
        std::strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return BFM_OK;
}

}  // extern "C"
