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
#include "bfm_workers.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace bfm {
#include "bfm_window.cuh"
}

namespace {

using bfm::Problem;
using bfm::ScanParams;
using bfm::Segment;

constexpr int NT = 128;          // threads per scan CTA
constexpr int MIN_SEG_ROWS = 32; // smallest train range worth a CTA
// tapered tail of a batch (>= 8 problems): the last 10 % of the work is cut into segments of 1/2, then 1/4 of the
// length - 256-pair batch 1046 -> 1026 us (980 -> 998 G pairs/s), pinned host path 1120 -> 1104 us
// (tools/taper_probe.py, profiles/r01f_taper_probe.log)
constexpr int GSS_MAX_ROWS = 512;
constexpr long long PERSISTENT_MAX_PAIRS = 160ll * 1000 * 1000;   // launches above this (~0.15 ms) take the static form
constexpr long long TENSOR_MIN_PAIRS = 8ll * 1000 * 1000;         // eligible launches from this size on take the tensor form
constexpr int TAPER_AUTO = 4;
constexpr int TAPER_PCT_AUTO = 10;
constexpr int N_TABLE_SLOTS = 4;
constexpr int MAX_COPY_CHUNKS = 64;  // input chunks of the pipelined host path

std::string g_create_error;
std::mutex g_create_mutex;

using bfm::ScanFn;

// the kernel variants are instantiated in bfm_scan_inst.cu, one object per (register tile, mode)
ScanFn pick_scan(int r, int mode, int mask, int pm, bool bound, bool dyn) {
    if (bound) return bfm::pick_scan_r1_m2(mask, pm, true, dyn);   // k > 2 passes: R = 1, K = 2, transformed carry-save popcount
    switch (r * 10 + mode) {
        case 10: return bfm::pick_scan_r1_m0(mask, pm, false, dyn);
        case 11: return bfm::pick_scan_r1_m1(mask, pm, false, dyn);
        case 12: return bfm::pick_scan_r1_m2(mask, pm, false, dyn);
        case 20: return bfm::pick_scan_r2_m0(mask, pm, false, dyn);
        case 21: return bfm::pick_scan_r2_m1(mask, pm, false, dyn);
        case 22: return bfm::pick_scan_r2_m2(mask, pm, false, dyn);
        case 40: return bfm::pick_scan_r4_m0(mask, pm, false, dyn);
        case 41: return bfm::pick_scan_r4_m1(mask, pm, false, dyn);
        default: return bfm::pick_scan_r4_m2(mask, pm, false, dyn);
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
    DevBuf lower;    // k > 2: per-row lower bound handed from one pass to the next
    DevBuf finc;     // persistent form: look-back words of the finalize tiles (epoch-tagged, zeroed on allocation)
    uint32_t fin_epoch = 0;
    int persistent = 0;  // tuning: 0 auto (resident inputs take the persistent form), 1 off
    int gss_div = 0, gss_min = 0;   // tuning: guided item lengths
    int ctas_per_sm = 0;            // tuning: resident CTAs per SM of the persistent form (0 = as many as fit)
    std::vector<int2> cta_tiles_host;
    std::vector<bfm::FinTile> fin_tiles_host;
    int plan_ctas = 0;        // grid of the persistent form
    uint32_t *d_queue = nullptr;   // two ticket counters used by alternate launches (each launch zeroes the other one)
    int queue_phase = 0;
    DevBuf bins;     // binned window search: train rows in grid-cell order + cell table
    DevBuf xq, xt;   // tensor form: the descriptors expanded to one s8 per bit, two planes of [rows + slack][128] bytes
    int tensor_chunks = 0;           // tuning: copy chunks of the chunked host path (0 = auto)
    bool tables_by_kernel = false;   // set by the chunked host path around its launches
    int tensor = 0;  // tuning: 0 auto (large resident batches without mask / cross-check), 1 off, 2 whenever eligible
    long long x_rows[2] = {0, 0};   // rows of the expanded planes the cached plan was made for (queries, train)
    int bins_problems = 0;   // problems the counters of `bins` are laid out for
    void *h_tables[N_TABLE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};  // pinned staging ring
    size_t h_tables_cap[N_TABLE_SLOTS] = {0, 0, 0, 0};
    cudaEvent_t table_ev[N_TABLE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    int table_slot = 0;

    // host-mode staging (BFM_MEM_HOST): device copy of the inputs; pinned staging for results when
    // the caller's output arrays are pageable (the kernel writes results straight into pinned host memory)
    DevBuf d_in;
    void *h_out = nullptr;
    size_t h_out_cap = 0;
    // input gate of the pipelined host path
    unsigned long long *d_ready = nullptr;   // device: two watermarks
    unsigned long long *h_marks = nullptr;   // pinned: per-chunk watermark values (copy sources)
    uint32_t *h_status = nullptr;            // pinned: gate time-out flag written by the kernel
    unsigned long long seq = 0;              // call sequence number (watermark epoch)

    // tuning knobs
    int popc_mode = 0, qpt = 0, timing = 0, segment_rows = 0, waves = 0, pipeline_chunks = 0, window_bins = 0, feeders = 0, feed_rows = 0, test_stall = 0, pipeline_min_kb = 0, taper = 0, taper_pct = 0;
    uint32_t *d_prog = nullptr;   // SM-fed upload: progress words of the feeder CTAs
    uint32_t feed_epoch = 0;      // epoch of the last SM-fed call (1..65535)
    // pageable caller arrays: host threads stage them into pinned memory slice by slice for the feeders
    std::unique_ptr<WorkerPool> pool;
    void *h_stage = nullptr;
    size_t h_stage_cap = 0;
    uint32_t *h_ready = nullptr;  // pinned: staged rounds published by the host (read by the feeder CTAs)
    int host_threads = 0;         // tuning: 0 auto, -1 off
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // [0], [1] around a call's launches; [2], [3] around the tensor scan alone
    // stream hand-over: every call leaves an event on its stream; a call that arrives on a DIFFERENT stream waits
    // for it before it touches the shared workspace / tables (device calls are asynchronous)
    cudaEvent_t last_ev = nullptr;
    cudaStream_t last_stream = nullptr;
    bool last_pending = false;

    unsigned long long *trace = nullptr;   // bfm_debug_timeline
    int trace_cap = 0;
    bfm_launch_info_t info{};
    int64_t launches = 0;
    std::vector<Segment> segs_host;
    std::vector<int> seg_begin;  // first segment of every problem (+ end sentinel)
    std::vector<Problem> probs_host, plan_probs;
    // plan cache + workspace hygiene
    bool plan_valid = false, state_clean = false;
    bool check_clean = false;   // BFM_CHECK_CLEAN in the environment at bfm_create
    int plan_sig[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int plan_seg_rows = 0;
    std::vector<bfm_problem_t> plan_problems;
    // pipelined host path: the copy-in stream
    cudaStream_t in_stream = nullptr;
    cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // tensor form from host memory: one per copy chunk
    cudaEvent_t chunk_done_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // ... and one per matched chunk
    cudaStream_t out_stream = nullptr;   // ... whose results a third stream copies to the host
    DevBuf d_res;                        // ... out of this device block
    int occ_cache[3][3][3][6];  // [R idx][mode][mask][pm idx] -> CTAs per SM (0 = unknown)
};

namespace {

// the device tables of a launch, fetched from their pinned staging slot by the SMs (see run_device)
__global__ void upload_tables_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

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

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int pm_index(int pm) { return pm == 4 ? 0 : pm == 5 ? 1 : pm == 6 ? 2 : pm == 40 ? 4 : pm == 50 ? 5 : 3; }
int r_index(int r) { return r == 1 ? 0 : r == 2 ? 1 : 2; }

int occupancy(bfm_handle_t h, int r, int mode, int mask, int pm, int *out) {
    int &c = h->occ_cache[r_index(r)][mode][mask][pm_index(pm)];
    if (c == 0) {
        int n = 0;
        CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pick_scan(r, mode, mask, pm, false, false), NT, 0));
        c = std::max(n, 1);
    }
    *out = c;
    return BFM_OK;
}

// Cut every problem into (query block, train range) segments of near-equal cost so that the grid
// is a few balanced waves over all SMs, whatever the batch shape.
void plan_segments(bfm_handle_t h, const bfm_problem_t *problems, int n_problems, int r, int slots,
                   std::vector<Segment> &segs, std::vector<int> &seg_begin, int *seg_rows_out, bool guided = false) {   // guided: the persistent form
    const int bq = NT * r;
    long long steps = 0;  // sum over query blocks of their train rows
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        if (pr.q_count <= 0 || pr.t_count <= 0) continue;
        steps += (long long)((pr.q_count + bq - 1) / bq) * pr.t_count;
    }
    int L;
    if (h->segment_rows > 0) {
        L = h->segment_rows;
    } else if (h->waves > 0 || (guided && n_problems == 1)) {
        // (persistent form, one problem: ONE wave - every CTA takes one item and the ticket counter hands the few
        // that remain to whoever is done first; 2000 x 20000: 57.7 -> 56.6 us, 4096^2: 33.2 -> 31.0 us)
        const long long target = (long long)slots * (h->waves > 0 ? h->waves : 1);
        L = (int)std::max<long long>(MIN_SEG_ROWS, (steps + target - 1) / target);
        if (h->waves == 0) {
            // never a few items more than CTAs (8192 x 4096: 1216 items on 1184 CTAs took 71 us, the transposed problem
            // with exactly 1184 items 51 us): every query block gets floor(slots / blocks) equal train ranges
            const bfm_problem_t &pr = problems[0];
            const long long nqb = std::max(1, (pr.q_count + bq - 1) / bq);
            const long long nsp = std::max<long long>(1, slots / nqb);
            L = (int)std::max<long long>(L, (pr.t_count + nsp - 1) / nsp);
        }
    } else {
        // All CTAs of a launch cost about the same, so a grid of n CTAs over `slots` resident ones runs
        // about ceil(n / slots) waves; avoid a nearly empty last wave.  (Measured effect on the 256-pair
        // batch is small - 7168 CTAs = 6.05 waves: 1.042 ms, 8192 = 6.92 waves: 1.039 ms - because CTA
        // start times drift apart, but it costs nothing.)  Search the segment length between ~3 and ~12
        // waves for the best wave efficiency, discounted by the per-CTA fixed cost (~4 rows).
        const long long lo = std::max<long long>(MIN_SEG_ROWS, steps / ((long long)slots * 12));
        const long long hi = std::max<long long>(lo, steps / ((long long)slots * 3) + 1);
        // (the search visits ~256 candidate lengths; a batch is mostly a few distinct shapes, so it runs over the
        // histogram of (query blocks, train rows) instead of over every problem: 256 equal pairs cost 0.65 ms otherwise,
        // which a loop-closing step with a fresh candidate list would pay every time)
        std::vector<std::pair<std::pair<int, int>, long long>> shapes;
        for (int p = 0; p < n_problems; ++p) {
            const bfm_problem_t &pr = problems[p];
            if (pr.q_count <= 0 || pr.t_count <= 0) continue;
            const std::pair<int, int> key((pr.q_count + bq - 1) / bq, pr.t_count);
            if (!shapes.empty() && shapes.back().first == key) { ++shapes.back().second; continue; }
            bool found = false;
            for (size_t i = 0; i < shapes.size() && i < 16 && !found; ++i)
                if (shapes[i].first == key) { ++shapes[i].second; found = true; }
            if (!found) shapes.emplace_back(key, 1);
        }
        double best = -1.0;
        L = (int)lo;
        const long long n_cand = std::max<long long>(8, std::min<long long>(256, 20000 / std::max<size_t>(shapes.size(), 1)));   // bounded work for ragged batches
        const long long stride = std::max<long long>(1, (hi - lo) / n_cand);
        for (long long cand = hi; cand >= lo; cand -= stride) {
            long long n = 0;
            for (const auto &sh : shapes) n += sh.second * sh.first.first * ((sh.first.second + cand - 1) / cand);
            const double waves = (double)n / slots;
            const double eff = waves / std::ceil(waves - 1e-9);
            const double score = eff * (double)cand / ((double)cand + 4.0);
            if (score > best + 1e-9) {
                best = score;
                L = (int)cand;
            }
        }
    }
    // Guided item lengths (taper = 16; an experiment kept for profiles/r02_kernel_forms.md, never chosen automatically):
    // an item is 1 / gss_div of an even share of the work that is LEFT when it starts.  With the persistent form it
    // LOSES to equal lengths: the CTAs the warp schedulers serve last hold their first - longest - item to the end.
    const bool gss = h->taper == 16;
    const int gss_min = h->gss_min > 0 ? h->gss_min : (r == 1 ? 32 : 16);
    const long long gss_div = (long long)(h->gss_div > 0 ? h->gss_div : 3) * slots;
    if (gss) L = (int)std::max<long long>(gss_min, std::min<long long>(GSS_MAX_ROWS, steps / gss_div + 1));
    *seg_rows_out = L;
    // taper knob: 0 = auto, 1 = off, 2 / 4 / 8 = finest divisor of the segment length in the tail, 16 = guided
    const int taper_div = (gss || h->segment_rows > 0 || n_problems < 8) ? 1 : (h->taper > 0 ? h->taper : TAPER_AUTO);
    int taper_levels = 0;
    while ((1 << (taper_levels + 1)) <= taper_div) ++taper_levels;
    const double taper_frac = (h->taper_pct > 0 ? h->taper_pct : TAPER_PCT_AUTO) / 100.0;
    long long cum = 0;
    segs.clear();
    seg_begin.assign((size_t)n_problems + 1, 0);
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        seg_begin[p] = (int)segs.size();
        if (pr.q_count <= 0 || pr.t_count <= 0) {
            // an empty problem still needs one work item: the CTA that completes a problem finalizes it
            // (writes its match count and the "no neighbour" rows of its knn table)
            Segment sg;
            sg.q_row0 = 0; sg.q_valid = 0; sg.q_local0 = 0; sg.out_row0 = pr.out_begin;
            sg.t_row0 = 0; sg.t_count = 0; sg.t_local0 = 0; sg.problem = p;
            segs.push_back(sg);
            seg_begin[p + 1] = (int)segs.size();
            continue;
        }
        // tapered tail: the problems that make up the last part of the batch are cut finer, so the CTAs that
        // finish the launch are short and the SMs run dry together (all CTAs of one launch cost the same, their
        // start times drift apart, and the finishing times of the last wave spread over one CTA duration)
        int Lp = L;
        if (taper_div > 1 && steps > 0) {
            const double f = (double)cum / (double)steps;            // where this problem starts in the batch
            const double tail0 = 1.0 - taper_frac;
            if (f >= tail0) {
                const int level = 1 + (int)((f - tail0) / taper_frac * taper_levels);
                Lp = std::max(MIN_SEG_ROWS, L >> std::min(level, taper_levels));
            }
        }
        if (gss) {
            for (int qb = 0; qb * bq < pr.q_count; ++qb) {
                int t0 = 0;
                while (t0 < pr.t_count) {
                    int cnt = (int)std::max<long long>(gss_min, std::min<long long>(GSS_MAX_ROWS, (steps - cum) / gss_div + 1));
                    if (pr.t_count - t0 - cnt < gss_min) cnt = pr.t_count - t0;   // no slivers
                    cnt = std::min(cnt, pr.t_count - t0);
                    Segment sg;
                    sg.q_row0 = pr.q_begin + qb * bq;
                    sg.q_valid = std::min(bq, pr.q_count - qb * bq);
                    sg.q_local0 = qb * bq;
                    sg.out_row0 = pr.out_begin + qb * bq;
                    sg.t_row0 = pr.t_begin + t0;
                    sg.t_count = cnt;
                    sg.t_local0 = t0;
                    sg.problem = p;
                    segs.push_back(sg);
                    t0 += cnt;
                    cum += cnt;
                }
            }
            seg_begin[p + 1] = (int)segs.size();
            continue;
        }
        cum += (long long)((pr.q_count + bq - 1) / bq) * pr.t_count;
        const int nsp = (pr.t_count + Lp - 1) / Lp;
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
                sg.problem = p;
                segs.push_back(sg);
                t0 += cnt;
            }
        }
        seg_begin[p + 1] = (int)segs.size();
    }
}

// Tensor form (bfm_tensor.cuh): work items of 256 query rows x a train range that is a multiple of 128 rows (the last
// range of a problem ends where the problem ends).  Every problem gets an 8-aligned base in the expanded planes
// (xq0 / xt0).  One CTA per SM walks the items with a fixed stride, so the cost of a cut is rounds x (tiles per item +
// ~3 tiles of per-item prologue); the cut with the lowest cost over the batch's shape histogram is taken.
void tensor_bases(const bfm_problem_t *problems, int n_problems, std::vector<int32_t> &xq0, std::vector<int32_t> &xt0, long long *rows) {
    xq0.resize(n_problems);
    xt0.resize(n_problems);
    long long q = 0, t = 0;
    for (int p = 0; p < n_problems; ++p) {
        xq0[p] = (int32_t)q;
        xt0[p] = (int32_t)t;
        q += (std::max(0, problems[p].q_count) + 7) & ~7;
        t += (std::max(0, problems[p].t_count) + 7) & ~7;
    }
    rows[0] = q;
    rows[1] = t;
}

void plan_items_tensor(const bfm_problem_t *problems, int n_problems, int n_sms, const std::vector<int32_t> &xq0, const std::vector<int32_t> &xt0,
                       std::vector<Segment> &segs, std::vector<int> &seg_begin, int *seg_rows_out) {
    std::vector<std::pair<std::pair<int, int>, long long>> shapes;   // (query blocks, train tiles) -> problems
    int max_tiles = 1;
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        if (pr.q_count <= 0 || pr.t_count <= 0) continue;
        const std::pair<int, int> key((pr.q_count + bfm::TC_BQ - 1) / bfm::TC_BQ, (pr.t_count + bfm::TC_BT - 1) / bfm::TC_BT);
        max_tiles = std::max(max_tiles, key.second);
        if (!shapes.empty() && shapes.back().first == key) { ++shapes.back().second; continue; }
        bool found = false;
        for (size_t i = 0; i < shapes.size() && i < 16 && !found; ++i)
            if (shapes[i].first == key) { ++shapes[i].second; found = true; }
        if (!found) shapes.emplace_back(key, 1);
    }
    const int cap = 8192;   // tiles per item: the keys hold a 21-bit column offset
    int best_L = std::min(max_tiles, cap);
    double best_cost = -1.0;
    const int n_cand = std::min(std::min(max_tiles, cap), 512);
    for (int c = 0; c <= n_cand; ++c) {
        const int L = c == 0 ? std::min(max_tiles, cap) : c;
        long long items = 0;
        int seg_tiles = 1;
        for (const auto &sh : shapes) {
            const int nsp = (sh.first.second + L - 1) / L;
            items += sh.second * sh.first.first * nsp;
            seg_tiles = std::max(seg_tiles, (sh.first.second + nsp - 1) / nsp);
        }
        const long long rounds = (items + n_sms - 1) / std::max(n_sms, 1);
        const double cost = (double)rounds * (seg_tiles + 3.0);
        if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best_L = L; }
    }
    *seg_rows_out = best_L * bfm::TC_BT;
    segs.clear();
    seg_begin.assign((size_t)n_problems + 1, 0);
    for (int p = 0; p < n_problems; ++p) {
        const bfm_problem_t &pr = problems[p];
        seg_begin[p] = (int)segs.size();
        if (pr.q_count > 0 && pr.t_count > 0) {
            const int tiles = (pr.t_count + bfm::TC_BT - 1) / bfm::TC_BT;
            const int nsp = (tiles + best_L - 1) / best_L;
            const int base = tiles / nsp, rem = tiles % nsp;
            for (int qb = 0; qb * bfm::TC_BQ < pr.q_count; ++qb) {
                int t0 = 0;
                for (int sidx = 0; sidx < nsp; ++sidx) {
                    const int cnt = std::min((base + (sidx < rem ? 1 : 0)) * bfm::TC_BT, pr.t_count - t0);
                    Segment sg;
                    sg.q_row0 = xq0[p] + qb * bfm::TC_BQ;
                    sg.q_valid = std::min(bfm::TC_BQ, pr.q_count - qb * bfm::TC_BQ);
                    sg.q_local0 = qb * bfm::TC_BQ;
                    sg.out_row0 = pr.out_begin + qb * bfm::TC_BQ;
                    sg.t_row0 = xt0[p] + t0;
                    sg.t_count = cnt;
                    sg.t_local0 = t0;
                    sg.problem = p;
                    segs.push_back(sg);
                    t0 += cnt;
                }
            }
        }
        seg_begin[p + 1] = (int)segs.size();
    }
}

// The persistent form (resident inputs): at most ONE wave of CTAs; CTA c starts with work item c of plan_segments and
// draws further items from a ticket counter; when the queue is empty the problems are finalized tile by tile (FT rows
// of one problem per tile).  The tiles, in (problem, tile) order, are dealt in contiguous runs to the LOWEST-indexed
// quarter of the grid (one tile each while they last): a tile waits for the tiles before it - CTAs with a lower or
// equal index, dispatched no later than its own - and for its problem's work items, which may include the first items
// of CTAs that are not resident yet when the GPU is shared with another kernel.  Those CTAs start as soon as a slot is
// free, and slots do come free: every resident CTA that owns no tile exits when the queue is empty, and the owners
// are at most a quarter of the grid, so even three such launches sharing the GPU cannot fill it with waiting owners
// (tools/share_gpu_probe.py runs two of them against each other).  Every wait also has a 2 s time-out.
constexpr int FT_ROWS_TILE = NT * bfm::FT_RPT;

void plan_tiles(const bfm_problem_t *problems, int n_problems, int n_ctas, std::vector<bfm::FinTile> &tiles,
                std::vector<int2> &cta_tiles) {
    tiles.clear();
    int slot = 0;
    for (int p = 0; p < n_problems; ++p) {
        const int rows = std::max(0, problems[p].q_count);
        const int nt = std::max(1, (rows + FT_ROWS_TILE - 1) / FT_ROWS_TILE);
        for (int j = 0; j < nt; ++j) {
            bfm::FinTile t;
            t.problem = p; t.row0 = j * FT_ROWS_TILE; t.index = j; t.n_tiles = nt; t.slot0 = slot;
            t.pad[0] = t.pad[1] = t.pad[2] = 0;
            tiles.push_back(t);
        }
        slot += nt;
    }
    const long long n = (long long)tiles.size();
    const long long owners = std::min<long long>(n, std::max<long long>(1, n_ctas / 4));
    cta_tiles.assign((size_t)n_ctas, make_int2(0, 0));
    for (long long i = 0; i < n; ++i) {
        int2 &w = cta_tiles[(size_t)(i * owners / n)];   // contiguous runs, owner index non-decreasing in tile order
        if (w.y == 0) w.x = (int)i;
        ++w.y;
    }
}

int check_opts(bfm_handle_t h, const bfm_options_t *o, int n_problems) {
    if (!o) return fail(h, BFM_ERR_INVALID, "options is NULL");
    if (o->k < 1) return fail(h, BFM_ERR_INVALID, "k must be >= 1");
    if (o->k > BFM_MAX_K) return fail(h, BFM_ERR_UNSUPPORTED, "k > 16 is not supported by this build");
    if (o->k > 2 && o->ratio >= 0) return fail(h, BFM_ERR_INVALID, "the ratio test is defined on k == 2");
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

// Input gate of the pipelined host path (see run_host): the kernel's CTAs wait on these watermarks.
struct Gate {
    const unsigned long long *ready = nullptr;  // device: [0] query rows landed, [1] train rows landed (+ base)
    unsigned long long base = 0;
    uint32_t *status = nullptr;                 // pinned host word, device-visible
    // SM-fed variant (pinned caller arrays): the first n_feed CTAs of the launch do the upload themselves
    int n_feed = 0, rounds = 0, q_rows = 0, t_rows = 0;
    const void *src[4] = {nullptr, nullptr, nullptr, nullptr};
    void *dst[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t bytes[4] = {0, 0, 0, 0};
    uint32_t *prog = nullptr;
    uint32_t epoch = 0;
    const uint32_t *host_ready = nullptr;   // pinned word: rounds the host has staged so far (NULL: all staged)
};

// The device path: every data pointer is device-visible, work is queued on `st`.  ONE kernel launch.
int run_device(bfm_handle_t h, const uint8_t *q, int32_t nq_rows, const uint8_t *t, int32_t nt_rows,
               const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
               const bfm_options_t *o, const bfm_outputs_t *dests, int n_dests, cudaStream_t st,
               const Gate *gate = nullptr, const int32_t *t_limit = nullptr, int32_t t_plan_rows = 0) {
    h->info = bfm_launch_info_t{};
    if (n_problems <= 0) return BFM_OK;
    if (n_out_rows <= 0) {
        // every query set is empty: no kernel runs, but m_count is an output per problem and must read 0
        for (int d = 0; d < n_dests && dests; ++d)
            if (dests[d].m_count && !dests[d].multicast) CU_TRY(h, cudaMemsetAsync(dests[d].m_count, 0, (size_t)n_problems * 4, st));
        return BFM_OK;
    }
    // stream hand-over: the workspace and the device tables are shared by all calls on this handle
    if (h->last_pending && h->last_stream != st) {
        // (work left on the handle's own stream is marked lazily - a host call synchronises before it returns, and a
        // tracking call costs no event when nobody follows on another stream; caller streams get their event eagerly,
        // they may not exist any more when the next call arrives)
        if (h->last_stream == h->stream) CU_TRY(h, cudaEventRecord(h->last_ev, h->stream));
        CU_TRY(h, cudaStreamWaitEvent(st, h->last_ev, 0));
    }
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(t)) & 15)
        return fail(h, BFM_ERR_INVALID, "descriptor arrays must be 16-byte aligned");
    if (n_dests < 1 || n_dests > bfm::MAX_DEST || !dests)
        return fail(h, BFM_ERR_INVALID, "between 1 and 8 result destinations are supported");
    for (int d = 0; d < n_dests; ++d) {
        const bfm_outputs_t &od = dests[d];
        if (od.m_count && (!od.m_query || !od.m_train || !od.m_dist))
            return fail(h, BFM_ERR_INVALID, "m_query/m_train/m_dist/m_count must be given together");
        if ((od.knn_idx == nullptr) != (od.knn_dist == nullptr))
            return fail(h, BFM_ERR_INVALID, "knn_idx and knn_dist must be given together");
        if (od.knn_idx && o->k > 1 && ((reinterpret_cast<uintptr_t>(od.knn_idx) | reinterpret_cast<uintptr_t>(od.knn_dist)) & 7))
            return fail(h, BFM_ERR_INVALID, "knn_idx / knn_dist must be 8-byte aligned");
    }

    // projection-window search over a binned train set (bfm_window.cuh): finite radius, every problem's train set
    // small enough for the cell tables; a batch of window problems takes it too (blockIdx.y = problem)
    bool binned = o->mask_kind == BFM_MASK_WINDOW && (o->k + 1) / 2 == 1 && h->window_bins != 1 && !gate && n_problems <= 4096 &&
                  std::isfinite(o->window_radius) && o->window_radius > 0.0f && (t_limit == nullptr || n_problems == 1);
    int bin_max_q = 0, bin_max_t = 0;
    for (int p = 0; p < n_problems && binned; ++p) {
        bin_max_q = std::max(bin_max_q, problems[p].q_count);
        bin_max_t = std::max(bin_max_t, problems[p].t_count);
        if (problems[p].t_count > bfm::WB_MAX_ROWS) binned = false;
    }
    if (bin_max_q <= 0 || bin_max_t <= 0) binned = false;
    const int bin_grid = binned ? (bin_max_q + bfm::WS_NT / 32 - 1) / (bfm::WS_NT / 32) : 0;

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
        d.n_segs = 0;                // filled from the plan below
        d.pad = 0;
        h->probs_host[p] = d;
        if (o->cross_check || binned) col_rows += pr.t_count;   // (binned: also the row offset of the problem's binned rows)
    }
    if (col_rows >= (1ll << 31)) return fail(h, BFM_ERR_INVALID, "batch too large for cross-check");

    // -- choose the kernel variant ----------------------------------------------------------------
    const int mode = o->cross_check ? 1 : ((o->k >= 2 || o->ratio >= 0) ? 2 : 0);
    const int mask = o->mask_kind;
    const int passes = (o->k + 1) / 2;  // k > 2: two more neighbours per pass (see ScanParams::lower)
    const int pm = (h->popc_mode && passes == 1) ? h->popc_mode : 40;  // measured best for every mode: profiles/sweep_r01.json
    int r = passes > 1 ? 1 : h->qpt;
    int slots = 0, seg_rows = 0;
    // a device-side train count: plan the work items for the rows the caller expects to exist (the kernel re-cuts
    // whatever does exist evenly over them, so the guess only affects efficiency, never the result)
    bfm_problem_t hinted;
    const bfm_problem_t *plan_problems = problems;
    if (t_limit != nullptr && n_problems == 1 && !binned) {
        hinted = problems[0];
        const int guess = t_plan_rows > 0 ? t_plan_rows : std::max(1, problems[0].t_count / 2);
        hinted.t_count = std::min(problems[0].t_count, std::max(128, (guess + 1023) & ~1023));
        plan_problems = &hinted;
    }
    // The static form wins on long launches (the hardware hands a freed CTA slot to the next work item, and a CTA that
    // has waited longest is served first, so the items of a large batch drain in order), the persistent form on short
    // ones (one wave, no second and third wave of CTA launches, tile-parallel finalize): profiles/r02_kernel_forms.md
    long long total_pairs = 0;
    for (int p = 0; p < n_problems; ++p) total_pairs += (long long)std::max(0, problems[p].q_count) * std::max(0, problems[p].t_count);
    // Tensor form (bfm_tensor.cuh): distances as s8 dot products on tcgen05, 3x the POPC kernel on the headline batch.
    // Resident inputs, one pass (k <= 2), no mask, no cross-check, a train count known on the host.
    const bool tensor = !gate && !binned && mode != 1 && mask == BFM_MASK_NONE && passes == 1 && t_limit == nullptr && h->tensor != 1 &&
                        (h->tensor == 2 || total_pairs >= TENSOR_MIN_PAIRS);
    const bool persistent = !tensor && !gate && !binned && (h->persistent == 2 || (h->persistent == 0 && total_pairs <= PERSISTENT_MAX_PAIRS));
    // static form with a few large problems: the CTA that completes a problem would walk all its rows alone (64 tiles
    // x ~5 us at 64k rows, fully exposed when there is nothing else to scan) - finalize tiles in a second launch instead
    bool defer = false;
    if (!persistent && !binned && !gate && n_problems <= 16) {
        int max_rows = 0;
        for (int p = 0; p < n_problems; ++p) max_rows = std::max(max_rows, problems[p].q_count);
        defer = max_rows >= 4096;
    }
    if (tensor) defer = true;   // the tensor scan only reduces; the tile-parallel kernel finalizes
    if (binned) r = 1;
    if (r != 1 && r != 2 && r != 4) {
        // largest register tile that still leaves >= 2 work items per CTA slot
        for (int cand : {4, 2, 1}) {
            int occ = 0;
            int rc = occupancy(h, cand, mode, mask, pm, &occ);
            if (rc) return rc;
            long long units = 0;
            for (int p = 0; p < n_problems; ++p)
                if (plan_problems[p].q_count > 0 && plan_problems[p].t_count > 0)
                    units += (long long)((plan_problems[p].q_count + NT * cand - 1) / (NT * cand)) *
                             ((plan_problems[p].t_count + 2 * MIN_SEG_ROWS - 1) / (2 * MIN_SEG_ROWS));
            r = cand;
            // (four queries per thread pay off from ~40 keyframe pairs: 32 pairs 153 vs 149 us, 20 pairs 105 vs 101 us
            // with two - shorter work items, shorter tail; profiles/r02_kernel_forms.md)
            if (units >= (cand == 4 && persistent ? 4ll : 2ll) * occ * h->sm_count) break;
        }
    }
    {
        int occ = 0;
        int rc = occupancy(h, r, mode, mask, pm, &occ);
        if (rc) return rc;
        slots = occ * h->sm_count;
        if (persistent && h->ctas_per_sm > 0) slots = std::min(occ, h->ctas_per_sm) * h->sm_count;
    }
    // -- plan cache: same problems + same variant as the previous call -> the device tables are
    //    already in place (steady state of a tracking loop with fixed shapes, bench loops) ----------
    // resident inputs take the persistent form (at most one wave, tile-parallel finalize inside the same launch)
    const int plan_sig[8] = {n_problems, binned ? 100 : (tensor ? 200 : r), mode, h->segment_rows, h->waves + 4096 * (plan_problems == &hinted ? hinted.t_count : 0), slots,
                             h->taper * 1000 + h->taper_pct + 100000 * h->gss_div + 10000000 * h->gss_min, (persistent ? 1 : 0) + (defer ? 2 : 0)};
    const bool plan_hit = h->plan_valid && std::memcmp(plan_sig, h->plan_sig, sizeof(plan_sig)) == 0 &&
                          h->plan_problems.size() == (size_t)n_problems &&
                          std::memcmp(h->plan_problems.data(), problems, sizeof(bfm_problem_t) * (size_t)n_problems) == 0;
    if (!plan_hit) {
        // output rows belong to exactly one problem (checked once per problem table: a cached plan was checked before)
        std::vector<std::pair<int32_t, int32_t>> spans;
        spans.reserve(n_problems);
        for (int p = 0; p < n_problems; ++p)
            if (problems[p].q_count > 0) spans.emplace_back(problems[p].out_begin, problems[p].q_count);
        std::sort(spans.begin(), spans.end());
        for (size_t i = 1; i < spans.size(); ++i)
            if ((long long)spans[i - 1].first + spans[i - 1].second > spans[i].first)
                return fail(h, BFM_ERR_INVALID, "two problems share output rows (out_begin ranges overlap)");
    }
    // same shapes, other rows (a loop-closing step with a fresh candidate list: every keyframe has its 2000 descriptors,
    // only their places in the bank differ): the cut of the cached plan still holds, its work items are re-based on
    // the new rows instead of planned again (256 pairs: ~0.15 ms of planning -> ~10 us)
    bool shape_hit = false;
    if (!plan_hit && !binned && h->plan_valid && plan_problems == problems && std::memcmp(plan_sig, h->plan_sig, sizeof(plan_sig)) == 0 &&
        h->plan_problems.size() == (size_t)n_problems && !h->segs_host.empty()) {
        shape_hit = true;
        for (int p = 0; p < n_problems && shape_hit; ++p)
            shape_hit = problems[p].q_count == h->plan_problems[p].q_count && problems[p].t_count == h->plan_problems[p].t_count;
        if (shape_hit) {
            for (Segment &sg : h->segs_host) {
                const bfm_problem_t &pr = problems[sg.problem];
                const bool placeholder = sg.q_valid == 0 && sg.t_count == 0;   // the one item of an empty problem
                sg.out_row0 = pr.out_begin + sg.q_local0;
                if (tensor) continue;   // (its rows are rows of the expanded planes: a function of the shapes alone)
                sg.q_row0 = placeholder ? 0 : pr.q_begin + sg.q_local0;
                sg.t_row0 = placeholder ? 0 : pr.t_begin + sg.t_local0;
            }
            for (int p = 0; p < n_problems; ++p) {
                const int n_segs = h->plan_probs[p].n_segs;
                const int32_t xt0 = h->plan_probs[p].col0, xq0 = h->plan_probs[p].pad;
                h->plan_probs[p] = h->probs_host[p];
                h->plan_probs[p].n_segs = n_segs;
                if (tensor) { h->plan_probs[p].col0 = xt0; h->plan_probs[p].pad = xq0; }
            }
        }
    }
    if (shape_hit) {
        // (nothing to plan)
    } else if (!plan_hit && binned) {
        h->segs_host.clear();
        h->seg_begin.assign(2, 0);
        h->plan_seg_rows = 0;
        h->plan_probs = h->probs_host;
        for (int p = 0; p < n_problems; ++p) h->plan_probs[p].n_segs = bin_grid;   // the search kernel's CTAs play the role of segments
    } else if (!plan_hit) {
        // (a device-side train count re-cuts the rows that exist over equal items per query block: no guided lengths)
        std::vector<int32_t> xq0, xt0;
        if (tensor) {
            tensor_bases(problems, n_problems, xq0, xt0, h->x_rows);
            plan_items_tensor(problems, n_problems, h->sm_count, xq0, xt0, h->segs_host, h->seg_begin, &seg_rows);
        } else {
            plan_segments(h, plan_problems, n_problems, r, slots, h->segs_host, h->seg_begin, &seg_rows, persistent);
        }
        h->plan_seg_rows = seg_rows;
        h->plan_probs = h->probs_host;
        for (int p = 0; p < n_problems; ++p) h->plan_probs[p].n_segs = h->seg_begin[p + 1] - h->seg_begin[p];
        if (tensor)
            for (int p = 0; p < n_problems; ++p) { h->plan_probs[p].col0 = xt0[p]; h->plan_probs[p].pad = xq0[p]; }
        if (defer) {
            // (static form, a few large problems: one finalize tile per CTA of a second launch)
            long long n_tiles = 0;
            for (int p = 0; p < n_problems; ++p) n_tiles += std::max(1, (std::max(0, problems[p].q_count) + FT_ROWS_TILE - 1) / FT_ROWS_TILE);
            h->plan_ctas = (int)n_tiles;
            plan_tiles(problems, n_problems, h->plan_ctas, h->fin_tiles_host, h->cta_tiles_host);
            for (int p = 0; p < n_problems; ++p) h->plan_probs[p].n_segs = 0;   // the tiles do not wait: the scan has completed
        }
        if (persistent) {
            long long n_tiles = 0;
            for (int p = 0; p < n_problems; ++p) n_tiles += std::max(1, (std::max(0, problems[p].q_count) + FT_ROWS_TILE - 1) / FT_ROWS_TILE);
            h->plan_ctas = (int)std::min<long long>(slots, std::max<long long>((long long)h->segs_host.size(), std::min<long long>(n_tiles, slots)));
            h->plan_ctas = std::max(h->plan_ctas, 1);
            plan_tiles(problems, n_problems, h->plan_ctas, h->fin_tiles_host, h->cta_tiles_host);
        }
    }
    seg_rows = h->plan_seg_rows;
    const size_t n_segs = h->segs_host.size();
    const bool tiles = persistent || defer;
    const size_t n_ctas_p = tiles ? h->cta_tiles_host.size() : 0, n_tiles_p = tiles ? h->fin_tiles_host.size() : 0;

    // -- workspace: [row state u64 | column keys u32 | done counters u32], all-ones when idle.  It is
    //    self-cleaning (the finalizing CTA restores every slot it read), so a memset is only queued
    //    after (re)allocation or after a call that failed half-way --------------------------------------
    const size_t state_bytes = (size_t)n_out_rows * 8;
    const size_t col_bytes = (size_t)col_rows * 4;
    const size_t done_bytes = (size_t)n_problems * 4 * 2;   // done counters + finalized-tile counters (persistent form)
    const unsigned state_gen = h->state.generation;
    int rc = ensure(h, h->state, state_bytes + col_bytes + done_bytes);
    if (rc) return rc;
    if (h->state.generation != state_gen) h->state_clean = false;
    unsigned long long *rowstate = static_cast<unsigned long long *>(h->state.p);
    uint32_t *colkeys = reinterpret_cast<uint32_t *>(static_cast<char *>(h->state.p) + state_bytes);
    uint32_t *done = reinterpret_cast<uint32_t *>(static_cast<char *>(h->state.p) + state_bytes + col_bytes);
    uint32_t *fin_done = done + n_problems;
    if (tiles) {
        const unsigned finc_gen = h->finc.generation;
        rc = ensure(h, h->finc, n_tiles_p * 4);
        if (rc) return rc;
        if (h->finc.generation != finc_gen || h->fin_epoch + (uint32_t)passes >= (1u << 20)) {
            CU_TRY(h, cudaMemsetAsync(h->finc.p, 0, h->finc.cap, st));
            h->fin_epoch = 0;
        }
    }
    if (passes > 1) {
        rc = ensure(h, h->lower, (size_t)n_out_rows * 4);
        if (rc) return rc;
    }
    unsigned long long x_plane[2] = {0, 0};
    if (tensor) {
        // expanded planes; the slack rows (and whatever earlier calls left behind) are only ever multiplied into
        // columns and rows nobody reads, but they must exist - and be zero once, so that they hold s8 values
        DevBuf *xb[2] = {&h->xq, &h->xt};
        for (int a = 0; a < 2; ++a) {
            x_plane[a] = (unsigned long long)(h->x_rows[a] + bfm::TC_SLACK_ROWS) * 128ull;
            const unsigned gen = xb[a]->generation;
            rc = ensure(h, *xb[a], (size_t)(2 * x_plane[a]));
            if (rc) return rc;
            if (xb[a]->generation != gen) CU_TRY(h, cudaMemsetAsync(xb[a]->p, 0, xb[a]->cap, st));
        }
    }
    const size_t prob_bytes = (size_t)n_problems * sizeof(Problem);
    const size_t seg_bytes = n_segs * sizeof(Segment), work_bytes = (n_ctas_p * sizeof(int2) + 15) & ~(size_t)15;
    const size_t table_bytes = prob_bytes + seg_bytes + work_bytes + n_tiles_p * sizeof(bfm::FinTile);
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
        std::memcpy(h->h_tables[slot], h->plan_probs.data(), prob_bytes);
        std::memcpy(static_cast<char *>(h->h_tables[slot]) + prob_bytes, h->segs_host.data(), seg_bytes);
        if (tiles) {
            std::memcpy(static_cast<char *>(h->h_tables[slot]) + prob_bytes + seg_bytes, h->cta_tiles_host.data(), n_ctas_p * sizeof(int2));
            std::memcpy(static_cast<char *>(h->h_tables[slot]) + prob_bytes + seg_bytes + work_bytes, h->fin_tiles_host.data(),
                        n_tiles_p * sizeof(bfm::FinTile));
        }
        // NOTE: the device table is shared by consecutive calls on one handle; stream order keeps the
        // upload of call n+1 behind the kernel of call n when both use the same stream (the contract).
        if (h->tables_by_kernel && table_bytes <= ((size_t)1 << 20)) {
            // (chunked host path: a copy-engine upload on `st` would queue behind the descriptor chunks of `in_stream` -
            // the engine serves that stream while it has copies ready - so a few threads fetch the pinned table instead)
            upload_tables_kernel<<<8, 256, 0, st>>>(static_cast<const uint4 *>(h->h_tables[slot]), static_cast<uint4 *>(h->tables.p), (table_bytes + 15) / 16);
            CU_TRY(h, cudaGetLastError());
        } else {
            CU_TRY(h, cudaMemcpyAsync(h->tables.p, h->h_tables[slot], table_bytes, cudaMemcpyHostToDevice, st));
        }
        CU_TRY(h, cudaEventRecord(h->table_ev[slot], st));
        std::memcpy(h->plan_sig, plan_sig, sizeof(plan_sig));
        h->plan_problems.assign(problems, problems + n_problems);
        h->plan_valid = true;
    }
    const Problem *d_probs = static_cast<const Problem *>(h->tables.p);
    const Segment *d_segs = reinterpret_cast<const Segment *>(static_cast<char *>(h->tables.p) + prob_bytes);

    if (!h->state_clean) {
        CU_TRY(h, cudaMemsetAsync(h->state.p, 0xFF, h->state.cap, st));
        CU_TRY(h, cudaMemsetAsync(h->d_queue, 0, 256, st));
    }
    h->state_clean = false;  // set again once the kernel (which restores the state) is queued

    ScanParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.q = reinterpret_cast<const uint4 *>(q);
    sp.t = reinterpret_cast<const uint4 *>(t);
    sp.segs = d_segs;
    sp.problems = d_probs;
    sp.rowstate = rowstate;
    sp.colkeys = colkeys;
    sp.done = done;
    if (tiles) {
        sp.cta_tiles = reinterpret_cast<const int2 *>(static_cast<char *>(h->tables.p) + prob_bytes + seg_bytes);
        sp.n_items = (int32_t)n_segs;
        sp.n_ctas = (int32_t)n_ctas_p;
        sp.fin_tiles = reinterpret_cast<const bfm::FinTile *>(static_cast<char *>(h->tables.p) + prob_bytes + seg_bytes + work_bytes);
        sp.fin_done = fin_done;
        sp.fin_count = static_cast<uint32_t *>(h->finc.p);
    }
    sp.t_limit = t_limit;
    if (t_limit != nullptr && !binned) {
        const int nqb = (problems[0].q_count + NT * r - 1) / (NT * r);
        sp.limit_segs = std::max(1, (int)(n_segs / std::max(nqb, 1)));
    }
    if (gate) {
        sp.ready = gate->ready;
        sp.ready_base = gate->base;
        sp.status = gate->status;
    }
    const int n_feed = (gate && gate->n_feed > 0) ? gate->n_feed : 0;
    if (n_feed) {
        sp.n_feed = n_feed;
        sp.feed_rounds = gate->rounds;
        sp.feed_q_rows = gate->q_rows;
        sp.feed_t_rows = gate->t_rows;
        for (int a = 0; a < 4; ++a) {
            sp.feed_src[a] = static_cast<const uint4 *>(gate->src[a]);
            sp.feed_dst[a] = static_cast<uint4 *>(gate->dst[a]);
            sp.feed_bytes[a] = gate->bytes[a];
        }
        sp.feed_prog = gate->prog;
        sp.feed_epoch = gate->epoch;
        sp.feed_stall = h->test_stall;
        sp.feed_host_ready = gate->host_ready;
    }
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
    sp.k = o->k;
    sp.cross_check = o->cross_check;
    sp.max_distance = o->max_distance;
    sp.use_ratio = o->ratio >= 0;
    sp.ratio = o->ratio;
    sp.n_dest = n_dests;
    for (int d = 0; d < n_dests; ++d) {
        sp.dest[d].knn_idx = dests[d].knn_idx;
        sp.dest[d].knn_dist = dests[d].knn_dist;
        sp.dest[d].m_query = dests[d].m_query;
        sp.dest[d].m_train = dests[d].m_train;
        sp.dest[d].m_dist = dests[d].m_dist;
        sp.dest[d].m_count = dests[d].m_count;
        if (dests[d].multicast) sp.dest_multicast |= 1u << d;
    }
    if (h->timing) CU_TRY(h, cudaEventRecord(h->ev[0], st));
    if (binned) {
        const size_t T = (size_t)col_rows, P = (size_t)n_problems;
        // layout: [counters per problem + tickets | cell offsets per problem | rows in cell order: desc, xy, orig | cell, rank per row]
        const size_t o_cnt = 0, o_cs = align256((P * bfm::WB_CELLS + P) * 4), o_desc = align256(o_cs + P * (bfm::WB_CELLS + 1) * 4),
                     o_xy = align256(o_desc + T * 32), o_orig = align256(o_xy + T * 8), o_cell = align256(o_orig + T * 4),
                     o_rank = align256(o_cell + T * 4), total = align256(o_rank + T * 4);
        const unsigned bins_gen = h->bins.generation;
        rc = ensure(h, h->bins, total);
        if (rc) return rc;
        char *bb = static_cast<char *>(h->bins.p);
        // the counters (and the tickets behind them) are self-cleaning: zeroed once per allocation and when the number of
        // problems - the layout of the block - changes
        if (h->bins.generation != bins_gen || h->bins_problems != n_problems) CU_TRY(h, cudaMemsetAsync(h->bins.p, 0, o_cs, st));
        h->bins_problems = n_problems;
        bfm::BinView bv{reinterpret_cast<uint4 *>(bb + o_desc), reinterpret_cast<float2 *>(bb + o_xy),
                        reinterpret_cast<int32_t *>(bb + o_orig), reinterpret_cast<int32_t *>(bb + o_cs)};
        int32_t *d_cnt = reinterpret_cast<int32_t *>(bb + o_cnt);
        uint32_t *d_ticket = reinterpret_cast<uint32_t *>(d_cnt + P * bfm::WB_CELLS);
        int32_t *d_cell = reinterpret_cast<int32_t *>(bb + o_cell), *d_rank = reinterpret_cast<int32_t *>(bb + o_rank);
        const double cell = 2.0 * (double)o->window_radius * (1.0 + 1e-6) + 1e-30;
        const double inv_cell = 1.0 / cell;
        const dim3 bgrid((unsigned)((bin_max_t + 255) / 256), (unsigned)n_problems);
        bfm::wb_count_kernel<<<bgrid, 256, 0, st>>>(sp.t_xy, d_probs, t_limit, inv_cell, d_cnt, d_ticket, d_cell, d_rank, bv.cell_start);
        CU_TRY(h, cudaGetLastError());
        bfm::wb_scatter_kernel<<<bgrid, 256, 0, st>>>(sp.t, sp.t_xy, d_probs, t_limit, d_cell, d_rank, bv);
        CU_TRY(h, cudaGetLastError());
        ScanParams sq = sp;
        sq.knn_col0 = 0;
        sq.knn_cols = std::min(2, o->k);
        const dim3 sgrid((unsigned)bin_grid, (unsigned)n_problems);
        if (mode == 1) bfm::wb_search_kernel<1, true><<<sgrid, bfm::WS_NT, 0, st>>>(sq, bv, inv_cell);
        else if (mode == 2) bfm::wb_search_kernel<2, false><<<sgrid, bfm::WS_NT, 0, st>>>(sq, bv, inv_cell);
        else bfm::wb_search_kernel<1, false><<<sgrid, bfm::WS_NT, 0, st>>>(sq, bv, inv_cell);
        CU_TRY(h, cudaGetLastError());
    }
    if (gate && passes > 1 && !binned) {
        // CUDA loads a kernel when it is first launched, and that load can wait for the device to go idle.  The first pass
        // of a gated call spins until the host has queued the rest of the upload - which it only does after ALL passes
        // have been launched - so a later pass whose kernel is not loaded yet would stall the host for the gate's whole
        // time-out (tools/repro_gate.py).  Load every pass's kernel before the first launch.
        for (int pass = 0; pass < passes; ++pass) {
            cudaFuncAttributes attr;
            CU_TRY(h, cudaFuncGetAttributes(&attr, reinterpret_cast<const void *>(pick_scan(r, mode, mask, pm, pass > 0, persistent))));
        }
    }
    for (int pass = 0; pass < (binned ? 0 : passes); ++pass) {
        sp.knn_col0 = 2 * pass;
        sp.knn_cols = std::min(2, o->k - 2 * pass);
        sp.lower = pass > 0 ? static_cast<const uint32_t *>(h->lower.p) : nullptr;
        sp.lower_out = pass + 1 < passes ? static_cast<uint32_t *>(h->lower.p) : nullptr;
        if (pass > 0)  // the match list (gate on the nearest neighbour) was produced by the first pass
            for (int d = 0; d < n_dests; ++d) sp.dest[d].m_count = nullptr;
        const ScanFn fn = pick_scan(r, mode, mask, pm, pass > 0, persistent);
        if (pass > 0) sp.n_feed = 0;   // the inputs are resident after the first pass
        sp.trace = pass == 0 ? h->trace : nullptr;
        sp.trace_cap = h->trace_cap;
        sp.trace_base = 0;
        if (tiles) sp.fin_epoch = ++h->fin_epoch;
        if (persistent) {
            sp.queue = h->d_queue + 32 * h->queue_phase;          // (separate 128-byte lines)
            sp.queue_other = h->d_queue + 32 * (h->queue_phase ^ 1);
            h->queue_phase ^= 1;
        }
        sp.defer_finalize = defer ? 1 : 0;
        const unsigned grid = persistent ? (unsigned)n_ctas_p : (unsigned)(n_segs + (pass == 0 ? n_feed : 0));
        if (tensor) {
            bfm::TensorLaunch tl;
            tl.q = q; tl.t = t;
            tl.probs = d_probs;
            tl.n_problems = n_problems;
            int max_rows = 0;
            for (int p = 0; p < n_problems; ++p) max_rows = std::max(max_rows, std::max(problems[p].q_count, problems[p].t_count));
            tl.max_rows = max_rows;
            tl.xq = h->xq.p; tl.xt = h->xt.p;
            tl.xq_plane = x_plane[0]; tl.xt_plane = x_plane[1];
            tl.items = d_segs;
            tl.n_items = (int)n_segs;
            tl.grid = (int)std::min<size_t>(n_segs, (size_t)h->sm_count);
            tl.rowstate = rowstate;
            tl.status = h->h_status;
            tl.ev_scan[0] = h->timing ? h->ev[2] : nullptr;
            tl.ev_scan[1] = h->timing ? h->ev[3] : nullptr;
            const int trc = bfm::tensor_launch(tl, st);
            if (trc) return fail(h, BFM_ERR_CUDA, std::string("tensor form launch failed: ") + cudaGetErrorString((cudaError_t)trc));
        } else {
            fn<<<grid, NT, 0, st>>>(sp);
            CU_TRY(h, cudaGetLastError());
        }
        if (defer) {
            bfm::bfm_tiles_kernel<<<(unsigned)n_tiles_p, NT, 0, st>>>(sp);
            CU_TRY(h, cudaGetLastError());
        }
    }
    if (h->timing) CU_TRY(h, cudaEventRecord(h->ev[1], st));
    h->state_clean = true;  // every slot touched is restored by the CTA that finalizes its problem
    if (st != h->stream) CU_TRY(h, cudaEventRecord(h->last_ev, st));
    h->last_stream = st;
    h->last_pending = true;

    h->launches += binned ? 3 : passes * (defer ? 2 : 1) + (tensor ? 1 : 0);
    h->info.kernels_launched = binned ? 3 : passes * (defer ? 2 : 1) + (tensor ? 1 : 0);
    h->info.scan_grid = binned ? bin_grid : (tensor ? (int32_t)std::min<size_t>(n_segs, (size_t)h->sm_count) : persistent ? (int32_t)n_ctas_p : (int32_t)n_segs);
    h->info.scan_block = tensor ? bfm::TC_THREADS : NT;
    h->info.queries_per_thread = tensor ? 0 : r;   // (tensor form: a thread of the epilogue owns one query row)
    h->info.popc_mode = tensor ? 0 : pm;           // 0: no POPC at all - distances come from tcgen05.mma
    h->info.segments = (int32_t)n_segs;
    h->info.train_rows_per_segment = seg_rows;
    if (h->timing && !gate) {
        CU_TRY(h, cudaEventSynchronize(h->ev[1]));
        CU_TRY(h, cudaEventElapsedTime(&h->info.scan_ms, h->ev[0], h->ev[1]));
        h->info.total_ms = h->info.scan_ms;
        if (tensor) CU_TRY(h, cudaEventElapsedTime(&h->info.scan_ms, h->ev[2], h->ev[3]));   // the scan alone; total_ms: expansion + scan + finalize
    }
    if (h->check_clean && !gate) {  // debugging aid: the workspace must be all-ones after every call
        CU_TRY(h, cudaDeviceSynchronize());
        std::vector<unsigned char> host(h->state.cap);
        CU_TRY(h, cudaMemcpy(host.data(), h->state.p, h->state.cap, cudaMemcpyDeviceToHost));
        size_t bad = 0, first = 0;
        for (size_t i = 0; i < host.size(); ++i)
            if (host[i] != 0xFF) { if (!bad) first = i; ++bad; }
        if (bad)
            std::fprintf(stderr, "[bfm check] workspace dirty after call: %zu bytes, first at %zu (rows=%d col_rows=%lld state_bytes=%zu cap=%zu mode=%d P=%d)\n",
                         bad, first, n_out_rows, col_rows, state_bytes, h->state.cap, mode, n_problems);
    }
    return BFM_OK;
}


#include "bfm_pipeline.cuh"
#include "bfm_localmap.cuh"

}  // namespace

#include "bfm_localmap_host.cuh"

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
    h->check_clean = std::getenv("BFM_CHECK_CLEAN") != nullptr;
    std::memset(h->occ_cache, 0, sizeof(h->occ_cache));
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->last_ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_prog, 256) == cudaSuccess && cudaMemset(h->d_prog, 0, 256) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_queue, 256) == cudaSuccess && cudaMemset(h->d_queue, 0, 256) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_ready, 256) == cudaSuccess && cudaMemset(h->d_ready, 0, 256) == cudaSuccess &&
         cudaMallocHost(&h->h_marks, sizeof(unsigned long long) * 2 * MAX_COPY_CHUNKS) == cudaSuccess &&
         cudaMallocHost(&h->h_status, 64) == cudaSuccess && cudaMallocHost(&h->h_ready, 64) == cudaSuccess;
    if (ok) *h->h_status = 0;
    ok = ok && bfm::tensor_init() == 0;
    for (int i = 0; ok && i < N_TABLE_SLOTS; ++i) {
        ok = cudaEventCreateWithFlags(&h->table_ev[i], cudaEventDisableTiming) == cudaSuccess;
        if (ok) ok = cudaEventRecord(h->table_ev[i], h->stream) == cudaSuccess;
    }
    if (!ok) {
        g_create_error = std::string("stream/event/buffer creation failed: ") + cudaGetErrorString(cudaGetLastError());
        bfm_destroy(h);
        return BFM_ERR_CUDA;
    }
    *out = h;
    return BFM_OK;
}

int bfm_destroy(bfm_handle_t h) {
    if (!h) return BFM_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&h->state, &h->tables, &h->d_in, &h->lower, &h->bins, &h->finc, &h->xq, &h->xt, &h->d_res})
        if (b->p) cudaFree(b->p);
    if (h->d_ready) cudaFree(h->d_ready);
    if (h->d_prog) cudaFree(h->d_prog);
    if (h->d_queue) cudaFree(h->d_queue);
    if (h->h_marks) cudaFreeHost(h->h_marks);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->h_ready) cudaFreeHost(h->h_ready);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    h->pool.reset();
    for (int i = 0; i < N_TABLE_SLOTS; ++i) {
        if (h->h_tables[i]) cudaFreeHost(h->h_tables[i]);
        if (h->table_ev[i]) cudaEventDestroy(h->table_ev[i]);
    }
    if (h->h_out) cudaFreeHost(h->h_out);
    for (int i = 0; i < 4; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->last_ev) cudaEventDestroy(h->last_ev);
    for (cudaEvent_t e : h->chunk_ev)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->chunk_done_ev)
        if (e) cudaEventDestroy(e);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    if (h->in_stream) cudaStreamDestroy(h->in_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return BFM_OK;
}

const char *bfm_last_error(bfm_handle_t h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int check_call(bfm_handle_t h, const uint8_t *q, int32_t n_query_rows, const uint8_t *t, int32_t n_train_rows,
                      const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows, const bfm_options_t *opts) {
    if (n_problems < 0 || n_query_rows < 0 || n_train_rows < 0 || n_out_rows < 0)
        return fail(h, BFM_ERR_INVALID, "negative size");
    if (n_problems > 0 && !problems) return fail(h, BFM_ERR_INVALID, "problems is NULL");
    int rc = check_opts(h, opts, n_problems);
    if (rc) return rc;
    if (opts->mask_kind == BFM_MASK_DENSE && opts->mask_row_stride < (int64_t)problems[0].t_count)
        return fail(h, BFM_ERR_INVALID, "mask_row_stride must be >= t_count");
    if ((n_query_rows > 0 && !q) || (n_train_rows > 0 && !t)) return fail(h, BFM_ERR_INVALID, "descriptor pointer is NULL");
    CU_TRY(h, cudaSetDevice(h->device));
    return BFM_OK;
}

int bfm_match_batched(bfm_handle_t h, int mem, const uint8_t *q, int32_t n_query_rows, const uint8_t *t,
                      int32_t n_train_rows, const bfm_problem_t *problems, int32_t n_problems,
                      int32_t n_out_rows, const bfm_options_t *opts, int32_t *knn_idx, int32_t *knn_dist,
                      int32_t *m_query, int32_t *m_train, int32_t *m_dist, int32_t *m_count, void *stream) {
    if (!h) return BFM_ERR_INVALID;
    h->err.clear();
    int rc = check_call(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts);
    if (rc) return rc;
    const bfm_outputs_t out = {knn_idx, knn_dist, m_query, m_train, m_dist, m_count, 0, 0};
    if (mem == BFM_MEM_DEVICE)
        return run_device(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, &out, 1,
                          stream == BFM_STREAM_OWN ? h->stream : static_cast<cudaStream_t>(stream));
    if (mem == BFM_MEM_HOST)
        return run_host(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, out);
    return fail(h, BFM_ERR_INVALID, "mem must be BFM_MEM_HOST or BFM_MEM_DEVICE");
}

int bfm_match_batched_multi(bfm_handle_t h, const uint8_t *q, int32_t n_query_rows, const uint8_t *t,
                            int32_t n_train_rows, const bfm_problem_t *problems, int32_t n_problems,
                            int32_t n_out_rows, const bfm_options_t *opts, const bfm_outputs_t *dests,
                            int32_t n_dests, void *stream) {
    if (!h) return BFM_ERR_INVALID;
    h->err.clear();
    int rc = check_call(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts);
    if (rc) return rc;
    return run_device(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, dests, n_dests,
                      stream == BFM_STREAM_OWN ? h->stream : static_cast<cudaStream_t>(stream));
}

int bfm_match_batched_host_multi(bfm_handle_t h, const uint8_t *q, int32_t n_query_rows, const uint8_t *t,
                                 int32_t n_train_rows, const bfm_problem_t *problems, int32_t n_problems,
                                 int32_t n_out_rows, const bfm_options_t *opts, const bfm_outputs_t *host_out,
                                 const bfm_outputs_t *device_dests, int32_t n_device_dests) {
    if (!h) return BFM_ERR_INVALID;
    h->err.clear();
    if (!host_out) return fail(h, BFM_ERR_INVALID, "host_out is NULL");
    int rc = check_call(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts);
    if (rc) return rc;
    bfm_outputs_t user = *host_out;
    user.multicast = 0;
    return run_host(h, q, n_query_rows, t, n_train_rows, problems, n_problems, n_out_rows, opts, user, device_dests,
                    n_device_dests);
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
        if (value < 0 || value > MAX_COPY_CHUNKS) return fail(h, BFM_ERR_INVALID, "pipeline_chunks must be 0 (auto), 1 (off) .. 64");
        h->pipeline_chunks = value;
    } else if (k == "feeders") {
        if (value < -1 || value > bfm::FEED_MAX) return fail(h, BFM_ERR_INVALID, "feeders must be -1 (off: copy engine), 0 (auto) or 1..32");
        h->feeders = value;
    } else if (k == "pipeline_min_kb") {
        if (value < 0) return fail(h, BFM_ERR_INVALID, "pipeline_min_kb must be >= 0");
        h->pipeline_min_kb = value;
    } else if (k == "host_threads") {
        if (value < -1 || value > 64) return fail(h, BFM_ERR_INVALID, "host_threads must be -1 (off), 0 (auto) or 1..64");
        if (value != h->host_threads) h->pool.reset();
        h->host_threads = value;
    } else if (k == "test_stall") {   // test hook: the next SM-fed call's feeders deliver nothing (gate time-out path)
        h->test_stall = value != 0;
    } else if (k == "feed_rows") {
        if (value < 0) return fail(h, BFM_ERR_INVALID, "feed_rows must be >= 0");
        h->feed_rows = value;
    } else if (k == "window_bins") {
        if (value != 0 && value != 1) return fail(h, BFM_ERR_INVALID, "window_bins must be 0 (auto) or 1 (brute force)");
        h->window_bins = value;
    } else if (k == "taper") {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
            return fail(h, BFM_ERR_INVALID, "taper must be 0 (auto), 1 (off), 2, 4, 8 or 16 (guided)");
        h->taper = value;
    } else if (k == "taper_pct") {
        if (value < 0 || value > 90) return fail(h, BFM_ERR_INVALID, "taper_pct must be 0 (auto) .. 90");
        h->taper_pct = value;
    } else if (k == "ctas_per_sm") {
        if (value < 0 || value > 32) return fail(h, BFM_ERR_INVALID, "ctas_per_sm must be 0 (auto) .. 32");
        h->ctas_per_sm = value;
    } else if (k == "gss_div" || k == "gss_min") {
        if (value < 0) return fail(h, BFM_ERR_INVALID, k + " must be >= 0");
        (k == "gss_div" ? h->gss_div : h->gss_min) = value;
    } else if (k == "persistent") {
        if (value < 0 || value > 2) return fail(h, BFM_ERR_INVALID, "persistent must be 0 (auto), 1 (off) or 2 (always, for resident inputs)");
        h->persistent = value;
    } else if (k == "tensor") {
        if (value < 0 || value > 2) return fail(h, BFM_ERR_INVALID, "tensor must be 0 (auto), 1 (off) or 2 (whenever the call is eligible)");
        h->tensor = value;
    } else if (k == "tensor_chunks") {
        if (value < 0 || value > 8) return fail(h, BFM_ERR_INVALID, "tensor_chunks must be 0 (auto) .. 8 (copy chunks of a host batch that takes the tensor form)");
        h->tensor_chunks = value;
    } else if (k == "waves") {
        if (value < 0) return fail(h, BFM_ERR_INVALID, "waves must be >= 0");
        h->waves = value;
    } else {
        return fail(h, BFM_ERR_INVALID, "unknown tuning knob: " + k);
    }
    return BFM_OK;
}

int64_t bfm_kernel_launch_count(bfm_handle_t h) { return h ? h->launches : 0; }

int bfm_synchronize(bfm_handle_t h) {
    if (!h) return BFM_ERR_INVALID;
    h->err.clear();
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return BFM_OK;
}

int bfm_debug_timeline(bfm_handle_t h, uint64_t *device_buf, int32_t capacity_ctas) {
    if (!h || capacity_ctas < 0) return BFM_ERR_INVALID;
    h->trace = capacity_ctas > 0 ? reinterpret_cast<unsigned long long *>(device_buf) : nullptr;
    h->trace_cap = h->trace ? capacity_ctas : 0;
    return BFM_OK;
}

// Host-only: the work-item plan a batch would get.  No device is touched, so the planner's invariants are
// testable on a box without a GPU (tests/test_planner_cpu.py).
int bfm_plan_preview(const bfm_problem_t *problems, int32_t n_problems, int32_t queries_per_thread, int32_t slots,
                     int32_t segment_rows, int32_t waves, int32_t taper, int32_t taper_pct, int32_t *items_out,
                     int32_t capacity, int32_t *n_items, int32_t *segment_rows_out) {
    if (!problems || n_problems <= 0 || !n_items || slots <= 0 || capacity < 0 || (capacity > 0 && !items_out)) return BFM_ERR_INVALID;
    if (queries_per_thread != 1 && queries_per_thread != 2 && queries_per_thread != 4) return BFM_ERR_INVALID;
    if (segment_rows < 0 || waves < 0 || taper_pct < 0 || taper_pct > 90 ||
        (taper != 0 && taper != 1 && taper != 2 && taper != 4 && taper != 8 && taper != 16))
        return BFM_ERR_INVALID;
    for (int p = 0; p < n_problems; ++p)
        if (problems[p].q_count < 0 || problems[p].t_count < 0 || problems[p].q_begin < 0 || problems[p].t_begin < 0 ||
            problems[p].out_begin < 0 || problems[p].t_count >= BFM_MAX_TRAIN_ROWS || problems[p].q_count >= BFM_MAX_QUERY_ROWS)
            return BFM_ERR_INVALID;
    bfm_handle_s tmp;   // knobs only: plan_segments reads nothing else
    tmp.segment_rows = segment_rows;
    tmp.waves = waves;
    tmp.taper = taper;
    tmp.taper_pct = taper_pct;
    std::vector<Segment> segs;
    std::vector<int> seg_begin;
    int L = 0;
    plan_segments(&tmp, problems, n_problems, queries_per_thread, slots, segs, seg_begin, &L);
    *n_items = (int32_t)segs.size();
    if (segment_rows_out) *segment_rows_out = L;
    const size_t n = std::min(segs.size(), (size_t)capacity);
    static_assert(sizeof(Segment) == 8 * sizeof(int32_t), "a work item is eight int32");
    if (n) std::memcpy(items_out, segs.data(), n * sizeof(Segment));
    return BFM_OK;
}

int bfm_plan_preview_tiles(const bfm_problem_t *problems, int32_t n_problems, int32_t n_ctas, int32_t *tiles_out,
                           int32_t *tile_cta_out, int32_t tile_capacity, int32_t *n_tiles) {
    if (!problems || n_problems <= 0 || !n_tiles || n_ctas <= 0 || tile_capacity < 0) return BFM_ERR_INVALID;
    for (int p = 0; p < n_problems; ++p)
        if (problems[p].q_count < 0 || problems[p].q_count >= BFM_MAX_QUERY_ROWS) return BFM_ERR_INVALID;
    std::vector<int2> work;
    std::vector<bfm::FinTile> tiles;
    plan_tiles(problems, n_problems, n_ctas, tiles, work);
    *n_tiles = (int32_t)tiles.size();
    for (size_t c = 0; c < work.size(); ++c)
        for (int f = work[c].x; f < work[c].x + work[c].y; ++f) {
            if (f >= tile_capacity) break;
            if (tiles_out) {
                const int32_t row[5] = {tiles[f].problem, tiles[f].row0, tiles[f].index, tiles[f].n_tiles, tiles[f].slot0};
                std::memcpy(tiles_out + 5 * (size_t)f, row, sizeof(row));
            }
            if (tile_cta_out) tile_cta_out[f] = (int32_t)c;
        }
    return BFM_OK;
}

int bfm_plan_preview_tensor(const bfm_problem_t *problems, int32_t n_problems, int32_t n_sms, int32_t *items_out, int32_t capacity,
                            int32_t *n_items, int32_t *segment_rows_out, int32_t *plane_rows_out) {
    if (!problems || n_problems <= 0 || !n_items || n_sms <= 0 || capacity < 0 || (capacity > 0 && !items_out)) return BFM_ERR_INVALID;
    for (int p = 0; p < n_problems; ++p)
        if (problems[p].q_count < 0 || problems[p].t_count < 0 || problems[p].q_begin < 0 || problems[p].t_begin < 0 ||
            problems[p].out_begin < 0 || problems[p].t_count >= BFM_MAX_TRAIN_ROWS || problems[p].q_count >= BFM_MAX_QUERY_ROWS)
            return BFM_ERR_INVALID;
    std::vector<int32_t> xq0, xt0;
    long long rows[2] = {0, 0};
    tensor_bases(problems, n_problems, xq0, xt0, rows);
    std::vector<Segment> segs;
    std::vector<int> seg_begin;
    int L = 0;
    plan_items_tensor(problems, n_problems, n_sms, xq0, xt0, segs, seg_begin, &L);
    *n_items = (int32_t)segs.size();
    if (segment_rows_out) *segment_rows_out = L;
    if (plane_rows_out) { plane_rows_out[0] = (int32_t)rows[0]; plane_rows_out[1] = (int32_t)rows[1]; }
    const size_t n = std::min(segs.size(), (size_t)capacity);
    if (n) std::memcpy(items_out, segs.data(), n * sizeof(Segment));
    return BFM_OK;
}

int bfm_plan_preview_host_chunks(const bfm_problem_t *problems, int32_t n_problems, int32_t n_query_rows, int32_t n_train_rows,
                                 int32_t n_sms, int32_t forced_chunks, int32_t *n_chunks, int32_t *problems_per_chunk) {
    if (!problems || n_problems <= 0 || n_sms <= 0 || forced_chunks < 0 || forced_chunks > 8 || n_query_rows < 0 || n_train_rows < 0 || !n_chunks || !problems_per_chunk)
        return BFM_ERR_INVALID;
    int per = 0;
    *n_chunks = plan_host_chunks(problems, n_problems, ((size_t)n_query_rows + (size_t)n_train_rows) * 32, n_sms, forced_chunks, &per);
    *problems_per_chunk = per;
    return BFM_OK;
}

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
