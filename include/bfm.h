/*
 * bfm.h - C ABI of the B200-native brute-force Hamming matcher (libbfm_b200.so).
 *
 * This is the drop-in boundary for boslam's one data-parallel hot path: everything
 * `cv2.BFMatcher_create(cv2.NORM_HAMMING, crossCheck)` does for
 *     reference slam/tracking.py:45      (matcher construction; same at slam/local_mapping.py:21,
 *                                         slam/covisibility_graph.py:34, experiments/pnp_*_tracking.py:11)
 *     reference slam/tracking.py:56      matcher.match(frame.des, kf_ref.desf())      frame <-> ref keyframe
 *     reference slam/tracking.py:121     matcher.match(frame.des, feats)              frame <-> local map
 *     reference experiments/pnp_one_way_tracking.py:30, pnp_two_way_tracking.py:42    frame <-> frame
 * plus the batched keyframe-pair form the local-mapping (slam/local_mapping.py:41-44, a stub) and
 * loop-closing (slam/loop_closing.py:13-29, stubs) workloads need.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary.
 *   - descriptors are 256-bit ORB rows: uint8[N][32], row-major, base pointer 16-byte aligned.
 *   - every function returns a bfm_status (0 = OK); nothing throws.  Text for the last failure
 *     of a handle: bfm_last_error(handle) (bfm_last_error(NULL) for a failed bfm_create).
 *   - `mem` says where the caller's buffers live:
 *       BFM_MEM_DEVICE  all data pointers are device pointers on the handle's GPU, the call is
 *                       asynchronous on `stream` (a cudaStream_t passed as void*; NULL is CUDA's
 *                       legacy default stream, exactly as in the runtime API; BFM_STREAM_OWN =
 *                       the handle's own stream) and outputs are valid once it has drained.
 *       BFM_MEM_HOST    all data pointers are host pointers; the upload overlaps the ONE kernel
 *                       launch of the call (see bfm_host_alloc below), the kernel writes results into
 *                       pinned host memory, and the call returns after the results are in the
 *                       caller's buffers (this is the e2e path bench.py times).
 *     The `problems` table and the options struct are always host memory.
 *   - the caller owns every input and output buffer; the handle owns only its workspace.
 *   - streams: calls on one handle may use different streams (BFM_MEM_DEVICE with caller streams, the
 *     handle's own stream for BFM_MEM_HOST / bfm_track_local_map).  The handle's workspace is shared, so
 *     every call leaves an event behind and a call arriving on another stream first waits for it
 *     (cudaStreamWaitEvent): calls on one handle execute in the order they were issued, whatever their streams.
 *   - a handle is NOT re-entrant: one thread at a time per handle.  Different handles are
 *     independent (own stream, own workspace), which is how boslam's two threads
 *     (tracking in main, local mapping in a daemon thread, reference slam/main.py:37-47) use it.
 *   - result semantics are cv2.BFMatcher's (SURVEY.md section 8(c) rules R1-R10):
 *     distance = popcount(q XOR t); neighbours ascend by (distance, trainIdx); an unfilled
 *     neighbour slot has idx = -1 and dist = -1; cross-check is the true mutual-nearest test
 *     with lowest-index ties on both sides; a masked pair never competes.
 */
#ifndef BFM_H
#define BFM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFM_ABI_VERSION 5
#define BFM_DESC_BYTES 32
/* per-problem limits of the packed (distance, index) keys the kernels reduce over;
 * cv2 itself refuses train sets of 2^18 rows or more (matchers.cpp:860, rule R7). */
#define BFM_MAX_TRAIN_ROWS (1 << 22)
#define BFM_MAX_QUERY_ROWS (1 << 22)
#define BFM_MAX_K 16
#define BFM_STREAM_OWN ((void *)(intptr_t)-1) /* `stream` value meaning "the handle's own stream" */

typedef struct bfm_handle_s *bfm_handle_t;

typedef enum bfm_status {
    BFM_OK = 0,
    BFM_ERR_INVALID = 1,     /* bad argument (message says which) */
    BFM_ERR_CUDA = 2,        /* CUDA runtime error (message carries cudaGetErrorString) */
    BFM_ERR_NOMEM = 3,
    BFM_ERR_UNSUPPORTED = 4  /* valid in cv2 but outside this engine (e.g. k > BFM_MAX_K) */
} bfm_status;

typedef enum bfm_mem { BFM_MEM_HOST = 0, BFM_MEM_DEVICE = 1 } bfm_mem;

typedef enum bfm_mask_kind {
    BFM_MASK_NONE = 0,
    BFM_MASK_DENSE = 1,   /* cv2-style uint8[Q][T], non-zero = allowed; single problem only */
    BFM_MASK_WINDOW = 2   /* allowed iff |qx-tx| < r and |qy-ty| < r (float32), per-row pixel coords */
} bfm_mask_kind;

/* One (query set, train set) problem of a batch.  Rows index the caller's descriptor arrays,
 * so several problems may share a query block (loop closing: one keyframe vs N candidates). */
typedef struct bfm_problem {
    int32_t q_begin;    /* first query row */
    int32_t q_count;
    int32_t t_begin;    /* first train row */
    int32_t t_count;
    int32_t out_begin;  /* first output row; problem p owns output rows [out_begin, out_begin+q_count) */
    int32_t reserved;
} bfm_problem_t;

typedef struct bfm_options {
    int32_t k;               /* neighbours per query, 1..BFM_MAX_K (cross_check needs k == 1) */
    int32_t cross_check;     /* 0/1: keep only mutual nearest neighbours (cv2 crossCheck=True) */
    int32_t mask_kind;       /* bfm_mask_kind */
    int32_t max_distance;    /* < 0: off; else keep matches with distance <= max_distance
                                (slam/tracking.py:121 `<= 30`; pass 29 for :57's strict `< 30`) */
    double ratio;            /* < 0: off; else keep a row iff it has 2 neighbours and
                                (double)d1 < ratio * (double)d2  (fp64, as the Python test does) */
    float window_radius;     /* BFM_MASK_WINDOW */
    int32_t reserved0;
    const uint8_t *mask;     /* BFM_MASK_DENSE: [q_count][mask_row_stride] bytes */
    int64_t mask_row_stride; /* bytes between mask rows (>= t_count) */
    const float *q_xy;       /* BFM_MASK_WINDOW: float32[n_query_rows][2], same row indexing as q */
    const float *t_xy;       /* BFM_MASK_WINDOW: float32[n_train_rows][2], same row indexing as t */
} bfm_options_t;

/* ---- lifetime ---------------------------------------------------------------------------- */
int bfm_abi_version(void);
int bfm_create(int device, bfm_handle_t *out);
int bfm_destroy(bfm_handle_t h);
const char *bfm_last_error(bfm_handle_t h);

/* ---- the hot path ------------------------------------------------------------------------- */
/*
 * General batched entry point: replaces one cv2 `match` / `knnMatch` call per problem.
 *   q, t            descriptor arrays, uint8[n_query_rows][32] / uint8[n_train_rows][32]
 *   problems        host table of n_problems entries
 *   n_out_rows      number of output rows (max over problems of out_begin + q_count)
 * Outputs (each may be NULL to skip it):
 *   knn_idx/knn_dist  int32[n_out_rows][k]: the raw k-NN table (cv2 knnMatch), -1 = no neighbour
 *   m_query/m_train/m_dist  int32[n_out_rows]: the filtered match list of problem p is packed,
 *                     ascending queryIdx, into [out_begin, out_begin + m_count[p]); a row is a
 *                     match iff it has a neighbour and passes cross_check / ratio / max_distance
 *                     (cv2 `match` + the caller-side filters of slam/tracking.py:57,121).
 *                     m_query / m_train are problem-local indices (cv2 queryIdx / trainIdx).
 *   m_count           int32[n_problems]
 */
int bfm_match_batched(bfm_handle_t h, int mem,
                      const uint8_t *q, int32_t n_query_rows,
                      const uint8_t *t, int32_t n_train_rows,
                      const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
                      const bfm_options_t *opts,
                      int32_t *knn_idx, int32_t *knn_dist,
                      int32_t *m_query, int32_t *m_train, int32_t *m_dist, int32_t *m_count,
                      void *stream);

/*
 * Multi-destination form (device memory only): the same batch, with every result written to
 * n_dests (1..8) sets of output buffers by the kernel's epilogue.  dests[0] is normally this GPU's
 * own buffers; dests[1..] are the corresponding buffers of NVLink peers (peer-mapped device
 * pointers, e.g. from CUDA IPC / torch symmetric memory), already offset to this rank's slice.
 * This is the multi-GPU gather of SURVEY.md 8(e) fused into the matching kernel: the match lists
 * of the sharded keyframe-pair batches (the shape slam/loop_closing.py:13-29 would issue) land in
 * every rank's table without a separate collective; the caller only needs a barrier afterwards.
 */
typedef struct bfm_outputs {
    int32_t *knn_idx, *knn_dist;             /* int32[n_out_rows][k], both or neither */
    int32_t *m_query, *m_train, *m_dist;     /* int32[n_out_rows] */
    int32_t *m_count;                        /* int32[n_problems]; NULL skips the match list */
    int32_t multicast;                       /* 1: the pointers are addresses of an NVSwitch multicast object (NVLS)
                                                spanning every rank's buffer; results are written with multimem.st,
                                                one store reaching all GPUs */
    int32_t reserved;
} bfm_outputs_t;

int bfm_match_batched_multi(bfm_handle_t h,
                            const uint8_t *q, int32_t n_query_rows,
                            const uint8_t *t, int32_t n_train_rows,
                            const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
                            const bfm_options_t *opts,
                            const bfm_outputs_t *dests, int32_t n_dests, void *stream);

/*
 * The same with HOST caller arrays (the BFM_MEM_HOST path of bfm_match_batched: upload overlapped with the one
 * kernel launch, results written into `host_out`, host pointers) plus n_device_dests (0..7) destination sets in
 * DEVICE memory - this GPU's, NVLink peers' or an NVSwitch multicast address - written by the same epilogue.
 * One call = host copies + match + multi-GPU exchange of a rank's pair block; the caller only
 * needs a barrier afterwards.  Synchronous like every BFM_MEM_HOST call.
 */
int bfm_match_batched_host_multi(bfm_handle_t h,
                                 const uint8_t *q, int32_t n_query_rows,
                                 const uint8_t *t, int32_t n_train_rows,
                                 const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
                                 const bfm_options_t *opts, const bfm_outputs_t *host_out,
                                 const bfm_outputs_t *device_dests, int32_t n_device_dests);

/* Single-problem conveniences (what slam/tracking.py:56,121 bind to). */
int bfm_knn(bfm_handle_t h, int mem, const uint8_t *q, int32_t nq, const uint8_t *t, int32_t nt,
            const bfm_options_t *opts, int32_t *knn_idx, int32_t *knn_dist, void *stream);
int bfm_match(bfm_handle_t h, int mem, const uint8_t *q, int32_t nq, const uint8_t *t, int32_t nt,
              const bfm_options_t *opts, int32_t *m_query, int32_t *m_train, int32_t *m_dist,
              int32_t *m_count, void *stream);

/* ---- frame <-> local map tracking, device resident ---------------------
 *
 * bfm_map_t is the device copy of what the reference keeps per map point (slam/nodes.py:115-118:
 * MapPoint.feat uint8[32], MapPoint.pt3d float64[3], MapPoint.n float64[3]), addressed by a slot
 * number the caller assigns (e.g. the MapPoint id).  bfm_map_update is the upsert the reference
 * performs at slam/covisibility_graph.py:128-134 (new point) and slam/nodes.py:153-154 (descriptor /
 * normal refresh); pass NULL for a field to leave it unchanged.
 *
 * bfm_track_local_map replaces reference slam/tracking.py:96-128 in one call:
 *   edges[e]          map-point slot of the e-th (local keyframe, map point) edge, in the order the
 *                     reference's double loop visits them (duplicates are expected: SURVEY.md 0.4)
 *   visibility        pixel = cam_map(SE3Quat(R, t) * pt3d); keep iff 0 <= u < width, 0 <= v < height
 *                     and dot(see_vector, normal) < cos_max   (:102-104, exactly as written)
 *   train set         the kept edges, in edge order: trainIdx j = j-th kept edge (:107-110 + np.stack)
 *   match             opts as in bfm_match (cross_check / k, ratio / max_distance); mask_kind may be
 *                     BFM_MASK_WINDOW: allowed iff |kp - pixel| < window_radius in both axes
 *   outputs (host)    visible_edges[n_visible] edge numbers of the train rows (-> ids_matching_kfs /
 *                     ids_matching_mps by the caller's own tables), visible_pixels[n_visible][2];
 *                     m_query / m_train / m_dist [n_matches] as bfm_match; m_edge[n_matches] =
 *                     visible_edges[m_train]; m_pts3d[n_matches][3] = pts3d[inds], m_kp[n_matches][2] =
 *                     kp_arr[inds_frame] (:128, the arrays CamOnlyBA consumes).  Any may be NULL.
 * The quaternion is (w, x, y, z) of g2o::SE3Quat(R, t).rotation(); arithmetic is fp64.
 */
typedef struct bfm_map_s *bfm_map_t;
typedef struct bfm_track_params {
    double q[4];            /* unit quaternion w, x, y, z */
    double t[3];
    double see_vector[3];   /* Frame.see_vector, camera.py:24-29 */
    double fx, fy, cx, cy;  /* config.py:36-41 */
    double cos_max;         /* cos(60 deg) in the reference, slam/tracking.py:20,104 */
    int32_t width, height;  /* config.py:40-41 */
} bfm_track_params_t;

int bfm_map_create(bfm_handle_t h, int32_t capacity, bfm_map_t *out);
int bfm_map_destroy(bfm_map_t m);
int bfm_map_update(bfm_map_t m, int32_t n, const int32_t *slots, const uint8_t *desc, const double *pt3d,
                   const double *normal);
int bfm_track_local_map(bfm_map_t m, const bfm_track_params_t *tp, const int32_t *edges, int32_t n_edges,
                        const uint8_t *q_desc, const double *q_kp, int32_t nq, const bfm_options_t *opts,
                        int32_t *visible_edges, double *visible_pixels, int32_t *m_query, int32_t *m_train,
                        int32_t *m_dist, int32_t *m_edge, double *m_pts3d, double *m_kp, int32_t *n_visible,
                        int32_t *n_matches);

/* Keyframe vote, reference slam/tracking.py:154:
 *     Counter(ids_matching_kfs[inds[inliers]]).most_common(top)
 * over the match list of the LAST bfm_track_local_map call on this map (still on the device; no other call on the map
 * in between).  edge_kf[e] = id of the keyframe of edge e (the caller's table behind ids_matching_kfs, :108, one
 * entry per edge of the tracking call); inliers[i] = positions in that call's match list which the pose optimisation
 * kept (CamOnlyBA's inlier set, :128-139; positions outside the list are ignored).  kf_ids / kf_counts [top]: the
 * keyframes most frequent first, ties in order of first appearance among the inliers, exactly as Counter.most_common
 * orders them; *n_kfs = how many were written.  At most 4096 inliers (a frame has 2000 features).  Host pointers. */
int bfm_keyframe_vote(bfm_map_t m, const int32_t *edge_kf, int32_t n_edges, const int32_t *inliers, int32_t n_inliers,
                      int32_t top, int32_t *kf_ids, int32_t *kf_counts, int32_t *n_kfs);

/* ---- representative descriptor of map points --------------------------------
 * reference slam/nodes.py:146-153 (MapPoint.add_observation), batched over the map points a new keyframe
 * touches (slam/covisibility_graph.py:124-134): obs uint8[n_points][max_obs][32] holds each point's stored
 * observations (first counts[p] rows valid, max_obs <= 16; the reference keeps at most 10);
 * out_idx[p] = argmin_j median_i Hamming(obs_i, obs_j) with numpy's conventions (median of an even count
 * = mean of the two middle values, argmin = lowest j on ties), -1 for a point without observations.
 * Host pointers. */
int bfm_select_representative(bfm_handle_t h, const uint8_t *obs, const int32_t *counts, int32_t n_points,
                              int32_t max_obs, int32_t *out_idx);

/* ---- introspection / tuning (used by bench.py and the tests; not needed by a call site) --- */
typedef struct bfm_launch_info {
    int32_t kernels_launched;   /* CUDA kernels launched by the last call on this handle (1: scan and
                                   finalize are one kernel) */
    int32_t scan_grid;          /* CTAs of the matching kernel */
    int32_t scan_block;         /* threads per CTA */
    int32_t queries_per_thread; /* register tile R */
    int32_t popc_mode;          /* popcount evaluation: 8 plain, 6/5/4 carry-save, 50/40 transformed carry-save */
    int32_t segments;           /* (query block, train range) work items */
    int32_t train_rows_per_segment;
    int32_t copy_chunks;        /* BFM_MEM_HOST: slices the upload was delivered in while the kernel ran - feed rounds
                                   (pinned inputs, POPC kernel), copy-engine chunks (pageable inputs; pinned inputs of the
                                   tensor form, where every chunk has its own launches); 1 = no overlap */
    float scan_ms;              /* device time of the kernel of the last call when timing is on */
    float total_ms;             /* same (kept for ABI stability) */
} bfm_launch_info_t;

int bfm_get_launch_info(bfm_handle_t h, bfm_launch_info_t *out);
/* knob: "popc_mode" {0=auto,8,6,5,4,50,40}, "queries_per_thread" {0=auto,1,2,4}, "timing" {0,1},
 *       "segment_rows" {0=auto, n}, "waves" {0=auto, n}, "pipeline_chunks" {0=auto, 1=off, n <= 64},
 *       "feeders" {0=auto (24), -1=off: copy-engine chunks, n <= 32}, "feed_rows" {0=auto, rows per feed round},
 *       "host_threads" {0=auto (<= 8), -1=off, n: staging threads for pageable caller arrays},
 *       "pipeline_min_kb" {0=auto (1024): smallest upload, in KB, that overlaps the kernel},
 *       "window_bins" {0=auto, 1=brute-force window kernel},
 *       "taper" {0=auto (4), 1=off, 2, 4, 8: finest divisor of the segment length in the tail of a batch},
 *       "taper_pct" {0=auto (10): percent of the batch's work, at its end, that is cut finer},
 *       "persistent" {0=auto, 1=off, 2=always}: resident inputs can take the persistent form of the kernel - at most
 *                     one wave of CTAs drawing work items from a ticket counter, then finalizing the problems tile by
 *                     tile inside the same launch; auto uses it for launches of up to 0.16 G pairs (a tracking frame, a
 *                     local-mapping batch, a rank's share of a sharded loop-closing batch) and the static form (one
 *                     work item per CTA, the CTA that completes a problem finalizes it) for longer ones and on the
 *                     gated host path,
 *       "taper" = 16 with "gss_div" / "gss_min": guided item lengths (an experiment, see profiles/r02_kernel_forms.md),
 *       "tensor" {0=auto, 1=off, 2=whenever eligible}: calls without mask / window / cross-check and k <= 2 can take the
 *                     tensor form (bfm_tensor.cuh: the descriptors expanded to one s8 per bit, distances as dot products
 *                     on tcgen05, the same finalize) - the same integers out; auto uses it from 8 M pairs per launch on.
 *                     A BFM_MEM_HOST batch of >= 8 problems and >= 4 MB in pinned memory that is eligible is uploaded by
 *                     the copy engine in chunks of whole problems, each matched by the tensor launches as it lands,
 *       "tensor_chunks" {0=auto, 2..8}: number of those chunks */
int bfm_set_tuning(bfm_handle_t h, const char *knob, int32_t value);
/* Wait until everything queued on the handle's own stream (BFM_STREAM_OWN calls) has completed. */
int bfm_synchronize(bfm_handle_t h);
/* total kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t bfm_kernel_launch_count(bfm_handle_t h);
/* Debug timeline of the matching kernel (tools/timeline_probe.py): with a device buffer uint64[capacity_ctas][8],
 * CTA b of every following call's launch stores the GPU's %globaltimer (ns) at points of its life -
 * [0] entry, [1] inputs requested, [2] first train chunk in shared memory, [3] scan done, [4] row keys committed,
 * [5] completion counted, [6] problem finalized (only the CTA that did it) - into row b (the tile-parallel finalize
 * kernels of a large single problem append their CTAs: [0] entry, [6] exit).  NULL / 0 switches it off (the default). */
int bfm_debug_timeline(bfm_handle_t h, uint64_t *device_buf, int32_t capacity_ctas);

/* Host-only preview of the work-item plan of a batch (no device needed; the CPU tests of the planner use it).
 * A work item is one CTA's share: a block of 128 x queries_per_thread query rows against a contiguous train
 * range of one problem.  `slots` = CTAs resident on the device at once (148 SMs x 8 on a B200); the knobs have
 * bfm_set_tuning's meaning (0 = auto).  items_out: int32[capacity][8] = {q_row0, q_valid, q_local0, out_row0,
 * t_row0, t_count, t_local0, problem} in launch order; *n_items is the full count even when it exceeds capacity. */
int bfm_plan_preview(const bfm_problem_t *problems, int32_t n_problems, int32_t queries_per_thread, int32_t slots,
                     int32_t segment_rows, int32_t waves, int32_t taper, int32_t taper_pct, int32_t *items_out,
                     int32_t capacity, int32_t *n_items, int32_t *segment_rows_out);

/* Host-only preview of the finalize tiles of the persistent form (resident inputs: a grid of at most one wave draws the
 * work items of bfm_plan_preview from a ticket counter, then finalizes the problems tile by tile, 512 query rows per
 * tile).  tiles_out: int32[tile_capacity][5] = {problem, row0, tile number, tiles of the problem, look-back slot of the
 * problem's tile 0}; tile_cta_out[f] = CTA (0 .. n_ctas - 1) that finalizes tile f.  A tile only waits for tiles of CTAs
 * with a lower or equal index. */
int bfm_plan_preview_tiles(const bfm_problem_t *problems, int32_t n_problems, int32_t n_ctas, int32_t *tiles_out,
                           int32_t *tile_cta_out, int32_t tile_capacity, int32_t *n_tiles);

/* Host-only preview of the work items of the tensor form (bfm_tensor.cuh): blocks of 256 query rows against train ranges
 * that are multiples of 128 rows (the last range of a problem ends where the problem ends), walked by one persistent CTA
 * per SM (`n_sms`).  items_out as in bfm_plan_preview, except that q_row0 / t_row0 are rows of the EXPANDED planes, where
 * every problem has an 8-row-aligned base; plane_rows_out[2] = rows of the query / train planes. */
int bfm_plan_preview_tensor(const bfm_problem_t *problems, int32_t n_problems, int32_t n_sms, int32_t *items_out, int32_t capacity,
                            int32_t *n_items, int32_t *segment_rows_out, int32_t *plane_rows_out);

/* Host-only preview of the copy chunks a BFM_MEM_HOST batch of the tensor form is uploaded in (whole problems; with equal
 * shapes a chunk is a whole number of rounds of the persistent scan).  forced_chunks = the "tensor_chunks" knob. */
int bfm_plan_preview_host_chunks(const bfm_problem_t *problems, int32_t n_problems, int32_t n_query_rows, int32_t n_train_rows,
                                 int32_t n_sms, int32_t forced_chunks, int32_t *n_chunks, int32_t *problems_per_chunk);

/* Integer-pipe micro-benchmark: the roofline denominator.
 * test: 0 POPC, 1 LOP3, 2 IADD3, 3 POPC+LOP3 1:1, 4 POPC+2xLOP3, 5 REDUX.MIN, 6 IMAD, 7 VIMNMX,
 *       8 POPC+IMAD 1:1, 9 XOR+POPC+IADD pair loop (8:8:4, the plain per-pair mix).
 * Writes thread-level ops per clock per SM (clock64 based), ops per second (event based) and
 * the SM clock in MHz implied by the two. */
int bfm_microbench(int device, int test, int iters, double *ops_per_clk_per_sm,
                   double *ops_per_s, double *sm_mhz);
/* Pinned (page-locked) host memory.  On the BFM_MEM_HOST path pinned inputs are streamed into HBM by the
 * matching kernel's own feeder CTAs and pinned outputs are written by the kernel directly (bench.py's e2e
 * leg uses such buffers); pageable buffers work too, through copy-engine chunks and a pinned staging block. */
int bfm_host_alloc(uint64_t bytes, void **out);
int bfm_host_free(void *p);
int bfm_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz,
                    char *name, int name_len);

#ifdef __cplusplus
}
#endif
#endif /* BFM_H */
