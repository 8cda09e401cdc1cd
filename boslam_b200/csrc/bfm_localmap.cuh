// bfm_localmap.cuh - frame-to-local-map tracking on the device.
// (included by bfm_api.cu inside its anonymous namespace, after run_device)
//
// What it replaces: the candidate loop of reference slam/tracking.py:96-111 (for every (keyframe,
// map point) edge of the local map: g2o pose * point, CameraParameters.cam_map, image-bounds test,
// viewing-angle test, collect descriptor + 3-D point), the np.stack + matcher.match + distance gate
// of :119-121, and the index gathers of :126-128 that feed CamOnlyBA.  Here the map points live
// in device memory (bfm_map_t: descriptor, 3-D point, normal per slot; reference slam/nodes.py:115-118)
// and one call runs
//     project + visibility      one thread per edge (fp64, explicit IEEE operations - no FMA
//                               contraction - so the CPU oracle reproduces every decision bit for bit)
//     ordered compaction        surviving edges keep the reference's iteration order, so train row j
//                               is the j-th appended feature exactly as in the reference's list;
//                               descriptors, pixels and points are gathered into contiguous arrays
//     match                     the fused matching kernel, train-set size read from device memory
//     gather                    matched (3-D point, frame pixel) pairs into the contiguous float64
//                               arrays CamOnlyBA / solvePnPRansac consume
// with one H2D copy (pinned staging) in front and one D2H copy behind.

struct MapView {
    const uint8_t *desc;   // [capacity][32]
    const double *pt3d;    // [capacity][3]
    const double *normal;  // [capacity][3]
};

struct ProjectParams {
    double qw, qx, qy, qz;   // unit quaternion of the frame pose (g2o::SE3Quat(R, t).rotation())
    double tx, ty, tz;
    double sx, sy, sz;       // frame.see_vector
    double fx, fy, cx, cy;
    double width, height;
    double cos_max;          // keep iff dot(see_vector, normal) < cos_max   (slam/tracking.py:104, as written)
};

constexpr int LM_NT = 256;

// pose * X as Eigen evaluates a quaternion-vector product (g2o::SE3Quat::map = _r * xyz + _t):
//   uv = 2 * (q.vec x v);  r = v + w * uv + q.vec x uv
// then g2o::CameraParameters::cam_map: (x / z) * f + c.  Every operation is an explicit
// round-to-nearest fp64 instruction, in the order the oracle (oracle/localmap_oracle.py) uses.
__device__ __forceinline__ bool lm_project(const ProjectParams &pp, const double X[3], const double N[3], double *px, double *py) {
    const double ux = __dsub_rn(__dmul_rn(pp.qy, X[2]), __dmul_rn(pp.qz, X[1]));
    const double uy = __dsub_rn(__dmul_rn(pp.qz, X[0]), __dmul_rn(pp.qx, X[2]));
    const double uz = __dsub_rn(__dmul_rn(pp.qx, X[1]), __dmul_rn(pp.qy, X[0]));
    const double vx = __dadd_rn(ux, ux), vy = __dadd_rn(uy, uy), vz = __dadd_rn(uz, uz);
    const double cx_ = __dsub_rn(__dmul_rn(pp.qy, vz), __dmul_rn(pp.qz, vy));
    const double cy_ = __dsub_rn(__dmul_rn(pp.qz, vx), __dmul_rn(pp.qx, vz));
    const double cz_ = __dsub_rn(__dmul_rn(pp.qx, vy), __dmul_rn(pp.qy, vx));
    const double rx = __dadd_rn(__dadd_rn(X[0], __dmul_rn(pp.qw, vx)), cx_);
    const double ry = __dadd_rn(__dadd_rn(X[1], __dmul_rn(pp.qw, vy)), cy_);
    const double rz = __dadd_rn(__dadd_rn(X[2], __dmul_rn(pp.qw, vz)), cz_);
    const double x = __dadd_rn(rx, pp.tx), y = __dadd_rn(ry, pp.ty), z = __dadd_rn(rz, pp.tz);
    const double u = __dadd_rn(__dmul_rn(__ddiv_rn(x, z), pp.fx), pp.cx);
    const double v = __dadd_rn(__dmul_rn(__ddiv_rn(y, z), pp.fy), pp.cy);
    *px = u;
    *py = v;
    const bool in_image = (0.0 <= u) && (u < pp.width) && (0.0 <= v) && (v < pp.height);   // NaN -> false
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn(pp.sx, N[0]), __dmul_rn(pp.sy, N[1])), __dmul_rn(pp.sz, N[2]));
    return in_image && (dot < pp.cos_max);
}

// pass 1: visibility flag + pixel per edge, survivors per block
__global__ void __launch_bounds__(LM_NT) lm_project_kernel(const ProjectParams pp, const MapView map, const int32_t *edges,
                                                           int32_t n_edges, int32_t capacity, double2 *pix, uint8_t *flag,
                                                           int32_t *block_count, const double *q_kp, float2 *q_xy, int32_t nq) {
    const int e = blockIdx.x * LM_NT + threadIdx.x;
    // the frame's keypoints as float2 for the window predicate (grid-stride: nq may exceed n_edges)
    if (q_xy != nullptr)
        for (int i = e; i < nq; i += gridDim.x * LM_NT) q_xy[i] = make_float2((float)q_kp[2 * (size_t)i], (float)q_kp[2 * (size_t)i + 1]);
    bool vis = false;
    if (e < n_edges) {
        const int slot = edges[e];
        double u = 0.0, v = 0.0;
        if (slot >= 0 && slot < capacity) {
            const double *Xp = map.pt3d + 3 * (size_t)slot, *Np = map.normal + 3 * (size_t)slot;
            const double X[3] = {Xp[0], Xp[1], Xp[2]}, N[3] = {Np[0], Np[1], Np[2]};
            vis = lm_project(pp, X, N, &u, &v);
        }
        pix[e] = make_double2(u, v);
        flag[e] = vis ? 1 : 0;
    }
    const int c = __syncthreads_count(vis);
    if (threadIdx.x == 0) block_count[blockIdx.x] = c;
}

// pass 2: ordered compaction + gathers.  Survivor number j (in edge order) becomes train row j.
__global__ void __launch_bounds__(LM_NT) lm_compact_kernel(const MapView map, const int32_t *edges, int32_t n_edges,
                                                           const double2 *pix, const uint8_t *flag, const int32_t *block_count,
                                                           int32_t *vis_edge, double2 *vis_pix, float2 *t_xy, uint4 *t_desc,
                                                           double *vis_pt3d, int32_t *n_visible) {
    __shared__ int s_red[LM_NT / 32];
    __shared__ int s_warp[LM_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // survivors in the blocks before this one
    int part = 0;
    for (int b = tid; b < (int)blockIdx.x; b += LM_NT) part += block_count[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[warp] = part;
    const int e = blockIdx.x * LM_NT + tid;
    const bool vis = e < n_edges && flag[e] != 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int pos = 0, mine = 0;
#pragma unroll
    for (int w = 0; w < LM_NT / 32; ++w) {
        pos += s_red[w];
        if (w < warp) pos += s_warp[w];
        mine += s_warp[w];
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        int before = 0;
        for (int w = 0; w < LM_NT / 32; ++w) before += s_red[w];
        n_visible[0] = before + mine;
        n_visible[1] = 0;   // match count: stays 0 unless the matcher runs (it writes hdr[1] in its finalize)
    }
    if (!vis) return;
    pos += __popc(bal & ((1u << lane) - 1u));
    const int slot = edges[e];
    const double2 p = pix[e];
    vis_edge[pos] = e;
    if (vis_pix) vis_pix[pos] = p;
    t_xy[pos] = make_float2((float)p.x, (float)p.y);
    const uint4 *d = reinterpret_cast<const uint4 *>(map.desc + 32 * (size_t)slot);
    t_desc[2 * (size_t)pos] = d[0];
    t_desc[2 * (size_t)pos + 1] = d[1];
    const double *Xp = map.pt3d + 3 * (size_t)slot;
    vis_pt3d[3 * (size_t)pos] = Xp[0];
    vis_pt3d[3 * (size_t)pos + 1] = Xp[1];
    vis_pt3d[3 * (size_t)pos + 2] = Xp[2];
}

// pass 4: everything the caller gets, written to pinned host memory by ONE kernel (a kernel that writes host
// memory only completes when its PCIe writes have drained, ~10 us: paying that once beats paying it in every
// stage): header, visible edges (+ pixels), match list, and the matched (3-D point, frame pixel, edge) triples
// gathered contiguously (slam/tracking.py:126-128).
struct TrackHostOut {
    int32_t *hdr;                 // n_visible, n_matches
    int32_t *vis_edge;
    double2 *vis_pix;             // nullable
    int32_t *m_query, *m_train, *m_dist, *m_edge;
    double *m_pts3d, *m_kp;
};
__global__ void __launch_bounds__(LM_NT) lm_gather_kernel(const int32_t *hdr, const int32_t *vis_edge, const double2 *vis_pix,
                                                          const int32_t *m_query, const int32_t *m_train, const int32_t *m_dist,
                                                          const double *vis_pt3d, const double *q_kp, int32_t have_matches,
                                                          TrackHostOut out) {
    const int nv = hdr[0], nm = have_matches ? hdr[1] : 0;
    const int i = blockIdx.x * LM_NT + threadIdx.x;
    if (i == 0) { out.hdr[0] = nv; out.hdr[1] = nm; }
    for (int e = i; e < nv; e += gridDim.x * LM_NT) {
        out.vis_edge[e] = vis_edge[e];
        if (out.vis_pix) out.vis_pix[e] = vis_pix[e];
    }
    for (int k = i; k < nm; k += gridDim.x * LM_NT) {
        const int qi = m_query[k], ti = m_train[k];
        out.m_query[k] = qi;
        out.m_train[k] = ti;
        out.m_dist[k] = m_dist[k];
        out.m_edge[k] = vis_edge[ti];
        out.m_pts3d[3 * (size_t)k] = vis_pt3d[3 * (size_t)ti];
        out.m_pts3d[3 * (size_t)k + 1] = vis_pt3d[3 * (size_t)ti + 1];
        out.m_pts3d[3 * (size_t)k + 2] = vis_pt3d[3 * (size_t)ti + 2];
        out.m_kp[2 * (size_t)k] = q_kp[2 * (size_t)qi];
        out.m_kp[2 * (size_t)k + 1] = q_kp[2 * (size_t)qi + 1];
    }
}

// ---- keyframe vote (reference slam/tracking.py:154) -------------------------------------
// Counter(ids_matching_kfs[inds[inliers]]).most_common(top): the keyframes that own the map points the pose
// optimisation kept, most frequent first, ties in order of first appearance (Counter.most_common sorts stably by
// count).  One CTA: the (keyframe id, position) pairs of the inliers are sorted in shared memory (bitonic), runs
// of equal ids become (count, first position) records, and those are sorted again by count.
constexpr int VOTE_NT = 1024, VOTE_MAX = 4096;
constexpr size_t VOTE_SMEM = (size_t)VOTE_MAX * 12;

template <typename T>
__device__ __forceinline__ void vote_bitonic(T *a, const int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += VOTE_NT) {
                const int l = i ^ j;
                if (l > i) {
                    const T x = a[i], y = a[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(VOTE_NT) lm_vote_kernel(const int32_t *m_train, const int32_t *vis_edge, const int32_t *edge_kf,
                                                          const int32_t *inliers, const int32_t n_inliers, const int32_t n_matches,
                                                          const int32_t top, int32_t *out /* [0] n, [1 .. top] ids, [1 + top ..] counts */) {
    extern __shared__ __align__(16) unsigned char s_vote[];   // VOTE_SMEM bytes (more than 48 KB: opted in by the host)
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(s_vote);
    uint32_t *s_run = reinterpret_cast<uint32_t *>(s_vote + (size_t)VOTE_MAX * 8);
    __shared__ int s_n;
    int n_pow2 = 1;
    while (n_pow2 < n_inliers) n_pow2 <<= 1;
    if (threadIdx.x == 0) s_n = 0;
    for (int i = threadIdx.x; i < n_pow2; i += VOTE_NT) {
        unsigned long long key = ~0ull;
        if (i < n_inliers) {
            const int k = inliers[i];
            if (k >= 0 && k < n_matches) {
                const uint32_t kf = (uint32_t)edge_kf[vis_edge[m_train[k]]] ^ 0x80000000u;   // signed order
                key = ((unsigned long long)kf << 32) | (uint32_t)i;
            }
        }
        s_key[i] = key;
    }
    __syncthreads();
    vote_bitonic(s_key, n_pow2);
    // runs of equal keyframe ids: (count, first position, keyframe) - the first element of a run holds its lowest position
    for (int i = threadIdx.x; i < n_pow2; i += VOTE_NT) {
        const unsigned long long key = s_key[i];
        if (key != ~0ull && (i == 0 || (s_key[i - 1] >> 32) != (key >> 32))) {
            int e = i + 1;
            while (e < n_pow2 && s_key[e] != ~0ull && (s_key[e] >> 32) == (key >> 32)) ++e;   // (the padding sorts behind keyframe INT_MAX)
            const int r = atomicAdd(&s_n, 1);
            // most frequent first, then first appearance (a position < 4096: 12 bits)
            s_run[r] = ((uint32_t)(VOTE_MAX - (e - i)) << 12) | (uint32_t)(key & 0xFFFu);
        }
    }
    __syncthreads();
    const int n_runs = s_n;
    int r_pow2 = 1;
    while (r_pow2 < n_runs) r_pow2 <<= 1;
    for (int i = n_runs + threadIdx.x; i < r_pow2; i += VOTE_NT) s_run[i] = 0xFFFFFFFFu;
    __syncthreads();
    vote_bitonic(s_run, r_pow2);
    const int n_out = min(n_runs, top);
    if (threadIdx.x == 0) out[0] = n_out;
    for (int r = threadIdx.x; r < n_out; r += VOTE_NT) {
        const uint32_t rec = s_run[r];
        out[1 + r] = edge_kf[vis_edge[m_train[inliers[rec & 0xFFFu]]]];   // the keyframe of the run's first inlier
        out[1 + top + r] = VOTE_MAX - (int32_t)(rec >> 12);
    }
}

// scatter of an update batch into the store (slots may repeat: last writer in batch order wins is NOT
// guaranteed, callers pass each slot once per call)
__global__ void lm_scatter_kernel(int32_t n, const int32_t *slots, int32_t capacity, const uint4 *desc, const double *pt3d,
                                  const double *normal, uint4 *s_desc, double *s_pt3d, double *s_normal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slots[i];
    if (s < 0 || s >= capacity) return;
    if (desc) { s_desc[2 * (size_t)s] = desc[2 * (size_t)i]; s_desc[2 * (size_t)s + 1] = desc[2 * (size_t)i + 1]; }
    if (pt3d) for (int c = 0; c < 3; ++c) s_pt3d[3 * (size_t)s + c] = pt3d[3 * (size_t)i + c];
    if (normal) for (int c = 0; c < 3; ++c) s_normal[3 * (size_t)s + c] = normal[3 * (size_t)i + c];
}

// ---- representative descriptor of a map point ---------------------------------
// reference slam/nodes.py:146-153 (MapPoint.add_observation): with the point's n <= 10 stored
// observations, D[i][j] = Hamming(obs_i, obs_j) (diagonal 0), feat = obs[argmin_j median_i D[i][j]],
// np.argmin -> lowest j on ties, np.median -> mean of the two middle values for even n.
// One 16-lane group per map point: lane j holds observation j, the descriptors go round by shuffle,
// every lane ranks its own column and the group takes the arg-min of 2 * median (an exact integer).
constexpr int REP_MAX_OBS = 16;

__global__ void __launch_bounds__(256) rep_select_kernel(const uint4 *obs, const int32_t *counts, int32_t n_points, int32_t max_obs,
                                                         int32_t *out_idx) {
    const int gid = (blockIdx.x * 256 + threadIdx.x) >> 4;       // map point
    const int j = threadIdx.x & 15;                              // observation / column
    const uint32_t gmask = 0xFFFFu << (threadIdx.x & 16);        // the 16 lanes of this group
    const bool live = gid < n_points;
    const int n = live ? min(max(counts[gid], 0), max_obs) : 0;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (j < n) {
        const uint4 a = obs[((size_t)gid * max_obs + j) * 2], b = obs[((size_t)gid * max_obs + j) * 2 + 1];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
    int d[REP_MAX_OBS];
#pragma unroll
    for (int i = 0; i < REP_MAX_OBS; ++i) {
        int dist = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) dist += __popc(w[c] ^ __shfl_sync(gmask, w[c], (threadIdx.x & 16) + i, 32));
        d[i] = dist;   // rows i >= n are ignored below
    }
    // 2 * median of d[0..n): the elements of rank (n-1)/2 and n/2 in the stable order
    const int r_lo = (n - 1) >> 1, r_hi = n >> 1;
    int med2 = 0;
#pragma unroll
    for (int a = 0; a < REP_MAX_OBS; ++a) {
        if (a < n) {
            int rank = 0;
#pragma unroll
            for (int b = 0; b < REP_MAX_OBS; ++b)
                if (b < n && (d[b] < d[a] || (d[b] == d[a] && b < a))) ++rank;
            if (rank == r_lo) med2 += d[a];
            if (rank == r_hi) med2 += d[a];
        }
    }
    uint32_t key = (j < n) ? (((uint32_t)med2 << 8) | (uint32_t)j) : 0xFFFFFFFFu;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(gmask, key, o, 32));
    if (live && j == 0) out_idx[gid] = n > 0 ? (int32_t)(key & 0xFFu) : -1;
}
