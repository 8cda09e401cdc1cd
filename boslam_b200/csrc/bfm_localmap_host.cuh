// bfm_localmap_host.cuh - host side of the local-map store and the fused tracking call.
// (included by bfm_api.cu at global scope, after the anonymous namespace)

struct bfm_map_s {
    bfm_handle_t h = nullptr;
    int32_t capacity = 0;
    uint8_t *desc = nullptr;   // device [capacity][32]
    double *pt3d = nullptr;    // device [capacity][3]
    double *normal = nullptr;  // device [capacity][3]
    DevBuf work;               // per-call device workspace
    void *h_in = nullptr, *h_out = nullptr;   // pinned staging
    size_t h_in_cap = 0, h_out_cap = 0;
    int32_t last_visible = 0;  // visible edges of the previous call: the planning hint for the next one
    // device views of the last bfm_track_local_map call (valid until the next call that uses the map's workspace):
    // what bfm_keyframe_vote reads
    const int32_t *last_m_train = nullptr, *last_vis_edge = nullptr;
    int32_t last_matches = -1, last_edges = 0;
};

namespace {

int ensure_pinned(bfm_handle_t h, void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return BFM_OK;
    if (*p) CU_TRY(h, cudaFreeHost(*p));
    *p = nullptr;
    *cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    CU_TRY(h, cudaMallocHost(p, want));
    *cap = want;
    return BFM_OK;
}

struct Carver {   // bump allocator over one block, 256-byte aligned pieces
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base(static_cast<char *>(b)) {}
    template <typename T>
    T *take(size_t n) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off = align256(off + n * sizeof(T));
        return p;
    }
};

}  // namespace

extern "C" {

int bfm_map_create(bfm_handle_t h, int32_t capacity, bfm_map_t *out) {
    if (!h || !out) return BFM_ERR_INVALID;
    *out = nullptr;
    h->err.clear();
    if (capacity <= 0) return fail(h, BFM_ERR_INVALID, "capacity must be positive");
    CU_TRY(h, cudaSetDevice(h->device));
    bfm_map_t m = new bfm_map_s();
    m->h = h;
    m->capacity = capacity;
    cudaError_t e = cudaMalloc(&m->desc, (size_t)capacity * 32);
    if (e == cudaSuccess) e = cudaMalloc(&m->pt3d, (size_t)capacity * 24);
    if (e == cudaSuccess) e = cudaMalloc(&m->normal, (size_t)capacity * 24);
    if (e == cudaSuccess) e = cudaMemset(m->desc, 0, (size_t)capacity * 32);
    if (e == cudaSuccess) e = cudaMemset(m->pt3d, 0, (size_t)capacity * 24);
    if (e == cudaSuccess) e = cudaMemset(m->normal, 0, (size_t)capacity * 24);
    if (e != cudaSuccess) {
        bfm_map_destroy(m);
        return fail(h, BFM_ERR_NOMEM, std::string("map store allocation: ") + cudaGetErrorString(e));
    }
    *out = m;
    return BFM_OK;
}

int bfm_map_destroy(bfm_map_t m) {
    if (!m) return BFM_OK;
    cudaSetDevice(m->h->device);
    cudaDeviceSynchronize();
    if (m->desc) cudaFree(m->desc);
    if (m->pt3d) cudaFree(m->pt3d);
    if (m->normal) cudaFree(m->normal);
    if (m->work.p) cudaFree(m->work.p);
    if (m->h_in) cudaFreeHost(m->h_in);
    if (m->h_out) cudaFreeHost(m->h_out);
    delete m;
    return BFM_OK;
}

int bfm_map_update(bfm_map_t m, int32_t n, const int32_t *slots, const uint8_t *desc, const double *pt3d,
                   const double *normal) {
    if (!m) return BFM_ERR_INVALID;
    bfm_handle_t h = m->h;
    h->err.clear();
    if (n < 0 || (n > 0 && !slots)) return fail(h, BFM_ERR_INVALID, "bad update batch");
    if (n == 0) return BFM_OK;
    m->last_matches = -1;   // the workspace is reused
    for (int i = 0; i < n; ++i)
        if (slots[i] < 0 || slots[i] >= m->capacity) return fail(h, BFM_ERR_INVALID, "slot " + std::to_string(slots[i]) + " is outside the store");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    Carver sz(nullptr);
    sz.take<int32_t>(n);
    if (desc) sz.take<uint8_t>((size_t)n * 32);
    if (pt3d) sz.take<double>((size_t)n * 3);
    if (normal) sz.take<double>((size_t)n * 3);
    int rc = ensure_pinned(h, &m->h_in, &m->h_in_cap, sz.off);
    if (rc) return rc;
    rc = ensure(h, m->work, sz.off);
    if (rc) return rc;
    Carver hc(m->h_in), dc(m->work.p);
    int32_t *h_slots = hc.take<int32_t>(n), *d_slots = dc.take<int32_t>(n);
    std::memcpy(h_slots, slots, (size_t)n * 4);
    uint8_t *d_desc = nullptr;
    double *d_pt = nullptr, *d_n = nullptr;
    if (desc) { std::memcpy(hc.take<uint8_t>((size_t)n * 32), desc, (size_t)n * 32); d_desc = dc.take<uint8_t>((size_t)n * 32); }
    if (pt3d) { std::memcpy(hc.take<double>((size_t)n * 3), pt3d, (size_t)n * 24); d_pt = dc.take<double>((size_t)n * 3); }
    if (normal) { std::memcpy(hc.take<double>((size_t)n * 3), normal, (size_t)n * 24); d_n = dc.take<double>((size_t)n * 3); }
    CU_TRY(h, cudaMemcpyAsync(m->work.p, m->h_in, hc.off, cudaMemcpyHostToDevice, st));
    lm_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_slots, m->capacity, reinterpret_cast<const uint4 *>(d_desc), d_pt, d_n,
                                                       reinterpret_cast<uint4 *>(m->desc), m->pt3d, m->normal);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaStreamSynchronize(st));   // the pinned staging block is reused by the next call
    h->launches += 1;
    return BFM_OK;
}

int bfm_track_local_map(bfm_map_t m, const bfm_track_params_t *tp, const int32_t *edges, int32_t n_edges,
                        const uint8_t *q_desc, const double *q_kp, int32_t nq, const bfm_options_t *opts,
                        int32_t *visible_edges, double *visible_pixels, int32_t *m_query, int32_t *m_train,
                        int32_t *m_dist, int32_t *m_edge, double *m_pts3d, double *m_kp, int32_t *n_visible,
                        int32_t *n_matches) {
    if (!m) return BFM_ERR_INVALID;
    bfm_handle_t h = m->h;
    h->err.clear();
    if (!tp || !opts || !n_visible || !n_matches) return fail(h, BFM_ERR_INVALID, "NULL argument");
    if (n_edges < 0 || nq < 0 || (n_edges > 0 && !edges) || (nq > 0 && (!q_desc || !q_kp)))
        return fail(h, BFM_ERR_INVALID, "bad sizes or NULL inputs");
    if (n_edges >= BFM_MAX_TRAIN_ROWS || nq >= BFM_MAX_QUERY_ROWS) return fail(h, BFM_ERR_INVALID, "more than 2^22 - 1 rows");
    if (opts->mask_kind == BFM_MASK_DENSE) return fail(h, BFM_ERR_INVALID, "a dense mask is not meaningful here: use the window");
    if (opts->k < 1 || opts->k > 2) return fail(h, BFM_ERR_INVALID, "k must be 1 or 2 on the tracking path");
    if (opts->cross_check && opts->k != 1) return fail(h, BFM_ERR_INVALID, "cross_check requires k == 1 (cv2 asserts the same)");
    if (opts->cross_check && opts->ratio >= 0) return fail(h, BFM_ERR_INVALID, "cross_check and ratio are exclusive");
    if (opts->mask_kind != BFM_MASK_NONE && opts->mask_kind != BFM_MASK_WINDOW) return fail(h, BFM_ERR_INVALID, "bad mask_kind");
    *n_visible = 0;
    *n_matches = 0;
    m->last_matches = -1;
    if (n_edges == 0) return BFM_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const bool trace = std::getenv("BFM_TRACE") != nullptr;
    const auto cpu0 = std::chrono::steady_clock::now();
    auto cpu_us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - cpu0).count(); };
    const int nblk = (n_edges + LM_NT - 1) / LM_NT;
    const int nqa = std::max(nq, 1);

    // ---- layout: device block [inputs | scratch]; pinned host blocks for the inputs (one H2D copy) and for the
    //      outputs, which the kernels write directly (zero-copy over PCIe: only what exists is transferred)
    Carver din(nullptr);
    din.take<int32_t>(n_edges);
    din.take<uint8_t>((size_t)nqa * 32);
    din.take<double>((size_t)nqa * 2);
    const size_t in_bytes = din.off;
    Carver hout(nullptr);
    hout.take<int32_t>(4);                        // n_visible, m_count
    hout.take<int32_t>(n_edges);                  // vis_edge
    hout.take<double2>(visible_pixels ? n_edges : 0);   // vis_pix
    hout.take<int32_t>((size_t)nqa * 3);          // m_query | m_train | m_dist
    hout.take<int32_t>(nqa);                      // m_edge
    hout.take<double>((size_t)nqa * 3);           // m_pts3d
    hout.take<double>((size_t)nqa * 2);           // m_kp
    const size_t out_bytes = hout.off;
    Carver dsc(nullptr);
    dsc.take<int32_t>(4);                         // header: n_visible, m_count
    dsc.take<double2>(n_edges);                   // pix
    dsc.take<uint8_t>(n_edges);                   // flag
    dsc.take<int32_t>(nblk);                      // block_count
    dsc.take<int32_t>(n_edges);                   // vis_edge
    dsc.take<float2>(n_edges);                    // t_xy
    dsc.take<uint4>((size_t)n_edges * 2);         // t_desc
    dsc.take<double>((size_t)n_edges * 3);        // vis_pt3d
    dsc.take<float2>(nqa);                        // q_xy
    dsc.take<int32_t>((size_t)nqa * 3);           // m_query | m_train | m_dist
    dsc.take<double2>(visible_pixels ? n_edges : 0);   // vis_pix
    const size_t scratch_bytes = dsc.off;
    int rc = ensure(h, m->work, in_bytes + scratch_bytes);
    if (rc) return rc;
    rc = ensure_pinned(h, &m->h_in, &m->h_in_cap, in_bytes);
    if (rc) return rc;
    rc = ensure_pinned(h, &m->h_out, &m->h_out_cap, out_bytes);
    if (rc) return rc;
    char *dbase = static_cast<char *>(m->work.p);
    Carver di(dbase), hi(m->h_in), ds(dbase + in_bytes), ho(m->h_out);
    int32_t *d_edges = di.take<int32_t>(n_edges);
    uint8_t *d_q = di.take<uint8_t>((size_t)nqa * 32);
    double *d_kp = di.take<double>((size_t)nqa * 2);
    std::memcpy(hi.take<int32_t>(n_edges), edges, (size_t)n_edges * 4);
    uint8_t *hq = hi.take<uint8_t>((size_t)nqa * 32);
    double *hkp = hi.take<double>((size_t)nqa * 2);
    if (nq) {
        std::memcpy(hq, q_desc, (size_t)nq * 32);
        std::memcpy(hkp, q_kp, (size_t)nq * 16);
    }
    int32_t *d_hdr = ds.take<int32_t>(4);
    double2 *d_pix = ds.take<double2>(n_edges);
    uint8_t *d_flag = ds.take<uint8_t>(n_edges);
    int32_t *d_bc = ds.take<int32_t>(nblk);
    int32_t *d_vedge = ds.take<int32_t>(n_edges);
    float2 *d_txy = ds.take<float2>(n_edges);
    uint4 *d_tdesc = ds.take<uint4>((size_t)n_edges * 2);
    double *d_vpt = ds.take<double>((size_t)n_edges * 3);
    float2 *d_qxy = ds.take<float2>(nqa);
    int32_t *d_m = ds.take<int32_t>((size_t)nqa * 3);
    double2 *d_vpix = ds.take<double2>(visible_pixels ? n_edges : 0);
    if (!visible_pixels) d_vpix = nullptr;
    int32_t *h_hdr = ho.take<int32_t>(4);
    int32_t *h_vedge = ho.take<int32_t>(n_edges);
    double2 *h_vpix = ho.take<double2>(visible_pixels ? n_edges : 0);
    if (!visible_pixels) h_vpix = nullptr;
    int32_t *h_m = ho.take<int32_t>((size_t)nqa * 3);
    int32_t *h_medge = ho.take<int32_t>(nqa);
    double *h_mpt = ho.take<double>((size_t)nqa * 3);
    double *h_mkp = ho.take<double>((size_t)nqa * 2);
    h_hdr[0] = h_hdr[1] = 0;

    const double t_staged = cpu_us();
    CU_TRY(h, cudaMemcpyAsync(dbase, m->h_in, in_bytes, cudaMemcpyHostToDevice, st));

    ProjectParams pp;
    pp.qw = tp->q[0]; pp.qx = tp->q[1]; pp.qy = tp->q[2]; pp.qz = tp->q[3];
    pp.tx = tp->t[0]; pp.ty = tp->t[1]; pp.tz = tp->t[2];
    pp.sx = tp->see_vector[0]; pp.sy = tp->see_vector[1]; pp.sz = tp->see_vector[2];
    pp.fx = tp->fx; pp.fy = tp->fy; pp.cx = tp->cx; pp.cy = tp->cy;
    pp.width = (double)tp->width; pp.height = (double)tp->height;
    pp.cos_max = tp->cos_max;
    MapView mv{m->desc, m->pt3d, m->normal};
    const bool window = opts->mask_kind == BFM_MASK_WINDOW && nq > 0;
    lm_project_kernel<<<nblk, LM_NT, 0, st>>>(pp, mv, d_edges, n_edges, m->capacity, d_pix, d_flag, d_bc, d_kp,
                                              window ? d_qxy : nullptr, nq);
    CU_TRY(h, cudaGetLastError());
    lm_compact_kernel<<<nblk, LM_NT, 0, st>>>(mv, d_edges, n_edges, d_pix, d_flag, d_bc, d_vedge, d_vpix, d_txy, d_tdesc, d_vpt, d_hdr);
    CU_TRY(h, cudaGetLastError());
    int kernels = 2;
    if (nq > 0) {
        bfm_options_t od = *opts;
        if (window) {
            od.q_xy = reinterpret_cast<const float *>(d_qxy);
            od.t_xy = reinterpret_cast<const float *>(d_txy);
        }
        // the train set is the compacted survivor list: its size lives in d_hdr[0]; the plan covers
        // the rows the previous frame saw and every work item re-cuts what exists on the device (no host round trip)
        const bfm_problem_t pr = {0, nq, 0, n_edges, 0, 0};
        const bfm_outputs_t outs = {nullptr, nullptr, d_m, d_m + nqa, d_m + 2 * (size_t)nqa, d_hdr + 1, 0, 0};
        rc = run_device(h, d_q, nq, reinterpret_cast<const uint8_t *>(d_tdesc), n_edges, &pr, 1, nq, &od, &outs, 1, st, nullptr, d_hdr,
                        m->last_visible);
        if (rc) return rc;
        kernels += h->info.kernels_launched;
    }
    TrackHostOut ho_ptrs{h_hdr, h_vedge, h_vpix, h_m, h_m + nqa, h_m + 2 * (size_t)nqa, h_medge, h_mpt, h_mkp};
    const int ggrid = std::min(64, (std::max(n_edges, nq) + LM_NT - 1) / LM_NT);
    lm_gather_kernel<<<ggrid, LM_NT, 0, st>>>(d_hdr, d_vedge, d_vpix, d_m, d_m + nqa, d_m + 2 * (size_t)nqa, d_vpt, d_kp, nq > 0 ? 1 : 0,
                                              ho_ptrs);
    CU_TRY(h, cudaGetLastError());
    ++kernels;
    const double t_queued = cpu_us();
    CU_TRY(h, cudaStreamSynchronize(st));
    const double t_synced = cpu_us();
    h->launches += kernels - (nq > 0 ? h->info.kernels_launched : 0);
    h->info.kernels_launched = kernels;

    const int nv = h_hdr[0], nm = h_hdr[1];
    *n_visible = nv;
    *n_matches = nm;
    m->last_visible = nv;   // tracking is temporally coherent: the next frame's plan is sized for about this many rows
    m->last_m_train = d_m + nqa;
    m->last_vis_edge = d_vedge;
    m->last_matches = nm;
    m->last_edges = n_edges;
    if (visible_edges) std::memcpy(visible_edges, h_vedge, (size_t)nv * 4);
    if (visible_pixels) std::memcpy(visible_pixels, h_vpix, (size_t)nv * 16);
    if (m_query) std::memcpy(m_query, h_m, (size_t)nm * 4);
    if (m_train) std::memcpy(m_train, h_m + nqa, (size_t)nm * 4);
    if (m_dist) std::memcpy(m_dist, h_m + 2 * (size_t)nqa, (size_t)nm * 4);
    if (m_edge) std::memcpy(m_edge, h_medge, (size_t)nm * 4);
    if (m_pts3d) std::memcpy(m_pts3d, h_mpt, (size_t)nm * 24);
    if (m_kp) std::memcpy(m_kp, h_mkp, (size_t)nm * 16);
    if (trace)
        std::fprintf(stderr, "[bfm trace] track_local_map: inputs staged %.1f us, %d kernels + copies queued %.1f us, synced %.1f us, "
                     "outputs copied %.1f us (in %zu B, out %zu B)\n", t_staged, kernels, t_queued, t_synced, cpu_us(), in_bytes, out_bytes);
    return BFM_OK;
}

int bfm_keyframe_vote(bfm_map_t m, const int32_t *edge_kf, int32_t n_edges, const int32_t *inliers, int32_t n_inliers,
                      int32_t top, int32_t *kf_ids, int32_t *kf_counts, int32_t *n_kfs) {
    if (!m) return BFM_ERR_INVALID;
    bfm_handle_t h = m->h;
    h->err.clear();
    if (!n_kfs || top < 0 || n_inliers < 0 || (n_inliers > 0 && (!inliers || !edge_kf)) || (top > 0 && (!kf_ids || !kf_counts)))
        return fail(h, BFM_ERR_INVALID, "NULL argument or negative size");
    *n_kfs = 0;
    if (m->last_matches < 0) return fail(h, BFM_ERR_INVALID, "bfm_keyframe_vote follows bfm_track_local_map on the same map (no call in between)");
    if (n_edges != m->last_edges) return fail(h, BFM_ERR_INVALID, "edge_kf must have one entry per edge of the tracking call");
    if (n_inliers > VOTE_MAX) return fail(h, BFM_ERR_UNSUPPORTED, "more than 4096 inliers");
    if (n_inliers == 0 || top == 0) return BFM_OK;
    top = std::min(top, n_inliers);
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    // inputs behind the tracking call's scratch (which holds the match list): one H2D copy, one kernel, one D2H copy
    Carver in(nullptr);
    in.take<int32_t>(n_edges);
    in.take<int32_t>(n_inliers);
    const size_t in_bytes = in.off, out_bytes = align256((size_t)(1 + 2 * top) * 4);
    DevBuf &vb = h->lower;   // (free on this path: the tracking call is k <= 2)
    int rc = ensure(h, vb, in_bytes + out_bytes);
    if (rc) return rc;
    rc = ensure_pinned(h, &m->h_in, &m->h_in_cap, in_bytes + out_bytes);
    if (rc) return rc;
    Carver hi(m->h_in), di(vb.p);
    int32_t *h_kf = hi.take<int32_t>(n_edges), *d_kf = di.take<int32_t>(n_edges);
    int32_t *h_inl = hi.take<int32_t>(n_inliers), *d_inl = di.take<int32_t>(n_inliers);
    std::memcpy(h_kf, edge_kf, (size_t)n_edges * 4);
    std::memcpy(h_inl, inliers, (size_t)n_inliers * 4);
    int32_t *d_out = reinterpret_cast<int32_t *>(static_cast<char *>(vb.p) + in_bytes);
    int32_t *h_out = reinterpret_cast<int32_t *>(static_cast<char *>(m->h_in) + in_bytes);
    CU_TRY(h, cudaMemcpyAsync(vb.p, m->h_in, in_bytes, cudaMemcpyHostToDevice, st));
    static bool opted_in = false;   // 48 KB of sort keys + 16 KB of run records
    if (!opted_in) {
        CU_TRY(h, cudaFuncSetAttribute(lm_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VOTE_SMEM));
        opted_in = true;
    }
    lm_vote_kernel<<<1, VOTE_NT, VOTE_SMEM, st>>>(m->last_m_train, m->last_vis_edge, d_kf, d_inl, n_inliers, m->last_matches, top, d_out);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(h_out, d_out, (size_t)(1 + 2 * top) * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    h->launches += 1;
    h->info = bfm_launch_info_t{};
    h->info.kernels_launched = 1;
    const int n = h_out[0];
    *n_kfs = n;
    std::memcpy(kf_ids, h_out + 1, (size_t)n * 4);
    std::memcpy(kf_counts, h_out + 1 + top, (size_t)n * 4);
    return BFM_OK;
}

int bfm_select_representative(bfm_handle_t h, const uint8_t *obs, const int32_t *counts, int32_t n_points,
                              int32_t max_obs, int32_t *out_idx) {
    if (!h) return BFM_ERR_INVALID;
    h->err.clear();
    if (n_points < 0 || max_obs < 1 || max_obs > REP_MAX_OBS) return fail(h, BFM_ERR_INVALID, "max_obs must be 1..16");
    if (n_points == 0) return BFM_OK;
    if (!obs || !counts || !out_idx) return fail(h, BFM_ERR_INVALID, "NULL argument");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const size_t ob = (size_t)n_points * max_obs * 32, cb = (size_t)n_points * 4;
    const size_t o_c = align256(ob), o_out = align256(o_c + cb), total = align256(o_out + cb);
    int rc = ensure(h, h->d_in, total);
    if (rc) return rc;
    if (h->h_out_cap < total) {
        if (h->h_out) CU_TRY(h, cudaFreeHost(h->h_out));
        h->h_out = nullptr;
        h->h_out_cap = 0;
        CU_TRY(h, cudaMallocHost(&h->h_out, total + total / 4 + 4096));
        h->h_out_cap = total + total / 4 + 4096;
    }
    char *din = static_cast<char *>(h->d_in.p), *hs = static_cast<char *>(h->h_out);
    std::memcpy(hs, obs, ob);
    std::memcpy(hs + o_c, counts, cb);
    CU_TRY(h, cudaMemcpyAsync(din, hs, o_c + cb, cudaMemcpyHostToDevice, st));
    const int groups_per_block = 256 / 16;
    rep_select_kernel<<<(n_points + groups_per_block - 1) / groups_per_block, 256, 0, st>>>(
        reinterpret_cast<const uint4 *>(din), reinterpret_cast<const int32_t *>(din + o_c), n_points, max_obs,
        reinterpret_cast<int32_t *>(din + o_out));
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(hs + o_out, din + o_out, cb, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    std::memcpy(out_idx, hs + o_out, cb);
    h->launches += 1;
    h->info = bfm_launch_info_t{};
    h->info.kernels_launched = 1;
    return BFM_OK;
}

}  // extern "C"
