// bfm_pipeline.cuh - pipelined BFM_MEM_HOST path for large keyframe-pair batches.
// (included by bfm_api.cu inside its anonymous namespace, after run_device / run_host)
//
// The batch is planned once (one table upload), then its P problems run as G launch groups of
// growing size.  Three streams overlap
//     in_stream   H2D of the descriptor rows group g+1 needs
//     stream      scan + finalize of group g             (waits on the group's copy-in event)
//     out_stream  D2H of group g-1's result rows         (waits on the group's finalize event)
// so the wall time tends to max(copy, compute) instead of their sum.  The first group is small
// (compute starts after a short copy) and groups double in size (fewer, more efficient launches).
// Eligible layouts are the packed ones the Python layer produces: q_begin / t_begin / out_begin
// non-decreasing in p, so "rows uploaded so far" is a single watermark per array.  Pinned caller
// buffers are copied directly (cudaPointerGetAttributes); pageable outputs go through the
// handle's pinned staging.  BFM_TRACE=1 in the environment prints the per-group timeline.

constexpr int MAX_CHUNKS = 16;

bool host_ptr_is_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

bool pipeline_eligible(bfm_handle_t h, int32_t nq_rows, int32_t nt_rows, const bfm_problem_t *problems,
                       int32_t n_problems, const bfm_options_t *o) {
    if (h->pipeline_chunks == 1 || n_problems < 4 || o->mask_kind == BFM_MASK_DENSE) return false;
    if (h->pipeline_chunks == 0 && ((size_t)nq_rows + (size_t)nt_rows) * 32 < ((size_t)4 << 20)) return false;
    for (int p = 1; p < n_problems; ++p)
        if (problems[p].q_begin < problems[p - 1].q_begin || problems[p].t_begin < problems[p - 1].t_begin ||
            problems[p].out_begin < problems[p - 1].out_begin + problems[p - 1].q_count)
            return false;
    return true;
}

struct PipeCtx {
    bfm_handle_t h;
    const bfm_problem_t *problems;
    const int *bounds;
    int32_t n_out_rows;
    int k;
    // device / host views of the output block
    int32_t *d_ki, *d_kd, *d_mq, *d_mt, *d_md, *d_mc;
    int32_t *h_ki, *h_kd, *h_mq, *h_mt, *h_md, *h_mc;
    // copy-in state: rows [0, q_mark) / [0, t_mark) of the query / train arrays are queued
    const uint8_t *q, *t;
    const float *q_xy, *t_xy;  // host pixel coordinates (window) or null
    char *din;
    size_t o_q, o_t, o_qxy, o_txy;
    int32_t q_mark, t_mark;
    int n_groups, copied_groups;
    bool trace;
    std::chrono::steady_clock::time_point cpu0;
    double cpu_ms[MAX_CHUNKS][2];
};

// queue the H2D copies of every row group g needs and has not been queued yet, then its event
int pipe_copy_in(PipeCtx *c, int g) {
    bfm_handle_t h = c->h;
    cudaStream_t sin = h->in_stream;
    int32_t q_need = c->q_mark, t_need = c->t_mark;
    for (int p = c->bounds[g]; p < c->bounds[g + 1]; ++p) {
        q_need = std::max(q_need, c->problems[p].q_begin + c->problems[p].q_count);
        t_need = std::max(t_need, c->problems[p].t_begin + c->problems[p].t_count);
    }
    if (q_need > c->q_mark) {
        const size_t b = (size_t)c->q_mark, n = (size_t)(q_need - c->q_mark);
        CU_TRY(h, cudaMemcpyAsync(c->din + c->o_q + b * 32, c->q + b * 32, n * 32, cudaMemcpyHostToDevice, sin));
        if (c->q_xy) CU_TRY(h, cudaMemcpyAsync(c->din + c->o_qxy + b * 8, c->q_xy + b * 2, n * 8, cudaMemcpyHostToDevice, sin));
        c->q_mark = q_need;
    }
    if (t_need > c->t_mark) {
        const size_t b = (size_t)c->t_mark, n = (size_t)(t_need - c->t_mark);
        CU_TRY(h, cudaMemcpyAsync(c->din + c->o_t + b * 32, c->t + b * 32, n * 32, cudaMemcpyHostToDevice, sin));
        if (c->t_xy) CU_TRY(h, cudaMemcpyAsync(c->din + c->o_txy + b * 8, c->t_xy + b * 2, n * 8, cudaMemcpyHostToDevice, sin));
        c->t_mark = t_need;
    }
    CU_TRY(h, cudaEventRecord(h->chunk_ev[2 * g], sin));
    return BFM_OK;
}

// group g may start once its descriptor rows have landed.  Copies are queued here (two groups ahead
// of the compute) rather than up front so that run_device's table upload, which shares the H2D
// engine, is not stuck behind the whole batch.
int pipe_before(void *vctx, int g) {
    PipeCtx *c = static_cast<PipeCtx *>(vctx);
    while (c->copied_groups < std::min(g + 3, c->n_groups)) {
        const int rc = pipe_copy_in(c, c->copied_groups);
        if (rc) return rc;
        ++c->copied_groups;
    }
    CU_TRY(c->h, cudaStreamWaitEvent(c->h->stream, c->h->chunk_ev[2 * g], 0));
    if (c->trace) c->cpu_ms[g][0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c->cpu0).count();
    return BFM_OK;
}

// group g finished: ship its result rows while the next group computes
int pipe_after(void *vctx, int g) {
    PipeCtx *c = static_cast<PipeCtx *>(vctx);
    bfm_handle_t h = c->h;
    cudaStream_t sout = h->out_stream;
    CU_TRY(h, cudaEventRecord(h->chunk_ev[2 * g + 1], h->stream));
    CU_TRY(h, cudaStreamWaitEvent(sout, h->chunk_ev[2 * g + 1], 0));
    const int p0 = c->bounds[g], p1 = c->bounds[g + 1];
    int32_t out_lo = c->n_out_rows, out_hi = 0;
    for (int p = p0; p < p1; ++p)
        if (c->problems[p].q_count > 0) {
            out_lo = std::min(out_lo, c->problems[p].out_begin);
            out_hi = std::max(out_hi, c->problems[p].out_begin + c->problems[p].q_count);
        }
    if (out_hi > out_lo) {
        const size_t rows = (size_t)(out_hi - out_lo), k = (size_t)c->k;
        if (c->d_ki) {
            CU_TRY(h, cudaMemcpyAsync(c->h_ki + out_lo * k, c->d_ki + out_lo * k, rows * k * 4, cudaMemcpyDeviceToHost, sout));
            CU_TRY(h, cudaMemcpyAsync(c->h_kd + out_lo * k, c->d_kd + out_lo * k, rows * k * 4, cudaMemcpyDeviceToHost, sout));
        }
        if (c->d_mc) {
            CU_TRY(h, cudaMemcpyAsync(c->h_mq + out_lo, c->d_mq + out_lo, rows * 4, cudaMemcpyDeviceToHost, sout));
            CU_TRY(h, cudaMemcpyAsync(c->h_mt + out_lo, c->d_mt + out_lo, rows * 4, cudaMemcpyDeviceToHost, sout));
            CU_TRY(h, cudaMemcpyAsync(c->h_md + out_lo, c->d_md + out_lo, rows * 4, cudaMemcpyDeviceToHost, sout));
        }
    }
    if (c->d_mc) CU_TRY(h, cudaMemcpyAsync(c->h_mc + p0, c->d_mc + p0, (size_t)(p1 - p0) * 4, cudaMemcpyDeviceToHost, sout));
    if (c->trace) c->cpu_ms[g][1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c->cpu0).count();
    return BFM_OK;
}

int run_host_pipelined(bfm_handle_t h, const uint8_t *q, int32_t nq_rows, const uint8_t *t, int32_t nt_rows,
                       const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows,
                       const bfm_options_t *o, int32_t *knn_idx, int32_t *knn_dist, int32_t *m_query,
                       int32_t *m_train, int32_t *m_dist, int32_t *m_count) {
    cudaStream_t st = h->stream, sin = h->in_stream, sout = h->out_stream;
    (void)sin;
    const bool window = o->mask_kind == BFM_MASK_WINDOW;
    const size_t qb = (size_t)nq_rows * 32, tb = (size_t)nt_rows * 32;
    const size_t qxy_b = window ? (size_t)nq_rows * 8 : 0, txy_b = window ? (size_t)nt_rows * 8 : 0;
    const size_t o_q = 0, o_t = align256(o_q + qb), o_qxy = align256(o_t + tb), o_txy = align256(o_qxy + qxy_b),
                 in_total = align256(o_txy + txy_b);
    int rc = ensure(h, h->d_in, in_total);
    if (rc) return rc;
    char *din = static_cast<char *>(h->d_in.p);
    bfm_options_t od = *o;
    if (window) {
        od.q_xy = reinterpret_cast<const float *>(din + o_qxy);
        od.t_xy = reinterpret_cast<const float *>(din + o_txy);
    }

    const int k = o->k;
    const size_t knn_b = knn_idx ? (size_t)n_out_rows * k * 4 : 0;
    const size_t m_b = m_count ? (size_t)n_out_rows * 4 : 0;
    const size_t cnt_b = m_count ? (size_t)n_problems * 4 : 0;
    const size_t out_total = 2 * knn_b + 3 * m_b + cnt_b;
    rc = ensure(h, h->d_out, std::max<size_t>(out_total, 16));
    if (rc) return rc;
    char *dout = static_cast<char *>(h->d_out.p);

    // results land directly in the caller's arrays when those are pinned, else in pinned staging
    const bool direct = (!knn_idx || (host_ptr_is_pinned(knn_idx) && host_ptr_is_pinned(knn_dist))) &&
                        (!m_count || (host_ptr_is_pinned(m_query) && host_ptr_is_pinned(m_train) &&
                                      host_ptr_is_pinned(m_dist) && host_ptr_is_pinned(m_count)));
    if (!direct && h->h_out_cap < out_total) {
        if (h->h_out) CU_TRY(h, cudaFreeHost(h->h_out));
        h->h_out = nullptr;
        const size_t want = out_total + out_total / 4 + 4096;
        CU_TRY(h, cudaMallocHost(&h->h_out, want));
        h->h_out_cap = want;
    }
    char *ho = static_cast<char *>(h->h_out);

    PipeCtx ctx{};
    ctx.h = h;
    ctx.problems = problems;
    ctx.n_out_rows = n_out_rows;
    ctx.k = k;
    ctx.d_ki = knn_idx ? reinterpret_cast<int32_t *>(dout) : nullptr;
    ctx.d_kd = knn_idx ? reinterpret_cast<int32_t *>(dout + knn_b) : nullptr;
    ctx.d_mq = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b) : nullptr;
    ctx.d_mt = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b + m_b) : nullptr;
    ctx.d_md = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b + 2 * m_b) : nullptr;
    ctx.d_mc = m_count ? reinterpret_cast<int32_t *>(dout + 2 * knn_b + 3 * m_b) : nullptr;
    ctx.h_ki = direct ? knn_idx : reinterpret_cast<int32_t *>(ho);
    ctx.h_kd = direct ? knn_dist : reinterpret_cast<int32_t *>(ho + knn_b);
    ctx.h_mq = direct ? m_query : reinterpret_cast<int32_t *>(ho + 2 * knn_b);
    ctx.h_mt = direct ? m_train : reinterpret_cast<int32_t *>(ho + 2 * knn_b + m_b);
    ctx.h_md = direct ? m_dist : reinterpret_cast<int32_t *>(ho + 2 * knn_b + 2 * m_b);
    ctx.h_mc = direct ? m_count : reinterpret_cast<int32_t *>(ho + 2 * knn_b + 3 * m_b);
    ctx.trace = std::getenv("BFM_TRACE") != nullptr;
    ctx.cpu0 = std::chrono::steady_clock::now();

    // -- group boundaries: cumulative cost shares 1 : 2 : 4 : ... (first copy short, launches few) ------
    double total_cost = 0;
    for (int p = 0; p < n_problems; ++p) total_cost += (double)problems[p].q_count * problems[p].t_count + 1.0;
    int G = h->pipeline_chunks > 1 ? h->pipeline_chunks : 4;
    G = std::min(std::min(G, MAX_CHUNKS), n_problems);
    int bounds[MAX_CHUNKS + 1];
    bounds[0] = 0;
    {
        const double denom = (double)((1u << G) - 1u);  // 1 + 2 + ... + 2^(G-1)
        double acc = 0;
        int g = 1;
        for (int p = 0; p < n_problems && g < G; ++p) {
            acc += (double)problems[p].q_count * problems[p].t_count + 1.0;
            if (acc >= total_cost * (double)((1u << g) - 1u) / denom) bounds[g++] = p + 1;
        }
        while (g <= G) bounds[g++] = n_problems;
    }
    ctx.bounds = bounds;

    ctx.q = q;
    ctx.t = t;
    ctx.q_xy = window ? o->q_xy : nullptr;
    ctx.t_xy = window ? o->t_xy : nullptr;
    ctx.din = din;
    ctx.o_q = o_q; ctx.o_t = o_t; ctx.o_qxy = o_qxy; ctx.o_txy = o_txy;
    ctx.n_groups = G;
    if (ctx.trace) CU_TRY(h, cudaEventRecord(h->ev[0], sin));

    GroupHooks hooks{G, bounds, &ctx, pipe_before, pipe_after};
    const int64_t launches_before = h->launches;
    rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows, reinterpret_cast<const uint8_t *>(din + o_t),
                    nt_rows, problems, n_problems, n_out_rows, &od, ctx.d_ki, ctx.d_kd, ctx.d_mq, ctx.d_mt, ctx.d_md,
                    ctx.d_mc, st, &hooks);
    if (rc) {
        cudaDeviceSynchronize();
        return rc;
    }
    if (ctx.trace) CU_TRY(h, cudaEventRecord(h->ev[3], sout));
    CU_TRY(h, cudaStreamSynchronize(sout));
    CU_TRY(h, cudaStreamSynchronize(st));
    CU_TRY(h, cudaStreamSynchronize(sin));
    if (ctx.trace) {
        float a = 0, b = 0;
        for (int g = 0; g < G; ++g) {
            if (bounds[g + 1] <= bounds[g]) continue;
            cudaEventElapsedTime(&a, h->ev[0], h->chunk_ev[2 * g]);
            cudaEventElapsedTime(&b, h->ev[0], h->chunk_ev[2 * g + 1]);
            std::fprintf(stderr, "[bfm trace] group %d (%d problems): copy-in done %.3f ms, compute done %.3f ms | cpu: queued at %.3f, out queued %.3f\n",
                         g, bounds[g + 1] - bounds[g], a, b, ctx.cpu_ms[g][0], ctx.cpu_ms[g][1]);
        }
        cudaEventElapsedTime(&a, h->ev[0], h->ev[3]);
        std::fprintf(stderr, "[bfm trace] copy-out done %.3f ms (direct=%d)\n", a, (int)direct);
    }
    if (!direct) {
        if (knn_idx) {
            std::memcpy(knn_idx, ctx.h_ki, knn_b);
            std::memcpy(knn_dist, ctx.h_kd, knn_b);
        }
        if (m_count) {
            std::memcpy(m_count, ctx.h_mc, cnt_b);
            for (int p = 0; p < n_problems; ++p) {
                const size_t b = (size_t)problems[p].out_begin, n = (size_t)m_count[p];
                std::memcpy(m_query + b, ctx.h_mq + b, n * 4);
                std::memcpy(m_train + b, ctx.h_mt + b, n * 4);
                std::memcpy(m_dist + b, ctx.h_md + b, n * 4);
            }
        }
    }
    h->info.kernels_launched = (int32_t)(h->launches - launches_before);
    h->info.reserved = G;
    return BFM_OK;
}
