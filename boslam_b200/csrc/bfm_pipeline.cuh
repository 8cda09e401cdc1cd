// bfm_pipeline.cuh - the BFM_MEM_HOST path.
// (included by bfm_api.cu inside its anonymous namespace, after run_device)
//
// POPC kernel: one kernel launch per call, however large the batch (what overlaps with it is listed below).  Tensor form
// (large unmasked batches in pinned memory): copy-engine chunks of whole problems, each matched by its own three launches
// as it lands, results staged on the device and copied out by a third stream - see the `tchunks` branch of run_host.
//
// One launch, and what overlaps with it:
//
//   inputs   (a) pinned caller arrays - SM-fed upload: the first 24 CTAs of the matching kernel stream the
//            arrays from pinned host memory into HBM themselves (zero-copy loads over PCIe, bfm_kernels.cuh:
//            feed_rows) in rounds of ~8k rows, the first round split in eight, and publish a per-feeder
//            progress word; every other CTA waits at its input gate until all feeders have passed the round
//            that covers its rows.  No copy-engine operation, event or host involvement per slice (each
//            cudaMemcpyAsync costs ~5 us of DMA set-up here, which is what limited the chunked variant).
//            (b) pageable caller arrays - copy-engine chunks: G chunks on `in_stream`, each followed by a
//            16-byte watermark copy; the kernel is launched right after the first chunk has been queued and
//            its CTAs wait on the watermarks.
//            Work items are in problem order and so is the upload, so the SMs chase the data through the
//            batch and the wall time tends to max(upload, compute).
//   outputs  The CTA that completes a problem writes its result rows straight into pinned host memory over
//            PCIe (the caller's arrays when those are pinned, else the handle's pinned staging block): there
//            is no device result buffer and no D2H copy stage.
//
// Small calls (a frame against a keyframe or the local map) use one copy on the compute stream and no gate.
// BFM_TRACE=1 in the environment prints the per-call timeline.

// progress words carry the call's epoch, so they are only zeroed when the 16-bit epoch wraps
uint32_t next_feed_epoch(bfm_handle_t h, cudaStream_t st) {
    if (++h->feed_epoch >= 65536u) {
        cudaMemsetAsync(h->d_prog, 0, 128, st);
        h->feed_epoch = 1;
    }
    return h->feed_epoch;
}

void ensure_pool(bfm_handle_t h) {
    if (h->pool || h->host_threads < 0) return;
    const int n = h->host_threads > 0 ? h->host_threads : (int)std::thread::hardware_concurrency() / 2;
    h->pool.reset(new WorkerPool(std::min(std::max(n, 2), 8) - 1));   // the calling thread works too
}

// Copy chunks of a host batch that takes the tensor form: whole problems, about 8 MB of descriptors each (2 .. 8 chunks;
// `forced` > 0 overrides the count); when all problems have one shape a chunk is rounded to a whole number of rounds of
// the persistent scan - its work items (256-row query blocks) divide by the SM count.  Returns the number of chunks,
// *per = problems per chunk (the last chunk takes what is left).
int plan_host_chunks(const bfm_problem_t *problems, int n_problems, size_t bytes, int sm_count, int forced, int *per_out) {
    int C = forced > 0 ? std::min(forced, n_problems) : (int)std::min<size_t>(8, std::max<size_t>(2, bytes / ((size_t)8 << 20)));
    C = std::max(1, std::min(C, n_problems));
    int per = (n_problems + C - 1) / C;
    bool uniform = true;
    for (int p = 1; p < n_problems && uniform; ++p) uniform = problems[p].q_count == problems[0].q_count && problems[p].t_count == problems[0].t_count;
    if (uniform && problems[0].q_count > 0 && sm_count > 0) {
        const int ipp = (problems[0].q_count + bfm::TC_BQ - 1) / bfm::TC_BQ;   // work items per problem
        int a = sm_count, b = ipp;
        while (b) { const int r = a % b; a = b; b = r; }
        const int unit = sm_count / a;                                         // problems per round of the scan
        if (unit <= 2 * per) per = std::max(1, (per + unit / 2) / unit) * unit;
    }
    *per_out = per;
    return (n_problems + per - 1) / per;
}

bool host_ptr_is_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int run_host(bfm_handle_t h, const uint8_t *q, int32_t nq_rows, const uint8_t *t, int32_t nt_rows,
             const bfm_problem_t *problems, int32_t n_problems, int32_t n_out_rows, const bfm_options_t *o,
             const bfm_outputs_t &user, const bfm_outputs_t *extra = nullptr, int n_extra = 0) {
    // `extra`: up to 7 more destination sets in DEVICE memory (this GPU's or NVLink peers' / a multicast address):
    // the same epilogue that writes the caller's host arrays also delivers the multi-GPU gather (bfm_match_batched_host_multi)
    if (n_extra < 0 || n_extra > bfm::MAX_DEST - 1 || (n_extra > 0 && !extra))
        return fail(h, BFM_ERR_INVALID, "between 0 and 7 extra device destinations are supported");
    if (n_problems <= 0) return BFM_OK;
    if (n_out_rows <= 0) {   // every query set is empty: m_count is still an output per problem
        if (user.m_count) std::memset(user.m_count, 0, (size_t)n_problems * 4);
        return BFM_OK;
    }
    cudaStream_t st = h->stream, sin = h->in_stream;
    // (a device call may still be running on a caller stream: run_device orders `st` behind it before the shared
    // workspace is touched; d_in and the staging blocks are only ever used by these synchronous host calls)
    const bool window = o->mask_kind == BFM_MASK_WINDOW;
    const bool dense = o->mask_kind == BFM_MASK_DENSE && o->mask;
    const size_t qb = (size_t)nq_rows * 32, tb = (size_t)nt_rows * 32;
    const size_t qxy_b = window ? (size_t)nq_rows * 8 : 0, txy_b = window ? (size_t)nt_rows * 8 : 0;
    const size_t mask_b = dense && problems[0].q_count > 0
                              ? (size_t)(problems[0].q_count - 1) * (size_t)o->mask_row_stride + (size_t)problems[0].t_count
                              : 0;
    const size_t o_q = 0, o_t = align256(o_q + qb), o_qxy = align256(o_t + tb), o_txy = align256(o_qxy + qxy_b),
                 o_m = align256(o_txy + txy_b), in_total = align256(o_m + mask_b);
    int rc = ensure(h, h->d_in, in_total);
    if (rc) return rc;
    char *din = static_cast<char *>(h->d_in.p);
    bfm_options_t od = *o;
    if (window) {
        od.q_xy = reinterpret_cast<const float *>(din + o_qxy);
        od.t_xy = reinterpret_cast<const float *>(din + o_txy);
    }
    if (dense) od.mask = reinterpret_cast<const uint8_t *>(din + o_m);

    // -- results: straight into the caller's arrays when those are pinned, else into pinned staging ----
    const int k = o->k;
    const bool want_knn = user.knn_idx != nullptr, want_m = user.m_count != nullptr;
    if ((user.knn_idx == nullptr) != (user.knn_dist == nullptr))
        return fail(h, BFM_ERR_INVALID, "knn_idx and knn_dist must be given together");
    if (want_m && (!user.m_query || !user.m_train || !user.m_dist))
        return fail(h, BFM_ERR_INVALID, "m_query/m_train/m_dist/m_count must be given together");
    const size_t knn_b = want_knn ? align256((size_t)n_out_rows * k * 4) : 0;
    const size_t m_b = want_m ? align256((size_t)n_out_rows * 4) : 0;
    const size_t cnt_b = want_m ? align256((size_t)n_problems * 4) : 0;
    const size_t out_total = 2 * knn_b + 3 * m_b + cnt_b;
    // small results always go through the staging block: asking the driver whether four or six pointers are pinned
    // (~1 us each) costs more than copying a few kilobytes
    const bool direct = out_total >= ((size_t)256 << 10) &&
                        (!want_knn || (host_ptr_is_pinned(user.knn_idx) && host_ptr_is_pinned(user.knn_dist) &&
                                       ((reinterpret_cast<uintptr_t>(user.knn_idx) | reinterpret_cast<uintptr_t>(user.knn_dist)) & 7) == 0)) &&
                        (!want_m || (host_ptr_is_pinned(user.m_query) && host_ptr_is_pinned(user.m_train) &&
                                     host_ptr_is_pinned(user.m_dist) && host_ptr_is_pinned(user.m_count)));
    bfm_outputs_t dests[bfm::MAX_DEST];
    bfm_outputs_t &dst = dests[0];
    dst = user;
    for (int d = 0; d < n_extra; ++d) dests[1 + d] = extra[d];
    const int n_dests = 1 + n_extra;
    if (!direct) {
        if (h->h_out_cap < out_total) {
            if (h->h_out) CU_TRY(h, cudaFreeHost(h->h_out));
            h->h_out = nullptr;
            h->h_out_cap = 0;
            const size_t want = out_total + out_total / 4 + 4096;
            CU_TRY(h, cudaMallocHost(&h->h_out, want));
            h->h_out_cap = want;
        }
        char *ho = static_cast<char *>(h->h_out);
        dst.knn_idx = want_knn ? reinterpret_cast<int32_t *>(ho) : nullptr;
        dst.knn_dist = want_knn ? reinterpret_cast<int32_t *>(ho + knn_b) : nullptr;
        dst.m_query = want_m ? reinterpret_cast<int32_t *>(ho + 2 * knn_b) : nullptr;
        dst.m_train = want_m ? reinterpret_cast<int32_t *>(ho + 2 * knn_b + m_b) : nullptr;
        dst.m_dist = want_m ? reinterpret_cast<int32_t *>(ho + 2 * knn_b + 2 * m_b) : nullptr;
        dst.m_count = want_m ? reinterpret_cast<int32_t *>(ho + 2 * knn_b + 3 * m_b) : nullptr;
    }

    // -- input chunks -----------------------------------------------------------------------------------
    // chunk boundaries are problem boundaries; shares grow geometrically (short first copy: the kernel
    // starts early; few large copies later: the copy engine runs at full rate)
    int G = 1;
    if (h->pipeline_chunks > 1) {
        G = std::min(h->pipeline_chunks, n_problems);
    } else if (h->pipeline_chunks == 0 && n_problems >= 4 && !dense && qb + tb >= ((size_t)(h->pipeline_min_kb > 0 ? h->pipeline_min_kb : 1024) << 10)) {
        G = std::min(12, n_problems);
    }
    // (one large problem - a frame against the local map, 0.7 MB - keeps the plain copy: the gated form measured
    // slower for it, 197 us with query-block-major work items and 212 us with train-range-major items and 2048-row feed
    // rounds against 137 us: seventeen latency-bound feed rounds cost more than the 25 us of DMA they would hide)
    const bool trace = std::getenv("BFM_TRACE") != nullptr;
    const auto cpu0 = std::chrono::steady_clock::now();
    auto cpu_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - cpu0).count(); };

    // SM-fed upload: pinned caller arrays are streamed into HBM by the first CTAs of the matching kernel itself
    // (bfm_kernels.cuh: feed_rows) - no copy-engine operation per slice, so slices are as fine as a keyframe pair
    auto feedable = [](const void *p) { return host_ptr_is_pinned(p) && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool feed = G > 1 && h->feeders >= 0 && feedable(q) && feedable(t) &&
                      (!window || (feedable(o->q_xy) && feedable(o->t_xy))) && o->k <= 2;
    // Tensor form for a batch from pinned host memory: the copy engine brings the descriptors in a few chunks on
    // `in_stream` (55 GB/s here, against ~30 GB/s of zero-copy reads by feeder CTAs) and every chunk is matched by the
    // tensor kernels (expansion, tcgen05 scan, finalize) as soon as it has landed.  The batch is then bound by the
    // upload alone: the last chunk's 50 us of matching is all that is not hidden.  Chunks are whole problems; with
    // equal shapes a chunk is a whole number of rounds of the persistent scan (its items divide by the SM count).
    long long total_pairs_h = 0;
    for (int p = 0; p < n_problems; ++p) total_pairs_h += (long long)std::max(0, problems[p].q_count) * std::max(0, problems[p].t_count);
    const bool tchunks = h->tensor != 1 && h->feeders >= 0 && h->pipeline_chunks == 0 && o->mask_kind == BFM_MASK_NONE && !o->cross_check && o->k <= 2 &&
                         n_problems >= 8 && qb + tb >= ((size_t)4 << 20) && total_pairs_h >= 16 * TENSOR_MIN_PAIRS && feedable(q) && feedable(t);
    if (tchunks) {
        int per = 0;
        const int C = plan_host_chunks(problems, n_problems, qb + tb, h->sm_count, h->tensor_chunks, &per);
        for (int c = 0; c < C; ++c) {
            if (!h->chunk_ev[c]) CU_TRY(h, cudaEventCreateWithFlags(&h->chunk_ev[c], cudaEventDisableTiming));
            if (!h->chunk_done_ev[c]) CU_TRY(h, cudaEventCreateWithFlags(&h->chunk_done_ev[c], cudaEventDisableTiming));
        }
        if (!h->out_stream) CU_TRY(h, cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
        *h->h_status = 0;
        // Results: the finalize kernel of a chunk writes into a device block and a third stream copies the chunk's rows to
        // the host (pinned caller arrays or the staging block) - stores over PCIe by the kernel itself, which the one-launch
        // paths hide under a millisecond of matching, would cost a chunk as much again as its scan (measured: 206 us per
        // 74 pairs instead of ~100), and the copy engine of the other direction is idle anyway.
        rc = ensure(h, h->d_res, out_total);
        if (rc) return rc;
        char *dr = static_cast<char *>(h->d_res.p);
        bfm_outputs_t ddst = dst;
        ddst.knn_idx = want_knn ? reinterpret_cast<int32_t *>(dr) : nullptr;
        ddst.knn_dist = want_knn ? reinterpret_cast<int32_t *>(dr + knn_b) : nullptr;
        ddst.m_query = want_m ? reinterpret_cast<int32_t *>(dr + 2 * knn_b) : nullptr;
        ddst.m_train = want_m ? reinterpret_cast<int32_t *>(dr + 2 * knn_b + m_b) : nullptr;
        ddst.m_dist = want_m ? reinterpret_cast<int32_t *>(dr + 2 * knn_b + 2 * m_b) : nullptr;
        ddst.m_count = want_m ? reinterpret_cast<int32_t *>(dr + 2 * knn_b + 3 * m_b) : nullptr;
        const cudaStream_t sout = h->out_stream;
        // The copy engine serves `in_stream` for as long as that stream has copies ready: a small table upload of a
        // chunk's launch on `st` waited until the whole descriptor upload was over (measured: the first kernel started
        // 0.4 ms late).  So the launches fetch their tables with a few threads instead (tables_by_kernel), and the copy of
        // chunk c + 1 is queued right after the launch of chunk c (~25 us of host time against ~85 us on the wire).
        int32_t q_mark = 0, t_mark = 0;
        cudaEvent_t tev0 = nullptr, tev_copy[8] = {nullptr}, tev_comp[8] = {nullptr};   // BFM_TRACE: the device's own timeline
        auto copy_chunk = [&](int c) -> int {
            const int p0 = c * per, p1 = std::min(n_problems, p0 + per);
            int32_t q_need = q_mark, t_need = t_mark;
            for (int p = p0; p < p1; ++p) {
                if (problems[p].q_count <= 0 || problems[p].t_count <= 0) continue;
                q_need = std::max(q_need, problems[p].q_begin + problems[p].q_count);
                t_need = std::max(t_need, problems[p].t_begin + problems[p].t_count);
            }
            if (q_need > q_mark) CU_TRY(h, cudaMemcpyAsync(din + o_q + (size_t)q_mark * 32, q + (size_t)q_mark * 32, (size_t)(q_need - q_mark) * 32, cudaMemcpyHostToDevice, sin));
            if (t_need > t_mark) CU_TRY(h, cudaMemcpyAsync(din + o_t + (size_t)t_mark * 32, t + (size_t)t_mark * 32, (size_t)(t_need - t_mark) * 32, cudaMemcpyHostToDevice, sin));
            q_mark = q_need; t_mark = t_need;
            CU_TRY(h, cudaEventRecord(h->chunk_ev[c], sin));
            if (trace) cudaEventRecord(tev_copy[c], sin);
            return BFM_OK;
        };
        if (trace) {
            cudaEventCreate(&tev0);
            for (int c = 0; c < C; ++c) { cudaEventCreate(&tev_copy[c]); cudaEventCreate(&tev_comp[c]); }
            cudaEventRecord(tev0, sin);
        }
        rc = copy_chunk(0);
        const double t_copies = cpu_ms();
        // every chunk takes the tensor form, whatever its size, and fetches its tables by kernel; restored on every way out
        struct ChunkMode {
            bfm_handle_t h;
            int saved;
            explicit ChunkMode(bfm_handle_t h_) : h(h_), saved(h_->tensor) { h->tensor = 2; h->tables_by_kernel = true; }
            ~ChunkMode() { h->tensor = saved; h->tables_by_kernel = false; }
        } chunk_mode(h);
        int launched = 0;
        double t_chunk[16] = {0};
        for (int c = 0; c < C && rc == BFM_OK; ++c) {
            const int p0 = c * per, p1 = std::min(n_problems, p0 + per);
            bfm_outputs_t cd[bfm::MAX_DEST];
            for (int d = 0; d < n_dests; ++d) {
                cd[d] = d == 0 ? ddst : dests[d];
                if (cd[d].m_count) cd[d].m_count += p0;   // per problem of the call; everything else is indexed by output row
            }
            if (cudaStreamWaitEvent(st, h->chunk_ev[c], 0) != cudaSuccess) { rc = fail(h, BFM_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
            rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows, reinterpret_cast<const uint8_t *>(din + o_t), nt_rows,
                            problems + p0, p1 - p0, n_out_rows, &od, cd, n_dests, st);
            launched += h->info.kernels_launched;
            if (rc == BFM_OK) {
                // this chunk's output rows (any order of out_begin is fine: a row copied again later carries the same bits)
                long long r0 = n_out_rows, r1 = 0;
                for (int p = p0; p < p1; ++p) {
                    if (problems[p].q_count <= 0) continue;
                    r0 = std::min<long long>(r0, problems[p].out_begin);
                    r1 = std::max<long long>(r1, (long long)problems[p].out_begin + problems[p].q_count);
                }
                CU_TRY(h, cudaEventRecord(h->chunk_done_ev[c], st));
                CU_TRY(h, cudaStreamWaitEvent(sout, h->chunk_done_ev[c], 0));
                if (r1 > r0) {
                    const size_t b = (size_t)r0, n = (size_t)(r1 - r0);
                    if (want_knn) {
                        CU_TRY(h, cudaMemcpyAsync(dst.knn_idx + b * k, ddst.knn_idx + b * k, n * k * 4, cudaMemcpyDeviceToHost, sout));
                        CU_TRY(h, cudaMemcpyAsync(dst.knn_dist + b * k, ddst.knn_dist + b * k, n * k * 4, cudaMemcpyDeviceToHost, sout));
                    }
                    if (want_m) {
                        CU_TRY(h, cudaMemcpyAsync(dst.m_query + b, ddst.m_query + b, n * 4, cudaMemcpyDeviceToHost, sout));
                        CU_TRY(h, cudaMemcpyAsync(dst.m_train + b, ddst.m_train + b, n * 4, cudaMemcpyDeviceToHost, sout));
                        CU_TRY(h, cudaMemcpyAsync(dst.m_dist + b, ddst.m_dist + b, n * 4, cudaMemcpyDeviceToHost, sout));
                    }
                }
                if (want_m) CU_TRY(h, cudaMemcpyAsync(dst.m_count + p0, ddst.m_count + p0, (size_t)(p1 - p0) * 4, cudaMemcpyDeviceToHost, sout));
            }
            if (trace) cudaEventRecord(tev_comp[c], st);
            t_chunk[2 * c] = cpu_ms();
            if (rc == BFM_OK && c + 1 < C) rc = copy_chunk(c + 1);
            t_chunk[2 * c + 1] = cpu_ms();
        }
        const double t_launched = cpu_ms();
        if (rc) {
            cudaDeviceSynchronize();
            h->state_clean = false;
            return rc;
        }
        CU_TRY(h, cudaStreamSynchronize(sout));   // (behind the last chunk's kernels on st)
        if (*h->h_status != 0) {
            h->state_clean = false;
            return fail(h, BFM_ERR_CUDA, "tensor scan timed out on a barrier");
        }
        h->info.copy_chunks = C;
        h->info.kernels_launched = launched;
        if (trace) std::fprintf(stderr, "[bfm trace] tensor form, %d copy-engine chunks of %d problems: copies queued %.3f, kernels queued %.3f, done %.3f ms (direct=%d)\n",
                                C, per, t_copies, t_launched, cpu_ms(), (int)direct);
        if (trace) {
            std::fprintf(stderr, "[bfm trace]   per chunk (launch queued / next copy queued, ms):");
            for (int c = 0; c < C; ++c) std::fprintf(stderr, " %.3f/%.3f", t_chunk[2 * c], t_chunk[2 * c + 1]);
            std::fprintf(stderr, "\n[bfm trace]   on the device (copy landed / chunk matched, ms after the first copy was queued):");
            for (int c = 0; c < C; ++c) {
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, tev0, tev_copy[c]);
                cudaEventElapsedTime(&b, tev0, tev_comp[c]);
                std::fprintf(stderr, " %.3f/%.3f", a, b);
                cudaEventDestroy(tev_copy[c]);
                cudaEventDestroy(tev_comp[c]);
            }
            cudaEventDestroy(tev0);
            std::fprintf(stderr, "\n");
        }
    } else if (feed) {
        Gate gate;
        gate.status = h->h_status;
        *h->h_status = 0;
        gate.n_feed = h->feeders > 0 ? h->feeders : 24;
        const int rows_per_round = std::max(h->feed_rows > 0 ? h->feed_rows : 8192, std::max(nq_rows, nt_rows) / 60000 + 1);  // rounds fit 16 bits
        const int S = std::max(1, (std::max(nq_rows, nt_rows) + rows_per_round - 1) / rows_per_round);
        gate.rounds = S + bfm::FEED_HEAD - 1;   // the first round is delivered as FEED_HEAD short ones
        // rows per round are multiples of 128, so even the eighth-rounds of the head end on 128-byte boundaries of
        // every array (16 rows x 8 B of pixel coordinates): no 32-byte sector is shared by two rounds
        gate.q_rows = std::max(128, (((nq_rows + S - 1) / S) + 127) & ~127);
        gate.t_rows = std::max(128, (((nt_rows + S - 1) / S) + 127) & ~127);
        gate.src[0] = q; gate.dst[0] = din + o_q; gate.bytes[0] = qb;
        gate.src[1] = t; gate.dst[1] = din + o_t; gate.bytes[1] = tb;
        if (window) {
            gate.src[2] = o->q_xy; gate.dst[2] = din + o_qxy; gate.bytes[2] = qxy_b;
            gate.src[3] = o->t_xy; gate.dst[3] = din + o_txy; gate.bytes[3] = txy_b;
        }
        gate.prog = h->d_prog;
        gate.epoch = next_feed_epoch(h, st);
        const double t_prep = cpu_ms();
        rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows, reinterpret_cast<const uint8_t *>(din + o_t),
                        nt_rows, problems, n_problems, n_out_rows, &od, dests, n_dests, st, &gate);
        const double t_launched = cpu_ms();
        if (rc) {
            cudaDeviceSynchronize();
            h->state_clean = false;
            return rc;
        }
        CU_TRY(h, cudaStreamSynchronize(st));
        if (*h->h_status != 0) {
            h->state_clean = false;
            return fail(h, BFM_ERR_CUDA, "input gate timed out: the feeder CTAs did not deliver the descriptors");
        }
        h->info.copy_chunks = gate.rounds;
        if (trace) std::fprintf(stderr, "[bfm trace] SM-fed upload: %d feeders, %d rounds: prepared %.3f, kernel queued %.3f, done %.3f ms (direct=%d)\n",
                                gate.n_feed, gate.rounds, t_prep, t_launched, cpu_ms(), (int)direct);
    } else if (G > 1 && h->feeders >= 0 && h->host_threads >= 0 && o->k <= 2) {
        // Pageable caller arrays: worker threads stage them into pinned memory round by round while the kernel
        // is already running; its feeder CTAs wait for the host's "rounds staged" word before each round.
        ensure_pool(h);
        if (h->h_stage_cap < in_total) {
            if (h->h_stage) CU_TRY(h, cudaFreeHost(h->h_stage));
            h->h_stage = nullptr;
            h->h_stage_cap = 0;
            CU_TRY(h, cudaMallocHost(&h->h_stage, in_total + in_total / 4 + 4096));
            h->h_stage_cap = in_total + in_total / 4 + 4096;
        }
        char *stage = static_cast<char *>(h->h_stage);
        Gate gate;
        gate.status = h->h_status;
        *h->h_status = 0;
        gate.n_feed = h->feeders > 0 ? h->feeders : 24;
        const int rows_per_round = std::max(h->feed_rows > 0 ? h->feed_rows : 8192, std::max(nq_rows, nt_rows) / 60000 + 1);  // rounds fit 16 bits
        const int S = std::max(1, (std::max(nq_rows, nt_rows) + rows_per_round - 1) / rows_per_round);
        gate.rounds = S + bfm::FEED_HEAD - 1;
        // rows per round are multiples of 128, so even the eighth-rounds of the head end on 128-byte boundaries of
        // every array (16 rows x 8 B of pixel coordinates): no 32-byte sector is shared by two rounds
        gate.q_rows = std::max(128, (((nq_rows + S - 1) / S) + 127) & ~127);
        gate.t_rows = std::max(128, (((nt_rows + S - 1) / S) + 127) & ~127);
        const void *user[4] = {q, t, window ? (const void *)o->q_xy : nullptr, window ? (const void *)o->t_xy : nullptr};
        const size_t offs[4] = {o_q, o_t, o_qxy, o_txy}, bytes[4] = {qb, tb, qxy_b, txy_b};
        for (int a = 0; a < 4; ++a) {
            if (!user[a] || !bytes[a]) continue;
            gate.src[a] = stage + offs[a];
            gate.dst[a] = din + offs[a];
            gate.bytes[a] = bytes[a];
        }
        gate.prog = h->d_prog;
        gate.host_ready = h->h_ready;
        *h->h_ready = 0;
        // jobs: one memcpy per (round, array); the byte ranges are exactly the feeders' (feed_rows)
        const int rounds = gate.rounds;
        struct Shared {
            std::vector<std::atomic<int>> left;
            std::mutex adv;
            int next = 0;
            explicit Shared(int n) : left(n) {}
        };
        auto sh = std::make_shared<Shared>(rounds);
        volatile uint32_t *ready = h->h_ready;
        auto advance = [sh, ready, rounds]() {
            std::lock_guard<std::mutex> lk(sh->adv);
            while (sh->next < rounds && sh->left[sh->next].load(std::memory_order_acquire) == 0) ++sh->next;
            std::atomic_thread_fence(std::memory_order_release);
            *ready = (uint32_t)sh->next;
        };
        struct Job { int r; char *dst; const char *src; size_t n; };
        std::vector<Job> jobs;
        for (int r = 0; r < rounds; ++r) {
            int cnt = 0;
            for (int a = 0; a < 4; ++a) {
                if (!user[a] || !bytes[a]) continue;
                const size_t rows = (a & 1) ? (size_t)gate.t_rows : (size_t)gate.q_rows;
                const size_t per_round = rows * (a < 2 ? 32 : 8);
                const size_t b0 = r < bfm::FEED_HEAD ? (size_t)r * (per_round / bfm::FEED_HEAD) : (size_t)(r - bfm::FEED_HEAD + 1) * per_round;
                const size_t b1 = r < bfm::FEED_HEAD ? b0 + per_round / bfm::FEED_HEAD : b0 + per_round;
                const size_t lo = std::min(b0, bytes[a]), hi = std::min(b1, bytes[a]);
                if (hi > lo) {
                    jobs.push_back({r, stage + offs[a] + lo, static_cast<const char *>(user[a]) + lo, hi - lo});
                    ++cnt;
                }
            }
            sh->left[r].store(cnt, std::memory_order_relaxed);
        }
        std::vector<std::function<void()>> work;
        work.reserve(jobs.size());
        for (const Job &j : jobs)
            work.emplace_back([j, sh, advance]() {
                stream_copy(j.dst, j.src, j.n);   // non-temporal: the next reader is the GPU, over PCIe
                if (sh->left[j.r].fetch_sub(1, std::memory_order_acq_rel) == 1) advance();
            });
        h->pool->submit_batch(std::move(work));
        advance();   // leading rounds without bytes
        const double t_submitted = cpu_ms();
        gate.epoch = next_feed_epoch(h, st);
        rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows, reinterpret_cast<const uint8_t *>(din + o_t),
                        nt_rows, problems, n_problems, n_out_rows, &od, dests, n_dests, st, &gate);
        const double t_launched = cpu_ms();
        h->pool->help_and_wait();   // also after a failed launch: the jobs reference this call's buffers
        advance();
        const double t_staged = cpu_ms();
        if (rc) {
            cudaDeviceSynchronize();
            h->state_clean = false;
            return rc;
        }
        CU_TRY(h, cudaStreamSynchronize(st));
        if (*h->h_status != 0) {
            h->state_clean = false;
            return fail(h, BFM_ERR_CUDA, "input gate timed out: the staged upload did not reach the GPU");
        }
        h->info.copy_chunks = gate.rounds;
        if (trace) std::fprintf(stderr, "[bfm trace] host-staged SM-fed upload: %d threads, %d feeders, %d rounds: jobs submitted %.3f, kernel "
                                "launched %.3f, staged %.3f, kernel done %.3f ms (direct=%d)\n",
                                h->pool->size() + 1, gate.n_feed, gate.rounds, t_submitted, t_launched, t_staged, cpu_ms(), (int)direct);
    } else if (G <= 1) {
        if (in_total <= ((size_t)160 << 10) && !mask_b) {
            // a small frame (two ORB frames: 64 KB): the arrays are gathered into the pinned staging block with the
            // device layout and go up in ONE copy - each cudaMemcpyAsync from pageable memory costs ~6 us of host time
            if (h->h_stage_cap < in_total) {
                if (h->h_stage) CU_TRY(h, cudaFreeHost(h->h_stage));
                h->h_stage = nullptr;
                h->h_stage_cap = 0;
                const size_t want = std::max(in_total + in_total / 4 + 4096, (size_t)256 << 10);
                CU_TRY(h, cudaMallocHost(&h->h_stage, want));
                h->h_stage_cap = want;
            }
            char *stage = static_cast<char *>(h->h_stage);
            if (qb) std::memcpy(stage + o_q, q, qb);
            if (tb) std::memcpy(stage + o_t, t, tb);
            if (qxy_b) std::memcpy(stage + o_qxy, o->q_xy, qxy_b);
            if (txy_b) std::memcpy(stage + o_txy, o->t_xy, txy_b);
            CU_TRY(h, cudaMemcpyAsync(din, stage, in_total, cudaMemcpyHostToDevice, st));
        } else {
            if (qb) CU_TRY(h, cudaMemcpyAsync(din + o_q, q, qb, cudaMemcpyHostToDevice, st));
            if (tb) CU_TRY(h, cudaMemcpyAsync(din + o_t, t, tb, cudaMemcpyHostToDevice, st));
            if (mask_b) CU_TRY(h, cudaMemcpyAsync(din + o_m, o->mask, mask_b, cudaMemcpyHostToDevice, st));
            if (qxy_b) {
                CU_TRY(h, cudaMemcpyAsync(din + o_qxy, o->q_xy, qxy_b, cudaMemcpyHostToDevice, st));
                CU_TRY(h, cudaMemcpyAsync(din + o_txy, o->t_xy, txy_b, cudaMemcpyHostToDevice, st));
            }
        }
        const double t_copies = cpu_ms();
        rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows, reinterpret_cast<const uint8_t *>(din + o_t),
                        nt_rows, problems, n_problems, n_out_rows, &od, dests, n_dests, st);
        if (rc) return rc;
        const double t_launch = cpu_ms();
        CU_TRY(h, cudaStreamSynchronize(st));
        h->info.copy_chunks = 1;
        if (trace)
            std::fprintf(stderr, "[bfm trace] 1 chunk: copies queued %.3f ms, kernel queued %.3f ms, synced %.3f ms (direct=%d)\n",
                         t_copies, t_launch, cpu_ms(), (int)direct);
    } else {
        int bounds[MAX_COPY_CHUNKS + 1];
        {
            double total = 0;
            for (int p = 0; p < n_problems; ++p) total += (double)problems[p].q_count * problems[p].t_count + 1.0;
            // shares 1, 1.5, 1.5^2, ... capped at 4x the first, normalised
            double share[MAX_COPY_CHUNKS], sum = 0;
            for (int g = 0; g < G; ++g) {
                share[g] = std::min(std::pow(1.5, g), 4.0);
                sum += share[g];
            }
            bounds[0] = 0;
            double acc = 0, target = 0;
            int g = 1;
            target = total * share[0] / sum;
            for (int p = 0; p < n_problems && g < G; ++p) {
                acc += (double)problems[p].q_count * problems[p].t_count + 1.0;
                if (acc >= target - 1e-9) {
                    bounds[g] = p + 1;
                    target += total * share[g] / sum;
                    ++g;
                }
            }
            while (g <= G) bounds[g++] = n_problems;
        }
        const unsigned long long base = (++h->seq) << 32;
        *h->h_status = 0;
        int32_t q_mark = 0, t_mark = 0;
        auto copy_chunk = [&](int g) -> int {
            int32_t q_need = q_mark, t_need = t_mark;
            for (int p = bounds[g]; p < bounds[g + 1]; ++p) {
                if (problems[p].q_count <= 0 || problems[p].t_count <= 0) continue;
                q_need = std::max(q_need, problems[p].q_begin + problems[p].q_count);
                t_need = std::max(t_need, problems[p].t_begin + problems[p].t_count);
            }
            if (g == G - 1) { q_need = nq_rows; t_need = nt_rows; }
            if (q_need > q_mark) {
                const size_t b = (size_t)q_mark, n = (size_t)(q_need - q_mark);
                CU_TRY(h, cudaMemcpyAsync(din + o_q + b * 32, q + b * 32, n * 32, cudaMemcpyHostToDevice, sin));
                if (window) CU_TRY(h, cudaMemcpyAsync(din + o_qxy + b * 8, o->q_xy + b * 2, n * 8, cudaMemcpyHostToDevice, sin));
                q_mark = q_need;
            }
            if (t_need > t_mark) {
                const size_t b = (size_t)t_mark, n = (size_t)(t_need - t_mark);
                CU_TRY(h, cudaMemcpyAsync(din + o_t + b * 32, t + b * 32, n * 32, cudaMemcpyHostToDevice, sin));
                if (window) CU_TRY(h, cudaMemcpyAsync(din + o_txy + b * 8, o->t_xy + b * 2, n * 8, cudaMemcpyHostToDevice, sin));
                t_mark = t_need;
            }
            h->h_marks[2 * g] = base + (unsigned long long)q_mark;
            h->h_marks[2 * g + 1] = base + (unsigned long long)t_mark;
            CU_TRY(h, cudaMemcpyAsync(h->d_ready, h->h_marks + 2 * g, 16, cudaMemcpyHostToDevice, sin));
            return BFM_OK;
        };
        // the watermark epoch (`base`) makes a reset of d_ready unnecessary: values of earlier calls are smaller
        cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
        if (trace) {
            for (auto &e : tev) cudaEventCreate(&e);
            cudaEventRecord(tev[0], sin);
            cudaEventRecord(tev[2], st);
        }
        rc = copy_chunk(0);
        if (rc) return rc;
        Gate gate;
        gate.ready = h->d_ready;
        gate.base = base;
        gate.status = h->h_status;
        const double t_launch0 = cpu_ms();
        rc = run_device(h, reinterpret_cast<const uint8_t *>(din + o_q), nq_rows, reinterpret_cast<const uint8_t *>(din + o_t),
                        nt_rows, problems, n_problems, n_out_rows, &od, dests, n_dests, st, &gate);
        const double t_launch1 = cpu_ms();
        // the remaining chunks MUST be queued even if the launch failed half-way: CTAs may be waiting
        int rc2 = BFM_OK;
        for (int g = 1; g < G && rc2 == BFM_OK; ++g) rc2 = copy_chunk(g);
        const double t_copies = cpu_ms();
        if (trace) {
            cudaEventRecord(tev[1], sin);
            cudaEventRecord(tev[3], st);
        }
        if (rc || rc2) {
            cudaDeviceSynchronize();
            h->state_clean = false;
            return rc ? rc : rc2;
        }
        CU_TRY(h, cudaStreamSynchronize(st));
        CU_TRY(h, cudaStreamSynchronize(sin));
        if (*h->h_status != 0) {
            h->state_clean = false;
            return fail(h, BFM_ERR_CUDA, "input gate timed out: the descriptor upload never completed");
        }
        h->info.copy_chunks = G;
        if (trace) {
            float copy_ms = 0, kern_ms = 0, skew_ms = 0;
            cudaEventElapsedTime(&copy_ms, tev[0], tev[1]);
            cudaEventElapsedTime(&kern_ms, tev[2], tev[3]);
            cudaEventElapsedTime(&skew_ms, tev[0], tev[2]);
            for (auto &e : tev) cudaEventDestroy(e);
            std::fprintf(stderr, "[bfm trace] %d chunks: launch queued %.3f-%.3f ms, copies queued %.3f ms, done %.3f ms (direct=%d) | "
                         "device: copies %.3f ms, kernel %.3f ms (starts %.3f ms after the first copy)\n",
                         G, t_launch0, t_launch1, t_copies, cpu_ms(), (int)direct, copy_ms, kern_ms, skew_ms);
        }
    }

    if (!direct) {
        // pinned staging -> the caller's pageable arrays; large results are copied by the worker threads
        if (out_total >= ((size_t)1 << 20)) ensure_pool(h);
        const bool par = h->pool && out_total >= ((size_t)1 << 20);
        auto run = [&](std::function<void()> f) { if (par) h->pool->submit(std::move(f)); else f(); };
        if (want_knn) {
            const size_t total = (size_t)n_out_rows * k * 4, parts = par ? 8 : 1, step = (total / parts + 63) & ~(size_t)63;
            for (size_t b = 0; b < total; b += std::max<size_t>(step, 64)) {
                const size_t n = std::min(std::max<size_t>(step, 64), total - b);
                char *ui = reinterpret_cast<char *>(user.knn_idx), *ud = reinterpret_cast<char *>(user.knn_dist);
                const char *si = reinterpret_cast<const char *>(dst.knn_idx), *sd = reinterpret_cast<const char *>(dst.knn_dist);
                run([=]() { std::memcpy(ui + b, si + b, n); std::memcpy(ud + b, sd + b, n); });
            }
        }
        if (want_m) {
            std::memcpy(user.m_count, dst.m_count, (size_t)n_problems * 4);
            // only the filled prefix of every problem's slice is meaningful; copy exactly that
            const int group = par ? std::max(1, n_problems / 16) : n_problems;
            for (int p0 = 0; p0 < n_problems; p0 += group) {
                const int p1 = std::min(n_problems, p0 + group);
                const bfm_outputs_t u = user, d2 = dst;
                run([=]() {
                    for (int p = p0; p < p1; ++p) {
                        const size_t b = (size_t)problems[p].out_begin, n = (size_t)u.m_count[p];
                        std::memcpy(u.m_query + b, d2.m_query + b, n * 4);
                        std::memcpy(u.m_train + b, d2.m_train + b, n * 4);
                        std::memcpy(u.m_dist + b, d2.m_dist + b, n * 4);
                    }
                });
            }
        }
        if (par) h->pool->help_and_wait();
        if (trace) std::fprintf(stderr, "[bfm trace] results copied to the caller's arrays at %.3f ms\n", cpu_ms());
    }
    return BFM_OK;
}
