/*
 * Plain-C CPU restatement of the ORB Hamming-matching hot path.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/ (as a second oracle at sizes numpy is slow at) and by bench.py's
 * cpu_baseline leg when cv2 is absent (kind "port").  Never linked into the product library.
 *
 * Restates what boslam gets from cv2.BFMatcher(NORM_HAMMING) at slam/tracking.py:56,121
 * (OpenCV: BFMatcher::knnMatchImpl -> cv::batchDistance -> hal::normHamming; source not under
 * /root/reference) as rules R1-R6 of SURVEY.md section 8(c):
 *   R1 distance = popcount(q ^ t) over 32 bytes, R2 ascending (distance, trainIdx),
 *   R3 -1 padding when fewer than k candidates, R4 mask non-zero = allowed,
 *   R5 cross-check = mutual argmin with lowest-index ties, R6 masked cross-check = +inf then R5.
 * Pinned against live cv2 and the cv2-generated fixtures by tests/test_oracle.py.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC) -> oracle/_build/liboracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_DESC_BYTES 32
#define ORC_INF 0x7fffffff

static inline int orc_hamming32(const uint8_t *a, const uint8_t *b) {
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

/* R1: full distance matrix, int32[nq*nt] row-major. */
void orc_hamming_matrix(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *D) {
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < nt; ++j)
            D[(size_t)i * nt + j] = orc_hamming32(q + (size_t)i * 32, t + (size_t)j * 32);
}

/* R2-R4: k nearest per query by insertion into an ascending (distance, index) list.
 * mask may be NULL; mask_stride is the row pitch in bytes.  out_* are [nq*k], -1 padded. */
void orc_knn(const uint8_t *q, int nq, const uint8_t *t, int nt, int k, const uint8_t *mask,
             int64_t mask_stride, int32_t *out_idx, int32_t *out_dist) {
    for (int i = 0; i < nq; ++i) {
        int32_t *bi = out_idx + (size_t)i * k, *bd = out_dist + (size_t)i * k;
        for (int c = 0; c < k; ++c) { bi[c] = -1; bd[c] = ORC_INF; }
        for (int j = 0; j < nt; ++j) {
            if (mask && !mask[(size_t)i * mask_stride + j]) continue;
            int d = orc_hamming32(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d >= bd[k - 1]) continue;           /* j ascends: strict '<' keeps the lowest index */
            int c = k - 1;
            while (c > 0 && d < bd[c - 1]) { bd[c] = bd[c - 1]; bi[c] = bi[c - 1]; --c; }
            bd[c] = d; bi[c] = j;
        }
        for (int c = 0; c < k; ++c) if (bi[c] < 0) bd[c] = -1;
    }
}

/* R5/R6: mutual nearest neighbours.  Writes up to nq matches, returns the count. */
int orc_cross_check(const uint8_t *q, int nq, const uint8_t *t, int nt, const uint8_t *mask,
                    int64_t mask_stride, int32_t *mq, int32_t *mt, int32_t *md) {
    if (nq <= 0 || nt <= 0) return 0;
    int32_t *rb = (int32_t *)malloc(sizeof(int32_t) * nq), *rd = (int32_t *)malloc(sizeof(int32_t) * nq);
    int32_t *cb = (int32_t *)malloc(sizeof(int32_t) * nt), *cd = (int32_t *)malloc(sizeof(int32_t) * nt);
    for (int j = 0; j < nt; ++j) { cb[j] = -1; cd[j] = ORC_INF; }
    for (int i = 0; i < nq; ++i) {
        rb[i] = -1; rd[i] = ORC_INF;
        for (int j = 0; j < nt; ++j) {
            if (mask && !mask[(size_t)i * mask_stride + j]) continue;
            int d = orc_hamming32(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < rd[i]) { rd[i] = d; rb[i] = j; }   /* lowest j on ties */
            if (d < cd[j]) { cd[j] = d; cb[j] = i; }   /* i ascends: lowest i on ties */
        }
    }
    int n = 0;
    for (int i = 0; i < nq; ++i)
        if (rb[i] >= 0 && cb[rb[i]] == i) { mq[n] = i; mt[n] = rb[i]; md[n] = rd[i]; ++n; }
    free(rb); free(rd); free(cb); free(cd);
    return n;
}
