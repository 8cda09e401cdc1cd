/* A plain-C caller of the C ABI (include/bfm.h): no Python, no torch.  Matches two small descriptor sets with
 * cross-check + gate through bfm_match (the call slam/tracking.py:56-57 binds to) and a k = 2 table through
 * bfm_knn, and checks both against a naive loop written here.  Built and run by tests/test_c_abi_gpu.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bfm.h"

static uint32_t rng_state = 12345u;
static uint32_t rnd(void) { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static int hamming(const uint8_t *a, const uint8_t *b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

int main(void) {
    enum { NQ = 300, NT = 500 };
    static uint8_t q[NQ][32] __attribute__((aligned(16))), t[NT][32] __attribute__((aligned(16)));
    for (int j = 0; j < NT; ++j) for (int b = 0; b < 32; ++b) t[j][b] = (uint8_t)rnd();
    for (int i = 0; i < NQ; ++i) {
        if (i % 3) { memcpy(q[i], t[rnd() % NT], 32); q[i][rnd() % 32] ^= (uint8_t)(1u << (rnd() % 8)); }
        else for (int b = 0; b < 32; ++b) q[i][b] = (uint8_t)rnd();
    }
    bfm_handle_t h = NULL;
    if (bfm_create(0, &h) != BFM_OK) { printf("bfm_create: %s\n", bfm_last_error(NULL)); return 2; }

    /* naive reference: 2-NN per query (ties to the lowest index) and the column minima */
    static int nn1[NQ], nn2[NQ], d1[NQ], d2[NQ], col[NT];
    for (int j = 0; j < NT; ++j) col[j] = -1;
    static int cold[NT];
    for (int i = 0; i < NQ; ++i) {
        nn1[i] = nn2[i] = -1; d1[i] = d2[i] = 1 << 30;
        for (int j = 0; j < NT; ++j) {
            const int d = hamming(q[i], t[j]);
            if (d < d1[i]) { d2[i] = d1[i]; nn2[i] = nn1[i]; d1[i] = d; nn1[i] = j; }
            else if (d < d2[i]) { d2[i] = d; nn2[i] = j; }
            if (col[j] < 0 || d < cold[j]) { col[j] = i; cold[j] = d; }
        }
    }
    bfm_options_t o;
    memset(&o, 0, sizeof o);
    o.k = 2; o.ratio = -1.0; o.max_distance = -1;
    static int32_t idx[NQ][2], dist[NQ][2];
    if (bfm_knn(h, BFM_MEM_HOST, &q[0][0], NQ, &t[0][0], NT, &o, &idx[0][0], &dist[0][0], NULL) != BFM_OK) {
        printf("bfm_knn: %s\n", bfm_last_error(h)); return 3;
    }
    for (int i = 0; i < NQ; ++i)
        if (idx[i][0] != nn1[i] || idx[i][1] != nn2[i] || dist[i][0] != d1[i] || dist[i][1] != d2[i]) {
            printf("knn mismatch at query %d\n", i); return 4;
        }
    o.k = 1; o.cross_check = 1; o.max_distance = 29;   /* slam/tracking.py:57: distance < 30 */
    static int32_t mq[NQ], mt[NQ], md[NQ];
    int32_t n = 0;
    if (bfm_match(h, BFM_MEM_HOST, &q[0][0], NQ, &t[0][0], NT, &o, mq, mt, md, &n, NULL) != BFM_OK) {
        printf("bfm_match: %s\n", bfm_last_error(h)); return 5;
    }
    int want = 0;
    for (int i = 0; i < NQ; ++i)
        if (col[nn1[i]] == i && d1[i] <= 29) {
            if (want >= n || mq[want] != i || mt[want] != nn1[i] || md[want] != d1[i]) { printf("match mismatch at query %d\n", i); return 6; }
            ++want;
        }
    if (want != n) { printf("match count %d != %d\n", (int)n, want); return 7; }
    bfm_launch_info_t li;
    bfm_get_launch_info(h, &li);
    /* misuse is reported, not thrown */
    o.k = 2;
    if (bfm_match(h, BFM_MEM_HOST, &q[0][0], NQ, &t[0][0], NT, &o, mq, mt, md, &n, NULL) != BFM_ERR_INVALID) { printf("expected BFM_ERR_INVALID\n"); return 8; }
    printf("c abi ok: %d mutual matches, %d kernel launch(es) per call, message for misuse: %s\n", want, li.kernels_launched, bfm_last_error(h));
    bfm_destroy(h);
    return 0;
}
