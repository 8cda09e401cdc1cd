// tools/tc_bench.cu - stand-alone harness of the tensor-core matching kernel (boslam_b200/csrc/bfm_tensor.cuh): checks
// its row states against a brute-force POPC kernel on ragged shapes and on the 256 x 2000^2 headline batch, then times
// expansion and scan.  Build (the binary is not committed):
//   nvcc -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o _ab/tc/tc_bench tools/tc_bench.cu
// Run on the GPU box:  gpurun -- 'timeout 100 _ab/tc/tc_bench'   ("quick" as first argument: correctness cases only)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>
#define BFM_TC_HARNESS 1
#include "../boslam_b200/csrc/bfm_tensor.cuh"
using namespace bfm_tc;

__global__ void ref_kernel(const uint32_t *q, const uint32_t *t, int nq, int nt, unsigned long long *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t a[8];
    for (int w = 0; w < 8; ++w) a[w] = q[(size_t)i * 8 + w];
    uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
    for (int j = 0; j < nt; ++j) {
        int h = 0;
        for (int w = 0; w < 8; ++w) h += __popc(a[w] ^ t[(size_t)j * 8 + w]);
        const uint32_t key = ((uint32_t)h << 22) | (uint32_t)j;
        b2 = min(b2, max(b1, key));
        b1 = min(b1, key);
    }
    out[i] = ((unsigned long long)b1 << 32) | b2;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <int NST, int NA = 2>
static int run_case(int P, int nq, int nt, bool check, int reps, int dbg = 0) {
    const size_t QR = (size_t)P * nq, TR = (size_t)P * nt;
    std::vector<uint32_t> hq(QR * 8), ht(TR * 8);
    srand(7 + P + nq + nt);
    for (auto &v : hq) v = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
    for (auto &v : ht) v = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
    // a few near-duplicates so that small distances and ties occur
    for (int p = 0; p < P; ++p)
        for (int i = 0; i < std::min(nq, nt); i += 3) {
            for (int w = 0; w < 8; ++w) ht[((size_t)p * nt + i) * 8 + w] = hq[((size_t)p * nq + (i * 7) % nq) * 8 + w];
            ht[((size_t)p * nt + i) * 8 + (i & 7)] ^= (uint32_t)(i * 2654435761u) & 0x0F0Fu;
        }
    std::vector<XProblem> probs(P);
    std::vector<Item> items;
    size_t xq_rows = 0, xt_rows = 0;
    for (int p = 0; p < P; ++p) {
        probs[p] = {p * nq, nq, p * nt, nt, (int)xq_rows, (int)xt_rows};
        for (int b = 0; b * BQ < nq; ++b) items.push_back({(int)xq_rows + b * BQ, std::min(BQ, nq - b * BQ), b * BQ, p * nq + b * BQ, (int)xt_rows, nt, 0, p});
        xq_rows += (nq + 7) & ~7;
        xt_rows += (nt + 7) & ~7;
    }
    const size_t xq_plane = (xq_rows + 512) * 128, xt_plane = (xt_rows + 512) * 128;
    uint32_t *dq, *dt; uint8_t *xq, *xt; XProblem *dp; Item *di; unsigned long long *state, *refst; uint32_t *status;
    CK(cudaMalloc(&dq, hq.size() * 4)); CK(cudaMalloc(&dt, ht.size() * 4));
    CK(cudaMalloc(&xq, 2 * xq_plane)); CK(cudaMalloc(&xt, 2 * xt_plane));
    CK(cudaMalloc(&dp, P * sizeof(XProblem))); CK(cudaMalloc(&di, items.size() * sizeof(Item)));
    CK(cudaMalloc(&state, QR * 8)); CK(cudaMalloc(&refst, QR * 8)); CK(cudaMalloc(&status, 4));
    CK(cudaMemcpy(dq, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dp, probs.data(), P * sizeof(XProblem), cudaMemcpyHostToDevice)); CK(cudaMemcpy(di, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice));
    CK(cudaMemset(xq, 0, 2 * xq_plane)); CK(cudaMemset(xt, 0, 2 * xt_plane)); CK(cudaMemset(status, 0, 4));
    CK(cudaFuncSetAttribute(scan_kernel<NST, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(NST, NA)));
    Params pa{xq, xt, xq_plane, xt_plane, di, (int)items.size(), state, status, dbg, (uint32_t)(-64), (uint32_t)(-64) << 16};
    const int grid = std::min<int>((int)items.size(), 148);
    const dim3 egrid((std::max(nq, nt) + EXPAND_ROWS - 1) / EXPAND_ROWS, P, 2);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float best_x = 1e9, best_s = 1e9;
    for (int rep = 0; rep < reps; ++rep) {
        CK(cudaMemset(state, 0xFF, QR * 8));
        CK(cudaEventRecord(e0));
        expand_kernel<<<egrid, 256>>>((const uint16_t *)dq, (const uint16_t *)dt, (const int32_t *)dp, 6, 4, 5, (uint4 *)xq, (uint4 *)xt, xq_plane / 16, xt_plane / 16);
        CK(cudaEventRecord(e1));
        scan_kernel<NST, NA><<<grid, NTHREADS, smem_bytes(NST, NA)>>>(pa);
        CK(cudaEventRecord(e2));
        CK(cudaDeviceSynchronize());
        float mx, ms; cudaEventElapsedTime(&mx, e0, e1); cudaEventElapsedTime(&ms, e1, e2);
        best_x = std::min(best_x, mx); best_s = std::min(best_s, ms);
    }
    uint32_t st = 0; CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
    const double pairs = (double)P * nq * nt;
    printf("[stages %d A x%d dbg %d] P=%d %dx%d: %zu items on %d CTAs | expand %.1f us, scan %.1f us -> %.0f G pairs/s (scan), %.0f G pairs/s (both) | status %u\n", NST, NA, dbg, P, nq, nt, items.size(), grid,
           best_x * 1e3, best_s * 1e3, pairs / best_s / 1e6, pairs / (best_s + best_x) / 1e6, st);
    if (check) {
        for (int p = 0; p < P; ++p)
            ref_kernel<<<(nq + 127) / 128, 128>>>(dq + (size_t)p * nq * 8, dt + (size_t)p * nt * 8, nq, nt, refst + (size_t)p * nq);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> a(QR), b(QR);
        CK(cudaMemcpy(a.data(), state, QR * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), refst, QR * 8, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < QR; ++i)
            if (a[i] != b[i]) { if (bad < 6) printf("   row %zu: got %016llx want %016llx\n", i, a[i], b[i]); ++bad; }
        printf("   mismatching rows: %zu of %zu\n", bad, QR);
    }
    cudaFree(dq); cudaFree(dt); cudaFree(xq); cudaFree(xt); cudaFree(dp); cudaFree(di); cudaFree(state); cudaFree(refst); cudaFree(status);
    return 0;
}

int main(int argc, char **argv) {
    const std::string mode = argc > 1 ? argv[1] : "all";
    if (mode == "ncu") return run_case<3>(256, 2000, 2000, false, 2);   // one shape, for a profiler capture
    if (run_case<2>(1, 256, 128, true, 1)) return 1;
    if (run_case<2>(1, 300, 1000, true, 1)) return 1;
    if (run_case<3>(3, 2000, 2000, true, 2)) return 1;
    if (run_case<3>(2, 777, 1234, true, 1)) return 1;
    if (mode == "quick") return 0;
    if (run_case<3>(20, 2000, 2000, false, 5)) return 1;
    if (run_case<3>(256, 2000, 2000, true, 5)) return 1;
    if (run_case<3>(1, 2000, 20000, false, 5)) return 1;
    if (run_case<3>(1, 8192, 8192, false, 5)) return 1;
    if (mode != "probes") return 0;
    // which of the three engines binds: each alone, and pairs (bits: 1 no folds, 2 no MMA, 4 no B loads)
    for (int dbg : {8, 5, 6, 7}) if (run_case<3>(256, 2000, 2000, false, 5, dbg)) return 1;
    if (run_case<2>(256, 2000, 2000, false, 5, 0)) return 1;        // two B stages
    if (run_case<5, 1>(256, 2000, 2000, false, 5, 0)) return 1;     // five B stages, one A buffer
    return 0;
}
