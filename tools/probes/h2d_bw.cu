// tools/probes/h2d_bw.cu - what the copy engine delivers from pinned host memory (one copy and chunked copies), next to
// the SM-fed upload's ~30 GB/s.  nvcc -O2 -o _ab/h2d_bw tools/probes/h2d_bw.cu ; gpurun -- _ab/h2d_bw
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
int main() {
    const size_t n = 32u << 20;
    void *h, *d, *h2;
    cudaHostAlloc(&h, n, cudaHostAllocDefault); cudaHostAlloc(&h2, n, cudaHostAllocWriteCombined); cudaMalloc(&d, n);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (void *src : {h, h2})
        for (int chunks : {1, 4, 8, 16, 64}) {
            float best = 1e9f; double best_wall = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaStreamSynchronize(s);
                auto t0 = std::chrono::steady_clock::now();
                cudaEventRecord(a, s);
                for (int c = 0; c < chunks; ++c) cudaMemcpyAsync((char *)d + c * (n / chunks), (char *)src + c * (n / chunks), n / chunks, cudaMemcpyHostToDevice, s);
                cudaEventRecord(b, s);
                cudaStreamSynchronize(s);
                auto t1 = std::chrono::steady_clock::now();
                float ms; cudaEventElapsedTime(&ms, a, b);
                best = ms < best ? ms : best;
                best_wall = std::min(best_wall, std::chrono::duration<double, std::milli>(t1 - t0).count());
            }
            printf("%s H2D 32 MiB in %2d chunks: %.3f ms device (%.1f GB/s), %.3f ms wall\n", src == h ? "pinned        " : "write-combined", chunks, best, n / best / 1e6, best_wall);
        }
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(a, s); cudaMemcpyAsync(h, d, 4u << 20, cudaMemcpyDeviceToHost, s); cudaEventRecord(b, s); cudaStreamSynchronize(s);
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    printf("D2H 4 MiB: %.3f ms (%.1f GB/s)\n", best, (4u << 20) / best / 1e6);
    return 0;
}
