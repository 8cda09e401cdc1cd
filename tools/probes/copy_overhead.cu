// Per-copy overhead of chunked H2D uploads with a watermark after every chunk (pinned source).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
typedef CUresult (*WriteValue64)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);

int main() {
    const size_t total = 32u << 20;
    char *h, *d; unsigned long long *hm, *dm;
    CK(cudaMallocHost(&h, total)); CK(cudaMalloc(&d, total));
    CK(cudaMallocHost(&hm, 4096)); CK(cudaMalloc(&dm, 4096));
    cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    WriteValue64 wv = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuStreamWriteValue64", (void **)&wv, cudaEnableDefault, &qr));
    printf("cuStreamWriteValue64 entry point: %p (query %d)\n", (void *)wv, (int)qr);
    for (int chunks : {1, 4, 16, 64, 256}) {
        for (int mode = 0; mode < 5; ++mode) {
            // 0: data copies only; 1: + 16-byte memcpy watermark; 2: + cuStreamWriteValue64 watermark;
            // 3: two halves per chunk on two streams, memcpy watermark each; 4: two streams, write-value watermark
            float best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0, s0));
                if (mode >= 3) { CK(cudaStreamWaitEvent(s1, e0, 0)); }
                const size_t cs = total / chunks;
                for (int c = 0; c < chunks; ++c) {
                    if (mode < 3) {
                        CK(cudaMemcpyAsync(d + c * cs, h + c * cs, cs / 2, cudaMemcpyHostToDevice, s0));
                        CK(cudaMemcpyAsync(d + c * cs + cs / 2, h + c * cs + cs / 2, cs / 2, cudaMemcpyHostToDevice, s0));
                        if (mode == 1) CK(cudaMemcpyAsync(dm, hm + 2 * (c % 64), 16, cudaMemcpyHostToDevice, s0));
                        if (mode == 2) wv(s0, (CUdeviceptr)dm, (cuuint64_t)c, 0);
                    } else {
                        CK(cudaMemcpyAsync(d + c * cs, h + c * cs, cs / 2, cudaMemcpyHostToDevice, s0));
                        CK(cudaMemcpyAsync(d + c * cs + cs / 2, h + c * cs + cs / 2, cs / 2, cudaMemcpyHostToDevice, s1));
                        if (mode == 3) {
                            CK(cudaMemcpyAsync(dm, hm + 2 * (c % 64), 8, cudaMemcpyHostToDevice, s0));
                            CK(cudaMemcpyAsync(dm + 8, hm + 2 * (c % 64) + 1, 8, cudaMemcpyHostToDevice, s1));
                        } else {
                            wv(s0, (CUdeviceptr)dm, (cuuint64_t)c, 0);
                            wv(s1, (CUdeviceptr)(dm + 8), (cuuint64_t)c, 0);
                        }
                    }
                }
                CK(cudaEventRecord(e1, s0));
                if (mode >= 3) { CK(cudaEventRecord(e2, s1)); CK(cudaStreamWaitEvent(s0, e2, 0)); CK(cudaEventRecord(e1, s0)); }
                CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            printf("chunks %3d mode %d: %.3f ms  (%.1f GB/s, %.1f us per chunk over the 1-chunk time)\n", chunks, mode, best,
                   total / best / 1e6, 0.0);
        }
    }
    return 0;
}
