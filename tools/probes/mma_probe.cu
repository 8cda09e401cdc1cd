// throughput of the legacy (mma.sync) tensor path on this GPU: s8 m16n8k32, e4m3 m16n8k32, f16 m16n8k16
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND>
__global__ void __launch_bounds__(256) probe(int iters, int *out) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int c[8][4];
    float f[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) { c[i][j] = 0; f[i][j] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k64.row.col.s32.s4.s4.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j] + (int)f[i][j];
    if (s == 0x12345678) out[0] = s;
}
template <int KIND>
void run(const char *name, double ops_per_mma, int ctas_per_sm) {
    int *out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = 148 * ctas_per_sm;
    probe<KIND><<<grid, 256>>>(100, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<KIND><<<grid, 256>>>(iters, out);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)grid * 8 /*warps*/ * iters * 8;
    std::printf("%-28s %d CTAs/SM: %.3f ms, %.1f TOP/s, %.0f ops/clk/SM at 1.965 GHz (%s)\n", name, ctas_per_sm, ms, mmas * ops_per_mma / ms / 1e9,
                mmas * ops_per_mma / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    for (int c : {1, 2, 4}) {
        run<0>("s8 m16n8k32", 2.0 * 16 * 8 * 32, c);
        run<1>("e4m3 m16n8k32", 2.0 * 16 * 8 * 32, c);
        run<2>("f16 m16n8k16", 2.0 * 16 * 8 * 16, c);
        run<3>("s4 m16n8k64", 2.0 * 16 * 8 * 64, c);
    }
    return 0;
}
