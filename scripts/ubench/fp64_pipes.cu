// Microbenchmark: do DMMA (mma.sync.m8n8k4.f64) and DFMA share an execution pipe on sm_100a?
// MODE 1 = DMMA only, 2 = DFMA only, 3 = both interleaved in the same warp, 4 = alternate warps (even DMMA, odd DFMA)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    double c[8][2]; double f[16];
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = i * 0.5;
    const bool do_mma = (MODE == 1) || (MODE == 3) || (MODE == 4 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_fma = (MODE == 2) || (MODE == 3) || (MODE == 4 && ((threadIdx.x >> 5) & 1) == 1);
    for (int it = 0; it < iters; it++) {
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        if (do_fma) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(double* out, int iters, int grid) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, iters, 0.5); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, iters, 0.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * 4, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * grid * 256);
    float t1 = run<1>(out, iters, grid), t2 = run<2>(out, iters, grid), t3 = run<3>(out, iters, grid), t4 = run<4>(out, iters, grid);
    double warps = (double)grid * 8;
    double mma_flops = warps * iters * 8 * 512.0, fma_flops = warps * iters * 64 * 64.0;
    printf("SMs %d\n", nsm);
    printf("DMMA only     %.3f ms  %.2f TFLOP/s\n", t1, mma_flops / t1 * 1e-9);
    printf("DFMA only     %.3f ms  %.2f TFLOP/s\n", t2, fma_flops / t2 * 1e-9);
    printf("both, same warp  %.3f ms  (sum %.3f, max %.3f)\n", t3, t1 + t2, t1 > t2 ? t1 : t2);
    printf("alternate warps (half the work of each)  %.3f ms  (sum/2 %.3f, max/2 %.3f)\n", t4, (t1 + t2) / 2, (t1 > t2 ? t1 : t2) / 2);
    return 0;
}
