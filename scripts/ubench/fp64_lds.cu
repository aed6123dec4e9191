// Do FP64 arithmetic and shared-memory loads overlap on an sm_100a SM?  Warps run NF DFMAs and NL LDS.128 per
// iteration (independent of each other); MODE 1 = DFMA only, 2 = LDS only, 3 = both in every warp,
// 4 = even warps DFMA / odd warps LDS (same totals as mode 3 at twice the per-warp counts).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    double f[16]; double2 acc = make_double2(0, 0);
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = i * 0.5 + threadIdx.x;
    const bool wf = MODE == 1 || MODE == 3 || (MODE == 4 && !((threadIdx.x >> 5) & 1));
    const bool wl = MODE == 2 || MODE == 3 || (MODE == 4 && ((threadIdx.x >> 5) & 1));
    const int rep = MODE == 4 ? 2 : 1;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (wf) {
#pragma unroll
            for (int r = 0; r < 4 * rep; r++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
        }
        if (wl) {
#pragma unroll
            for (int r = 0; r < 16 * rep; r++) {
                double2 v;
                asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(&sm[(idx + r * 256) & 4095])));
                acc.x += 0 * v.x; acc.y = v.y;
            }
            idx += 7;
        }
    }
    double s = acc.x + acc.y;
#pragma unroll
    for (int i = 0; i < 16; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(double* out, int iters, int grid) {
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256, 65536>>>(out, iters, 0.999, 1e-3); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<MODE><<<grid, 256, 65536>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * nsm * 2 * 256);
    const int iters = 4000, grid = nsm * 2;  // 16 warps per SM
    float t1 = run<1>(out, iters, grid), t2 = run<2>(out, iters, grid), t3 = run<3>(out, iters, grid), t4 = run<4>(out, iters, grid);
    printf("per iteration and warp: 64 DFMA (2 cycles each per scheduler) and 16 LDS.128 (4 wavefronts each); 16 warps per SM\n");
    printf("DFMA only %.3f ms | LDS.128 only %.3f ms | both in every warp %.3f ms | split over warps %.3f ms   (sum %.3f, max %.3f)\n", t1, t2, t3, t4, t1 + t2, t1 > t2 ? t1 : t2);
    return 0;
}
