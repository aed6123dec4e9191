// FP64 pipe on sm_100a: throughput of DFMA / DADD streams vs. warps per SM and independent chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP, int OP>
__global__ void k(double* out, int iters, double a, double b) {
    double f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) f[i] = i * 0.5 + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 64 / ILP; r++)
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
            else asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(f[i]) : "d"(a));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int OP> void run(double* out, int nsm, int warps_per_sm) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<ILP, OP><<<nsm, warps_per_sm * 32>>>(out, iters, 0.999, 1e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // cycles per warp-instruction per SMSP
    double instr_per_smsp = (double)iters * 64 * warps_per_sm / 4.0;
    double cyc = best * 1e-3 * 1.965e9;
    printf("%s ILP %d warps/SM %2d : %.3f ms  %.2f cyc per warp-instr per SMSP (pipe busy %.0f%% if 2 cyc each)\n", OP ? "DADD" : "DFMA", ILP, warps_per_sm, best,
           cyc / instr_per_smsp, 200.0 * instr_per_smsp / cyc);
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * nsm * 1024);
    for (int w : {4, 8, 12, 16}) {
        if (w == 4) { run<1,0>(out,nsm,4); run<2,0>(out,nsm,4); run<4,0>(out,nsm,4); run<8,0>(out,nsm,4); run<1,1>(out,nsm,4); run<4,1>(out,nsm,4);}
        if (w == 8) { run<1,0>(out,nsm,8); run<2,0>(out,nsm,8); run<4,0>(out,nsm,8); run<8,0>(out,nsm,8); run<2,1>(out,nsm,8);}
        if (w == 12) { run<1,0>(out,nsm,12); run<2,0>(out,nsm,12); run<4,0>(out,nsm,12);}
        if (w == 16) { run<1,0>(out,nsm,16); run<2,0>(out,nsm,16); run<4,0>(out,nsm,16);}
    }
    return 0;
}
