// Microbenchmark: are the FP64 FMA pipe (DFMA) and the FP64 tensor sub-pipe (DMMA m16n8k4) of B200 separate
// execution resources?  Times DFMA alone, DMMA alone, and both interleaved in the same warps (NF DFMAs per DMMA).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu ; run: ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b0));
}

// per loop iteration and warp: ND DMMAs (independent accumulators) and NF DFMAs per lane (independent accumulators)
template <int ND, int NF>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, double x) {
    double acc[ND > 0 ? ND : 1][4];
    double f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < (ND > 0 ? ND : 1); ++i)
        acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i)
        f[i] = threadIdx.x * 1e-7 + i;
    const double a0 = x, a1 = x * 0.5, b0 = 1.0 - x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i)
            dmma(acc[i], a0, a1, b0);
#pragma unroll
        for (int i = 0; i < NF; ++i)
            asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(a0), "d"(b0));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < (ND > 0 ? ND : 1); ++i)
        s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i)
        s += f[i];
    if (s == 123.456)
        out[0] = s;
}

template <int ND, int NF> void run(const char* name, int warps_per_sm) {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
    mix_kernel<ND, NF><<<blocks, threads>>>(out, 100, 0.25);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mix_kernel<ND, NF><<<blocks, threads>>>(out, iters, 0.25);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = double(blocks) * threads / 32;
    const double fl_d = warps * iters * ND * 1024.0, fl_f = warps * iters * NF * 64.0;
    printf("%-28s warps/SM %2d: %8.3f ms  DMMA %6.2f TF  DFMA %6.2f TF  total %6.2f TF\n", name, warps_per_sm, ms,
           fl_d / ms / 1e9, fl_f / ms / 1e9, (fl_d + fl_f) / ms / 1e9);
    cudaFree(out);
}

int main() {
    for (int w : {8, 16, 32}) {
        run<8, 0>("DMMA only (8/iter)", w);
        run<0, 32>("DFMA only (32/iter)", w);
        run<8, 32>("8 DMMA + 32 DFMA", w);
        run<8, 64>("8 DMMA + 64 DFMA", w);
        run<8, 128>("8 DMMA + 128 DFMA", w);
        run<4, 128>("4 DMMA + 128 DFMA", w);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
