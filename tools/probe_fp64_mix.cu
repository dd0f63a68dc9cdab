// Does the FP64 tensor instruction (DMMA m8n8k4) share the DFMA pipe on B200?  Three timed kernels of register-only work:
// all warps DFMA, all warps DMMA, and half / half.  If the mixed run reaches the sum of the two rates the pipes are separate
// (and the dense-block contractions could be split between them); if it stays at the single-pipe rate they are one resource.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_fp64_mix tools/probe_fp64_mix.cu && build/probe_fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) mix_kernel(int iters, int mode, double *out)
{
    const int warp = threadIdx.x >> 5;
    const bool use_mma = mode == 1 || (mode == 2 && (warp & 1));
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16];
    for (int t = 0; t < 16; t++) c[t] = t * 1e-3;
    if (use_mma) {
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int t = 0; t < 8; t++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[2 * t]), "+d"(c[2 * t + 1]) : "d"(a), "d"(b));
    } else {
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int r = 0; r < 4; r++)      // 64 DFMA per thread and iteration = 8 DMMA tiles' worth of FMAs per warp
#pragma unroll
                for (int t = 0; t < 16; t++) c[t] = fma(a, c[t], b);
    }
    double r = 0.0;
    for (int t = 0; t < 16; t++) r += c[t];
    if (r == 123.456) out[0] = r;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15, grid = sms * 8;
    const char *name[3] = {"all DFMA", "all DMMA", "half DFMA + half DMMA"};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            mix_kernel<<<grid, 256>>>(iters, mode, d);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double flop = 2.0 * 64.0 * 32.0 * 8.0 * (double)iters * grid;   // 64 FMA per thread-iteration in both forms
        printf("%-24s %8.3f ms  %6.2f TFLOP/s\n", name[mode], best, flop / (best * 1e-3) / 1e12);
    }
    return 0;
}
