// Micro-benchmark (diagnostic): which SM sub-partition does a warp run on?  4 CTAs of 4 warps per SM (the K1 launch
// shape); in every CTA only ONE warp does FP64 work (the others exit at once).  If the working warp of all four CTAs
// of an SM has the same index, they share one sub-partition (one FP64 pipe) when warps are placed by (warp id mod 4);
// if the index differs per CTA they spread over the four pipes.  Decides which warp should survive the tail compaction.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 4) k(double *sink, int iters, double a, double b, int mode, int sms) {
    const int warp = threadIdx.x >> 5;
    const int want = mode == 0 ? 0 : (mode == 1 ? (int)(blockIdx.x / sms) & 3 : (int)(blockIdx.x & 3));
    if (warp != want) return;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i];
    if (r == 123.456) sink[0] = r;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 4000;
    double *sink;
    cudaMalloc(&sink, 8);
    const char *names[] = {"warp 0 of every CTA", "warp (blockIdx / #SM) mod 4", "warp blockIdx mod 4"};
    for (int mode = 0; mode < 3; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k<<<sms * 4, 128>>>(sink, iters / 10, 0.999999, 1e-9, mode, sms);
        cudaEventRecord(e0);
        k<<<sms * 4, 128>>>(sink, iters, 0.999999, 1e-9, mode, sms);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-30s: %7.3f ms (4 working warps per SM; 64 DFMA x %d iterations each)\n", names[mode], ms, iters);
    }
    return 0;
}
