// Micro-benchmark (diagnostic): does an FP64 warp instruction cost less when only some of the warp's lanes are active?
// Every warp runs the same chain of independent DFMAs; lanes outside the mask skip the loop (divergent branch, so the
// warp executes the loop with a partial active mask).  Reported: cycles per warp-level DFMA per SM sub-partition, at
// 4 warps per sub-partition (the occupancy of the K1 integrator kernel), for several active-lane patterns.
// If the time follows the number of active lanes, idle lanes of a divergent warp cost no FP64-pipe time -- which decides
// whether compacting the surviving trajectories of the adaptive kernels across warps can pay at all.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(double *sink, int iters, double a, double b, unsigned mask) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    if ((mask >> (threadIdx.x & 31)) & 1u) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i];
    if (r == 123.456) sink[0] = r;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, w = 4, iters = 4000;
    double *sink;
    cudaMalloc(&sink, 8);
    struct { unsigned mask; const char *name; } cases[] = {
        {0xffffffffu, "32 lanes"},          {0x0000ffffu, "16 lanes (lower half)"}, {0x55555555u, "16 lanes (every other)"},
        {0x000000ffu, "8 lanes (0-7)"},     {0x11111111u, "8 lanes (every 4th)"},   {0x0000000fu, "4 lanes (0-3)"},
        {0x01010101u, "4 lanes (every 8th)"}, {0x00000001u, "1 lane"},              {0x7fffffffu, "31 lanes"},
        {0x0fffffffu, "28 lanes (0-27)"},   {0x00ffffffu, "24 lanes (0-23)"},
    };
    for (auto &c : cases) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k<<<sms * w, 128>>>(sink, iters / 10, 0.999999, 1e-9, c.mask);
        cudaEventRecord(e0);
        k<<<sms * w, 128>>>(sink, iters, 0.999999, 1e-9, c.mask);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 64 * w);
        printf("%-26s: %7.3f ms  %.2f cycles per warp-level DFMA per sub-partition\n", c.name, ms, cyc);
    }
    return 0;
}
