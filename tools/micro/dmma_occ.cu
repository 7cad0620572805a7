// Micro-benchmark (diagnostic): FP64 tensor-pipe (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) throughput of ONE block per SM as a
// function of its warps and of the independent accumulator tiles per warp -- the denominator that applies to K4's shape
// (one 8-warp CTA per SM, 4 .. 16 accumulator chains per warp), next to the pool-wide peak that nbk_fp64_mma_peak
// measures with 64 warps per SM.  Also with the operands re-loaded from shared memory before every DMMA (as K4 must).
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, bool LOADS>
__global__ void k(double *sink, int iters, double a0, double b0) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    double a = a0, b = b0;
    const double *p = sm + (threadIdx.x & 31);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (LOADS) {
                a = p[(it * NACC + i) * 8 & 1023];
                b = p[(((it * NACC + i) * 8) & 1023) + 512];
            }
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1];
    if (r == 123.456) sink[0] = r;
}

template <int NACC, bool LOADS>
void run(int sms, int warps, double *sink) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<NACC, LOADS><<<sms, 32 * warps, 2048 * 8>>>(sink, iters / 10, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<NACC, LOADS><<<sms, 32 * warps, 2048 * 8>>>(sink, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * 8 * 4 * (double)NACC * iters * warps * sms;
    printf("warps/SM %2d  accumulators/warp %2d  %s: %6.2f TFLOP/s\n", warps, NACC, LOADS ? "operands from shared memory" : "operands in registers    ",
           flop / ms * 1e3 / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *sink;
    cudaMalloc(&sink, 8);
    for (int warps : {4, 8, 16, 32}) {
        run<4, false>(p.multiProcessorCount, warps, sink);
        run<8, false>(p.multiProcessorCount, warps, sink);
        run<16, false>(p.multiProcessorCount, warps, sink);
        run<4, true>(p.multiProcessorCount, warps, sink);
        run<16, true>(p.multiProcessorCount, warps, sink);
    }
    return 0;
}
