// Micro-benchmark (diagnostic): issue cost of FP64 instruction flavours on B200 at 1 and 4 warps per SM
// sub-partition: register operands vs uniform-register (constant bank) operands, DFMA vs DMUL vs DADD.
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double cc[64];

template <int MODE>
__global__ void __launch_bounds__(128) k(double *sink, int iters, double a, double b) {
    double x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = 1.0 + 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);                       // register operands
                if (MODE == 1) x[i] = fma(x[i], cc[u * 8 + i], b);           // one constant-bank (UR) operand, 64 distinct
                if (MODE == 2) x[i] = x[i] * a;                              // DMUL
                if (MODE == 3) x[i] = x[i] + a;                              // DADD
                if (MODE == 4) x[i] = fma(x[i], y[(i + u) & 7], y[(i + 3) & 7]);  // three per-thread registers
                if (MODE == 5) x[i] = fma(x[i], cc[i], cc[8 + i]);           // two UR operands, 16 distinct (stay resident)
                if (MODE == 6) x[i] = x[i] * cc[u * 8 + i];                  // DMUL with UR operand
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += x[i] + y[i];
    if (r == 123.456) sink[0] = r;
}

template <int MODE>
void run(int sms, int w, double *sink, const char *name) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms * w, 128>>>(sink, iters / 10, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<sms * w, 128>>>(sink, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 64 * w);  // cycles per warp-instruction per SMSP
    printf("%-34s warps/SMSP %d : %7.3f ms  %.2f cycles per warp instr\n", name, w, ms, cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.999999 + 1e-9 * i;
    cudaMemcpyToSymbol(cc, h, sizeof h);
    double *sink;
    cudaMalloc(&sink, 8);
    for (int w : {1, 4}) {
        run<0>(sms, w, sink, "DFMA reg,reg,reg");
        run<1>(sms, w, sink, "DFMA reg,UR(64 distinct),reg");
        run<5>(sms, w, sink, "DFMA reg,UR,UR (16 resident)");
        run<2>(sms, w, sink, "DMUL reg,reg");
        run<6>(sms, w, sink, "DMUL reg,UR(64 distinct)");
        run<3>(sms, w, sink, "DADD reg,reg");
        run<4>(sms, w, sink, "DFMA three per-thread regs");
    }
    return 0;
}
