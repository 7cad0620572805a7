// Micro-benchmark (diagnostic, not product): FP64 pipe utilisation on B200 as a function of resident warps per
// SM sub-partition and independent DFMA chains per thread (ILP), optionally with ALU instructions interleaved.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_ilp fp64_ilp.cu ; run: ./fp64_ilp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MIX>
__global__ void __launch_bounds__(128) k(double *sink, int iters, double a, double b, int salt) {
    double x[ILP];
    int z = threadIdx.x ^ salt;
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                x[i] = fma(x[i], a, b);
                if (MIX) {
#pragma unroll
                    for (int m = 0; m < MIX; ++m) z = (z ^ (z >> 3)) + salt;  // 2 ALU ops (SHF/LOP3+IADD)
                }
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += x[i];
    if (r == 123.456 || z == 0x7fffffff) sink[0] = r + z;
}

template <int ILP, int MIX>
void run(int sms, int ctas_per_sm, double *sink) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP, MIX><<<sms * ctas_per_sm, 128>>>(sink, iters / 10, 0.999999, 1e-9, 12345);
    cudaEventRecord(e0);
    k<ILP, MIX><<<sms * ctas_per_sm, 128>>>(sink, iters, 0.999999, 1e-9, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)iters * 16 * ILP * sms * ctas_per_sm * 128;
    const double tf = 2 * dfma / (ms * 1e-3) / 1e12;
    // cycles per warp-DFMA per SM sub-partition
    printf("ILP %d MIX %d warps/SMSP %2d : %7.3f ms  %6.2f TFLOP/s\n", ILP, MIX, ctas_per_sm, ms, tf);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *sink;
    cudaMalloc(&sink, 8);
    printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    for (int w : {1, 2, 3, 4, 5, 6, 8, 12, 16}) {
        run<1, 0>(sms, w, sink);
        run<2, 0>(sms, w, sink);
        run<3, 0>(sms, w, sink);
        run<4, 0>(sms, w, sink);
        run<8, 0>(sms, w, sink);
    }
    for (int w : {2, 4, 6, 8}) {
        run<3, 1>(sms, w, sink);
        run<3, 2>(sms, w, sink);
        run<8, 1>(sms, w, sink);
    }
    return 0;
}
