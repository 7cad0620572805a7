// nbk_program_mma.cu -- init / run / step kernels of the shared-matrix linear program (K4,
// nbk_mma.cuh): y' = A y, n <= 128, RungeKutta23 / RungeKutta45, right-hand side as an FP64 DMMA GEMM
// over tiles of 16 trajectories.  Macros: NBK_METHOD (23 | 45), NBK_TAG.  Block = 256 threads.
#include "nbk_mma.cuh"

#define NBK_CAT2(a, b) a##b
#define NBK_CAT(a, b) NBK_CAT2(a, b)
#define NBK_KERNEL(stem) NBK_CAT(NBK_CAT(nbk_, stem), NBK_CAT(_, NBK_TAG))

#if NBK_METHOD == 45
typedef nbk::tab_rk45 nbk_tab_t;
#elif NBK_METHOD == 23
typedef nbk::tab_rk23 nbk_tab_t;
#else
#error "the shared-matrix linear program implements NBK_METHOD 23 or 45"
#endif

// construction is the CTA regime's (same trajectory-major state): one CTA per trajectory, E = 1
extern "C" __global__ void __launch_bounds__(nbk::MMA_THREADS) NBK_KERNEL(init)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ double nbk_smem[];
    nbk::cta_init_adaptive<nbk_tab_t, nbk::rhs_linear_cta, 1>(a, nbk_smem);
}
extern "C" __global__ void __launch_bounds__(nbk::MMA_THREADS, 1) NBK_KERNEL(run)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ double nbk_smem[];
    nbk::mma_advance<nbk_tab_t, 128, NBK_MODE_RUN>(a, nbk_smem);
}
extern "C" __global__ void __launch_bounds__(nbk::MMA_THREADS, 1) NBK_KERNEL(step)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ double nbk_smem[];
    nbk::mma_advance<nbk_tab_t, 128, NBK_MODE_STEP>(a, nbk_smem);
}

#ifndef __CUDACC_RTC__
#include "nbk_host.h"
extern "C" void NBK_KERNEL(program)(nbk_static_program *out) {
    out->method = NBK_METHOD;
    out->n = -1;
    out->n_params = -1;
    out->cta_elems = 1;
    out->cta_full = 0;
    out->group = 0;
    out->mma_tile = nbk::MMA_NT;
    out->mma_block = nbk::MMA_THREADS;
    // per group: Y staging (2 buffers), reduction scratch, columns; then the one copy of the matrix (see mma_advance)
    out->mma_smem = sizeof(double) * nbk::MMA_GROUPS * (2 * 128 * nbk::MMA_LD + nbk::MMA_WARPS * nbk::MMA_NT + (sizeof(nbk::mma_cols) + 7) / 8 + 2) +
                    16 + sizeof(double) * 128 * nbk::MMA_LDA;
    out->k_init = (const void *)&NBK_KERNEL(init);
    out->k_run = (const void *)&NBK_KERNEL(run);
    out->k_step = (const void *)&NBK_KERNEL(step);
}
#endif
