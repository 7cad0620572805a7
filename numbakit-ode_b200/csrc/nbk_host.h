// nbk_host.h -- host-side internals shared by nbk_api.cu, nbk_jit.cpp and nbk_program.cu.
#ifndef NBK_HOST_H
#define NBK_HOST_H

struct nbk_static_program {
    int method;            // 23, 45, 1..5
    int n, n_params;       // n == -1: any (run-time) size, CTA regime
    int cta_elems;         // 0: one trajectory per thread; E > 0: one CTA per trajectory, E components per thread
    int cta_full;          // CTA regime: 1 = compiled without bounds checks, needs n == block * E; 0 = any n
    int group;             // CTA regime: G > 0 = sub-warp run / step kernels (G lanes per trajectory, 128-thread persistent blocks)
    int mma_tile;          // > 0: K4, one CTA integrates a tile of that many trajectories (shared-matrix linear)
    int mma_block;         // K4: threads per CTA (32 x warps of the tile kernel)
    unsigned long mma_smem;  // dynamic shared memory of the K4 run/step kernels
    const void *k_init;    // host stubs of the __global__ kernels (cudaLaunchKernel)
    const void *k_run;
    const void *k_step;
};

#endif
