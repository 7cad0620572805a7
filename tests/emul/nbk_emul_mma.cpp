// nbk_emul_mma.cpp -- TEST INFRASTRUCTURE: host build of the shared-matrix linear program (K4, csrc/nbk_mma.cuh +
// nbk_program_mma.cu).  A block is 256 POSIX threads; __syncthreads is a block barrier; warp shuffles and the DMMA
// fragment exchange (mma.sync.m8n8k4: A(row, k) in lane 4 row + k, B(k, col) in lane 4 col + k, C(row, 2 (lane % 4) + q)
// in lane 4 row + lane % 4) go through per-warp scratch arrays between per-warp barriers.  Checks the tile logic --
// per-column controllers, frozen columns, partial tiles, emission from the accumulator layout -- in `pytest -m "not gpu"`,
// with the threads scheduled by the OS (any missing barrier shows as a run-to-run difference).
#include <pthread.h>

#include "cuda_host_shim.h"

static thread_local emul_dim3 gridDim;
struct double2 { double x, y; };
#define __shared__
#define __align__(x)

static pthread_barrier_t g_bar, g_wbar[8];
static double g_shfl[256], g_ma[256], g_mb[256];
alignas(16) double nbk_smem[1 << 16];

static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static int g_or[2];
static thread_local unsigned g_or_phase;
static inline int __syncthreads_or(int p) {  // barrier + OR over the block; two accumulators used alternately
    const unsigned k = g_or_phase++ & 1u;
    if (p) __atomic_store_n(&g_or[k], 1, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&g_bar);
    const int r = __atomic_load_n(&g_or[k], __ATOMIC_SEQ_CST);
    if (threadIdx.x == 0) __atomic_store_n(&g_or[k ^ 1u], 0, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&g_bar);
    return r;
}
static inline double __shfl_xor_sync(unsigned, double v, int o) {
    const int t = (int)threadIdx.x;
    g_shfl[t] = v;
    pthread_barrier_wait(&g_wbar[t >> 5]);
    const double r = g_shfl[t ^ o];
    pthread_barrier_wait(&g_wbar[t >> 5]);
    return r;
}
// the groups' named barriers (csrc/nbk_mma.cuh, NBK_MMA_GROUPS = 2): 128 threads each; with one group = the block barrier
static pthread_barrier_t g_gbar[2];
static int g_gor[2][2];
static thread_local unsigned g_gor_phase;
static int g_gthreads = 256;
static inline void nbk_emul_bar(int grp) { pthread_barrier_wait(g_gthreads == 256 ? &g_bar : &g_gbar[grp]); }
static inline int nbk_emul_bar_or(int grp, int p) {
    if (g_gthreads == 256) return __syncthreads_or(p);
    const unsigned k = g_gor_phase++ & 1u;
    if (p) __atomic_store_n(&g_gor[grp][k], 1, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&g_gbar[grp]);
    const int r = __atomic_load_n(&g_gor[grp][k], __ATOMIC_SEQ_CST);
    if ((int)threadIdx.x % g_gthreads == 0) __atomic_store_n(&g_gor[grp][k ^ 1u], 0, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&g_gbar[grp]);
    return r;
}
static inline void nbk_emul_dmma(double &c0, double &c1, double a, double b) {
    const int t = (int)threadIdx.x, w = t & ~31, lane = t & 31, row = lane >> 2, cq = 2 * (lane & 3);
    g_ma[t] = a;
    g_mb[t] = b;
    pthread_barrier_wait(&g_wbar[t >> 5]);
    for (int k = 0; k < 4; ++k) {
        c0 = std::fma(g_ma[w + 4 * row + k], g_mb[w + 4 * cq + k], c0);
        c1 = std::fma(g_ma[w + 4 * row + k], g_mb[w + 4 * (cq + 1) + k], c1);
    }
    pthread_barrier_wait(&g_wbar[t >> 5]);
}

#include "nbk_program_mma.cu"

#define EMUL_NAME NBK_CAT(emul_mma_, NBK_TAG)

namespace {
struct emul_arg { int kernel, tid; const nbk_kargs *a; };
void *emul_worker(void *p) {
    const emul_arg *g = (const emul_arg *)p;
    threadIdx.x = (unsigned)g->tid;
    blockDim.x = 256;
    blockIdx.x = 0;
    gridDim.x = 1;
    if (g->kernel == 0) NBK_KERNEL(init)(*g->a);
    else if (g->kernel == 1) NBK_KERNEL(run)(*g->a);
    else NBK_KERNEL(step)(*g->a);
    return nullptr;
}
}  // namespace

extern "C" int EMUL_NAME(int kernel, const nbk_kargs *a) {
    if (kernel < 0 || kernel > 2) return -1;
    pthread_barrier_init(&g_bar, nullptr, 256);
    g_or[0] = g_or[1] = 0;
    g_gthreads = nbk::MMA_GTHREADS;
    for (int g = 0; g < 2; ++g) {
        pthread_barrier_init(&g_gbar[g], nullptr, (unsigned)nbk::MMA_GTHREADS);
        g_gor[g][0] = g_gor[g][1] = 0;
    }
    for (int w = 0; w < 8; ++w) pthread_barrier_init(&g_wbar[w], nullptr, 32);
    pthread_t th[256];
    emul_arg args[256];
    for (int t = 0; t < 256; ++t) {
        args[t] = emul_arg{kernel, t, a};
        pthread_create(&th[t], nullptr, emul_worker, &args[t]);
    }
    for (int t = 0; t < 256; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&g_bar);
    for (int w = 0; w < 8; ++w) pthread_barrier_destroy(&g_wbar[w]);
    return 0;
}
