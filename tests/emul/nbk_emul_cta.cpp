// nbk_emul_cta.cpp -- TEST INFRASTRUCTURE: host build of one CTA-regime kernel program (csrc/nbk_program_cta.cu).
// A thread block is emulated by T POSIX threads: __syncthreads is a pthread barrier, a warp shuffle is an exchange
// through a scratch array between two barriers (every shuffle in nbk_cta.cuh is executed by all threads of the block),
// shared memory is one global array.  Checks the control flow and arithmetic of the same source the GPU runs
// (stage loop, block reductions, accept / reject, dense output, ring-buffered history, ragged and "full" shapes) in
// `pytest -m "not gpu"`; nothing in the product links this.
#include <pthread.h>

#include "cuda_host_shim.h"

static thread_local emul_dim3 gridDim;
struct double2 { double x, y; };
#define __shared__
#define __align__(x)

static pthread_barrier_t g_bar;
static double g_shfl[1024];
alignas(16) double nbk_smem[1 << 16];

static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline double __shfl_xor_sync(unsigned, double v, int o) {
    const int t = (int)threadIdx.x;
    g_shfl[t] = v;
    pthread_barrier_wait(&g_bar);
    const double r = g_shfl[t ^ o];
    pthread_barrier_wait(&g_bar);
    return r;
}

#ifdef EMUL_USER_PRELUDE  // a user right-hand side (nbk_rhs_i), spliced in the way csrc/nbk_jit.cpp does for NVRTC
#include EMUL_USER_PRELUDE
#endif
#include "nbk_program_cta.cu"

#define EMUL_NAME NBK_CAT(emul_cta_, NBK_TAG)

namespace {
struct emul_arg { int kernel, tid, T; const nbk_kargs *a; };
void *emul_worker(void *p) {
    const emul_arg *g = (const emul_arg *)p;
    threadIdx.x = (unsigned)g->tid;
    blockDim.x = (unsigned)g->T;
    blockIdx.x = 0;
    gridDim.x = 1;
    if (g->kernel == 0) NBK_KERNEL(init)(*g->a);
    else if (g->kernel == 1) NBK_KERNEL(run)(*g->a);
#ifdef NBK_NEVENTS
    else if (g->kernel == 3) NBK_KERNEL(runev)(*g->a);
#endif
    else NBK_KERNEL(step)(*g->a);
    return nullptr;
}
}  // namespace

extern "C" int EMUL_NAME(int kernel, const nbk_kargs *a, int T) {
    if (kernel < 0 || kernel > 3 || T < 32 || T > 1024 || T % 32) return -1;
    pthread_barrier_init(&g_bar, nullptr, (unsigned)T);
    pthread_t th[1024];
    emul_arg args[1024];
    for (int t = 0; t < T; ++t) {
        args[t] = emul_arg{kernel, t, T, a};
        pthread_create(&th[t], nullptr, emul_worker, &args[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&g_bar);
    return 0;
}
