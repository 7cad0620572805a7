// nbk_emul_group.cpp -- TEST INFRASTRUCTURE: host build of one program of the sub-warp regime (csrc/nbk_group.cuh through
// csrc/nbk_program_cta.cu with NBK_GROUP = G).  The construction kernel is the CTA regime's: a block of T POSIX threads as in
// nbk_emul_cta.cpp.  The run / step kernels are emulated as ONE warp of 32 POSIX threads: group_sync is a barrier of the
// group's G threads, the group shuffles exchange through a scratch array between two such barriers, warp_all is a
// reduction between two barriers of all 32 threads, and atomicAdd is a real atomic (the groups of the warp pull from the
// trajectory queue concurrently).  The OS schedules the 32 threads arbitrarily, so every hazard a missing group barrier
// would open shows up as run-to-run differences (tests/test_group_logic_host.py repeats runs and demands bit equality).
// Nothing in the product links this.
#include <pthread.h>

#include "cuda_host_shim.h"

static thread_local emul_dim3 gridDim;
struct double2 { double x, y; };
#define __shared__
#define __align__(x)

static pthread_barrier_t g_bar;            // all threads of the emulated block
static pthread_barrier_t g_gbar[32];       // one per group of the warp (indexed by group number)
static double g_shfl[1024];
static long long g_shfl_ll[32];
static int g_flag[32];
static int g_group = 32;                   // lanes per group of the run / step kernels
alignas(16) double nbk_smem[1 << 16];

static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline double __shfl_xor_sync(unsigned, double v, int o) {  // construction kernel: block_sum of nbk_cta.cuh
    const int t = (int)threadIdx.x;
    g_shfl[t] = v;
    pthread_barrier_wait(&g_bar);
    const double r = g_shfl[t ^ o];
    pthread_barrier_wait(&g_bar);
    return r;
}
static inline unsigned long long emul_atomic_add(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
#define atomicAdd emul_atomic_add

namespace nbk {
static inline void group_sync(unsigned) { pthread_barrier_wait(&g_gbar[threadIdx.x / g_group]); }
static inline double group_shfl_xor(unsigned, double v, int o) {
    const int t = (int)threadIdx.x;
    g_shfl[t] = v;
    pthread_barrier_wait(&g_gbar[t / g_group]);
    const double r = g_shfl[t ^ o];
    pthread_barrier_wait(&g_gbar[t / g_group]);
    return r;
}
static inline long long group_bcast(unsigned, long long v, int src_lane) {
    const int t = (int)threadIdx.x;
    g_shfl_ll[t] = v;
    pthread_barrier_wait(&g_gbar[t / g_group]);
    const long long r = g_shfl_ll[src_lane];
    pthread_barrier_wait(&g_gbar[t / g_group]);
    return r;
}
static inline bool warp_all(bool p) {
    g_flag[threadIdx.x] = p ? 1 : 0;
    pthread_barrier_wait(&g_bar);
    bool all = true;
    for (int i = 0; i < 32; ++i) all = all && g_flag[i] != 0;
    pthread_barrier_wait(&g_bar);
    return all;
}
}  // namespace nbk

#ifdef EMUL_USER_PRELUDE  // a user right-hand side (nbk_rhs_i), spliced in the way csrc/nbk_jit.cpp does for NVRTC
#include EMUL_USER_PRELUDE
#endif
#include "nbk_program_cta.cu"

#define EMUL_NAME NBK_CAT(emul_group_, NBK_TAG)

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

// kernel 0: construction with a block of T threads; 1 / 2 / 3: run / step / run_events with one warp (T is ignored)
extern "C" int EMUL_NAME(int kernel, const nbk_kargs *a, int T) {
    if (kernel < 0 || kernel > 3) return -1;
    if (kernel != 0 || NBK_TEAM_G > 0) T = 32;  // (with sub-warp teams in every kernel the construction runs in the warp too)
    if (T < 32 || T > 1024 || T % 32) return -1;
    g_group = NBK_GROUP;
    pthread_barrier_init(&g_bar, nullptr, (unsigned)T);
    for (int g = 0; g < 32 / NBK_GROUP; ++g) pthread_barrier_init(&g_gbar[g], nullptr, (unsigned)NBK_GROUP);
    pthread_t th[1024];
    emul_arg args[1024];
    for (int t = 0; t < T; ++t) {
        args[t] = emul_arg{kernel, t, T, a};
        pthread_create(&th[t], nullptr, emul_worker, &args[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&g_bar);
    for (int g = 0; g < 32 / NBK_GROUP; ++g) pthread_barrier_destroy(&g_gbar[g]);
    return 0;
}
