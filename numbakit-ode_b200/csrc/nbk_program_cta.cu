// nbk_program_cta.cu -- init / run / step kernels of one (method, RHS) pair in the
// CTA-per-trajectory regime (nbk_cta.cuh).  Macros: NBK_METHOD (23 | 45 | 853 | 1..5 | 202.. fixed-step), NBK_RHS_T (a CTA
// right-hand side type), NBK_E (state components per thread), NBK_TAG.  The block size is chosen at
// launch: ceil(n / NBK_E) rounded up to a warp; dynamic shared memory = (2*T*E + 34) doubles.
// NBK_GROUP = G > 0: sub-warp regime (nbk_group.cuh) -- the run / step kernels are persistent 128-thread blocks in which
// G lanes integrate one trajectory (RungeKutta23 / RungeKutta45, n <= G * NBK_E); the construction kernel stays the CTA one.
// Classes without a dedicated sub-warp kernel (DOP853, Adams-Bashforth, the fixed-step Runge-Kutta classes) run the CTA
// regime's source with the team = G lanes of a warp (NBK_TEAM_G, see nbk_cta.cuh): all three kernels then use
// NBK_GRP_BLOCK-thread blocks.
#ifndef NBK_GROUP
#define NBK_GROUP 0
#endif
#if NBK_GROUP > 0 && NBK_METHOD != 23 && NBK_METHOD != 45 && !defined(NBK_TEAM_G)
#define NBK_TEAM_G NBK_GROUP
#endif
#if NBK_GROUP > 0
#include "nbk_group.cuh"
#else
#include "nbk_cta.cuh"
#endif

#define NBK_CAT2(a, b) a##b
#define NBK_CAT(a, b) NBK_CAT2(a, b)
#define NBK_KERNEL(stem) NBK_CAT(NBK_CAT(nbk_, stem), NBK_CAT(_, NBK_TAG))

typedef NBK_RHS_T nbk_rhs_t;
#if NBK_METHOD == 45
typedef nbk::tab_rk45 nbk_tab_t;
#define NBK_CTA_ADAPTIVE 1
#elif NBK_METHOD == 23
typedef nbk::tab_rk23 nbk_tab_t;
#define NBK_CTA_ADAPTIVE 1
#elif NBK_METHOD == 853
typedef nbk::tab_dop853 nbk_tab_t;  // 13 stage vectors per thread: one or two components per thread (n <= 1024)
#define NBK_CTA_ADAPTIVE 1
#elif (NBK_METHOD >= 1 && NBK_METHOD <= 5) || NBK_METHOD == 202 || NBK_METHOD == 203 || NBK_METHOD == 213 || NBK_METHOD == 204 || NBK_METHOD == 238
#define NBK_CTA_ADAPTIVE 0  // fixed step: Adams-Bashforth-k / ForwardEuler and the fixed-step Runge-Kutta classes
#else
#error "the CTA regime implements RungeKutta23/45, DOP853, AdamsBashforth1..5 and the fixed-step Runge-Kutta classes"
#endif
#ifndef NBK_E
#define NBK_E 4
#endif
#ifndef NBK_CTA_MAXT
#define NBK_CTA_MAXT 512  // largest block the kernels are compiled for: n <= NBK_E * NBK_CTA_MAXT
#endif
#ifndef NBK_CTA_MINB
#define NBK_CTA_MINB 1    // resident CTAs per SM the advance kernels are compiled for (register budget)
#endif

// NBK_CTA_FULL: 1 = every thread owns NBK_E valid components (n == blockDim.x * NBK_E) and the adaptive kernels are
// compiled without bounds checks; 0 = any n.  Separate programs, not a branch inside one kernel: with both bodies in
// one function ptxas spills in the hot loop of the full body (C5: 568 -> 670 ms).
#ifndef NBK_CTA_FULL
#define NBK_CTA_FULL 0
#endif
#if NBK_GROUP > 0 && NBK_TEAM_G == 0
#ifndef NBK_GRP_MINB
#define NBK_GRP_MINB 4
#endif
#define NBK_ADV_BOUNDS __launch_bounds__(NBK_GRP_BLOCK, NBK_GRP_MINB)
#define NBK_INIT_BOUNDS __launch_bounds__(NBK_CTA_MAXT)
template <int MODE>
__device__ __forceinline__ void nbk_cta_advance_any(const nbk_kargs &a, double *sm) {
    nbk::group_advance<nbk_tab_t, nbk_rhs_t, NBK_GROUP, NBK_E, MODE, false, NBK_CTA_FULL != 0>(a, sm);
}
#elif NBK_TEAM_G > 0
// the CTA regime's source with sub-warp teams: resident blocks by register budget (DOP853 holds 15 vectors per lane)
#ifndef NBK_GRP_MINB
#define NBK_GRP_MINB (NBK_METHOD == 853 ? 3 : 4)
#endif
#define NBK_ADV_BOUNDS __launch_bounds__(NBK_GRP_BLOCK, NBK_GRP_MINB)
#define NBK_INIT_BOUNDS __launch_bounds__(NBK_GRP_BLOCK)
#if NBK_CTA_ADAPTIVE
template <int MODE>
__device__ __forceinline__ void nbk_cta_advance_any(const nbk_kargs &a, double *sm) {
    nbk::cta_advance<nbk_tab_t, nbk_rhs_t, NBK_E, MODE, false>(a, sm);
}
#endif
#else
#define NBK_INIT_BOUNDS __launch_bounds__(NBK_CTA_MAXT)
#define NBK_ADV_BOUNDS __launch_bounds__(NBK_CTA_MAXT, NBK_CTA_MINB)
#if NBK_CTA_ADAPTIVE
template <int MODE>
__device__ __forceinline__ void nbk_cta_advance_any(const nbk_kargs &a, double *sm) {
    nbk::cta_advance<nbk_tab_t, nbk_rhs_t, NBK_E, MODE, NBK_CTA_FULL != 0>(a, sm);
}
#endif
#endif

extern "C" __global__ void NBK_INIT_BOUNDS NBK_KERNEL(init)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ __align__(16) double nbk_smem[];
#if NBK_CTA_ADAPTIVE
    nbk::cta_init_adaptive<nbk_tab_t, nbk_rhs_t, NBK_E>(a, nbk_smem);
#else
    nbk::cta_init_fixed<NBK_METHOD, nbk_rhs_t, NBK_E>(a, nbk_smem);
#endif
}
extern "C" __global__ void NBK_ADV_BOUNDS NBK_KERNEL(run)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ __align__(16) double nbk_smem[];
#if NBK_CTA_ADAPTIVE
    nbk_cta_advance_any<NBK_MODE_RUN>(a, nbk_smem);
#else
    nbk::cta_advance_fixed<NBK_METHOD, nbk_rhs_t, NBK_E, NBK_MODE_RUN>(a, nbk_smem);
#endif
}
extern "C" __global__ void NBK_ADV_BOUNDS NBK_KERNEL(step)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ __align__(16) double nbk_smem[];
#if NBK_CTA_ADAPTIVE
    nbk_cta_advance_any<NBK_MODE_STEP>(a, nbk_smem);
#else
    nbk::cta_advance_fixed<NBK_METHOD, nbk_rhs_t, NBK_E, NBK_MODE_STEP>(a, nbk_smem);
#endif
}
#if defined(NBK_NEVENTS) && !NBK_CTA_ADAPTIVE
// run_events of the fixed-step classes (sub-warp teams and one CTA per trajectory): the run kernel with the team's event
// routine after every step; dynamic shared memory as below (one more whole-vector buffer per team)
extern "C" __global__ void NBK_ADV_BOUNDS NBK_KERNEL(runev)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ __align__(16) double nbk_smem[];
    nbk::cta_advance_fixed<NBK_METHOD, nbk_rhs_t, NBK_E, NBK_MODE_RUN, true>(a, nbk_smem);
}
#endif
#if defined(NBK_NEVENTS) && NBK_CTA_ADAPTIVE
// run_events variant of the run kernel (nbk_set_events).  Dynamic shared memory: the run kernel's + one whole-vector buffer
// for the event functions -- sub-warp regime 3 * NBK_GRP_BLOCK * NBK_E doubles, one CTA per trajectory 3 * T * NBK_E + 34.
extern "C" __global__ void NBK_ADV_BOUNDS NBK_KERNEL(runev)(const __grid_constant__ nbk_kargs a) {
    extern __shared__ __align__(16) double nbk_smem[];
#if NBK_GROUP > 0 && NBK_TEAM_G == 0
    nbk::group_advance<nbk_tab_t, nbk_rhs_t, NBK_GROUP, NBK_E, NBK_MODE_RUN, true, NBK_CTA_FULL != 0>(a, nbk_smem);
#elif NBK_TEAM_G > 0
    nbk::cta_advance<nbk_tab_t, nbk_rhs_t, NBK_E, NBK_MODE_RUN, false, true>(a, nbk_smem);
#else
    nbk::cta_advance<nbk_tab_t, nbk_rhs_t, NBK_E, NBK_MODE_RUN, NBK_CTA_FULL != 0, true>(a, nbk_smem);
#endif
}
#endif

#ifndef __CUDACC_RTC__
#include "nbk_host.h"
extern "C" void NBK_KERNEL(program)(nbk_static_program *out) {
    out->method = NBK_METHOD;
    out->n = -1;  // any n: run-time
    out->n_params = nbk_rhs_t::NP;
    out->cta_elems = NBK_E;
    out->cta_full = NBK_CTA_FULL;
    out->group = NBK_GROUP;
    out->mma_tile = 0;
    out->mma_block = 0;
    out->mma_smem = 0;
    out->k_init = (const void *)&NBK_KERNEL(init);
    out->k_run = (const void *)&NBK_KERNEL(run);
    out->k_step = (const void *)&NBK_KERNEL(step);
}
#endif
