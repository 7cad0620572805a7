// nbk_program.cu -- one "program" = the init / run / step kernels of one (method, RHS) pair.
//
// Compiled once per built-in pair by nvcc (see Makefile: -DNBK_METHOD=.. -DNBK_RHS_T=..
// -DNBK_TAG=..) and at run time by NVRTC for user right-hand sides (nbk_jit.cpp passes the
// same macros plus -DNBK_USER_RHS and prepends the user's nbk_rhs definition).
//
// NBK_METHOD: 23 / 45 / 853 = RungeKutta23 / RungeKutta45 / DOP853; 1..5 = AdamsBashforth-k (1 == ForwardEuler);
// 202 / 203 / 213 / 204 / 238 = Runge2 / Runge3 / Heun3 / RungeKutta4 / RungeKutta3_8 (fixed step);
// 11..15 = AdamsMoulton1..5 (BackwardEuler == 11), 31..36 = BDF1..6 (implicit multistep).
#include "nbk_device.cuh"

#define NBK_CAT2(a, b) a##b
#define NBK_CAT(a, b) NBK_CAT2(a, b)
#define NBK_KERNEL(stem) NBK_CAT(NBK_CAT(nbk_, stem), NBK_CAT(_, NBK_TAG))

typedef NBK_RHS_T nbk_rhs_t;

#if NBK_METHOD == 45
typedef nbk::tab_rk45 nbk_tab_t;
#define NBK_ADAPTIVE 1
#elif NBK_METHOD == 23
typedef nbk::tab_rk23 nbk_tab_t;
#define NBK_ADAPTIVE 1
#elif NBK_METHOD == 853
typedef nbk::tab_dop853 nbk_tab_t;
#define NBK_ADAPTIVE 1
#elif NBK_METHOD >= 1 && NBK_METHOD <= 5
#define NBK_ADAPTIVE 0
#elif (NBK_METHOD >= 11 && NBK_METHOD <= 15) || (NBK_METHOD >= 31 && NBK_METHOD <= 36)
#define NBK_ADAPTIVE 0
#define NBK_IMPLICIT 1
#elif NBK_METHOD == 202 || NBK_METHOD == 203 || NBK_METHOD == 213 || NBK_METHOD == 204 || NBK_METHOD == 238
#define NBK_ADAPTIVE 0
#define NBK_FIXED_RK 1
#else
#error "NBK_METHOD must be 23, 45, 1..5 or a fixed-step Runge-Kutta id (202, 203, 213, 204, 238)"
#endif
#ifndef NBK_FIXED_RK
#define NBK_FIXED_RK 0
#endif
#ifndef NBK_IMPLICIT
#define NBK_IMPLICIT 0
#endif

#ifndef NBK_MIN_BLOCKS
#define NBK_MIN_BLOCKS 1
#endif

extern "C" __global__ void __launch_bounds__(NBK_BLOCK) NBK_KERNEL(init)(const __grid_constant__ nbk_kargs a) {
#if NBK_ADAPTIVE
    nbk::init_adaptive<nbk_tab_t, nbk_rhs_t>(a);
#elif NBK_FIXED_RK
    nbk::init_erk<NBK_METHOD, nbk_rhs_t>(a);
#elif NBK_IMPLICIT
    nbk::init_ab<nbk::tab_im<NBK_METHOD>::K, nbk_rhs_t>(a);  // same history window and start-up as Adams-Bashforth-K
#else
    nbk::init_ab<NBK_METHOD, nbk_rhs_t>(a);
#endif
}

extern "C" __global__ void __launch_bounds__(NBK_BLOCK, NBK_MIN_BLOCKS) NBK_KERNEL(run)(const __grid_constant__ nbk_kargs a) {
#if NBK_ADAPTIVE
    nbk::advance_adaptive<nbk_tab_t, nbk_rhs_t, NBK_MODE_RUN>(a);
#elif NBK_FIXED_RK
    nbk::advance_erk<NBK_METHOD, nbk_rhs_t, NBK_MODE_RUN>(a);
#elif NBK_IMPLICIT
    nbk::advance_im<NBK_METHOD, nbk_rhs_t, NBK_MODE_RUN>(a);
#else
    nbk::advance_ab<NBK_METHOD, nbk_rhs_t, NBK_MODE_RUN>(a);
#endif
}

extern "C" __global__ void __launch_bounds__(NBK_BLOCK, NBK_MIN_BLOCKS) NBK_KERNEL(step)(const __grid_constant__ nbk_kargs a) {
#if NBK_ADAPTIVE
    nbk::advance_adaptive<nbk_tab_t, nbk_rhs_t, NBK_MODE_STEP>(a);
#elif NBK_FIXED_RK
    nbk::advance_erk<NBK_METHOD, nbk_rhs_t, NBK_MODE_STEP>(a);
#elif NBK_IMPLICIT
    nbk::advance_im<NBK_METHOD, nbk_rhs_t, NBK_MODE_STEP>(a);
#else
    nbk::advance_ab<NBK_METHOD, nbk_rhs_t, NBK_MODE_STEP>(a);
#endif
}

#ifdef NBK_NEVENTS
// run_events: the run kernel with the event functions compiled in (NVRTC only)
// (no occupancy bound: the event state and the bisection need registers the plain run kernel does not)
extern "C" __global__ void __launch_bounds__(NBK_BLOCK) NBK_KERNEL(runev)(const __grid_constant__ nbk_kargs a) {
#if NBK_ADAPTIVE
    nbk::advance_adaptive<nbk_tab_t, nbk_rhs_t, NBK_MODE_RUN, true>(a);
#elif NBK_FIXED_RK
    nbk::advance_erk<NBK_METHOD, nbk_rhs_t, NBK_MODE_RUN, true>(a);
#elif NBK_IMPLICIT
    nbk::advance_im<NBK_METHOD, nbk_rhs_t, NBK_MODE_RUN, true>(a);
#else
    nbk::advance_ab<NBK_METHOD, nbk_rhs_t, NBK_MODE_RUN, true>(a);
#endif
}
#endif

#ifndef __CUDACC_RTC__
#include "nbk_host.h"
// registration record picked up by nbk_api.cu (static catalogue)
extern "C" void NBK_KERNEL(program)(nbk_static_program *out) {
    out->method = NBK_METHOD;
    out->n = nbk_rhs_t::N;
    out->n_params = nbk_rhs_t::NP;
    out->cta_elems = 0;
    out->cta_full = 0;
    out->group = 0;
    out->mma_tile = 0;
    out->mma_block = 0;
    out->mma_smem = 0;
    out->k_init = (const void *)&NBK_KERNEL(init);
    out->k_run = (const void *)&NBK_KERNEL(run);
    out->k_step = (const void *)&NBK_KERNEL(step);
}
#endif
