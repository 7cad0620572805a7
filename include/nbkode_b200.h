/*
 * nbkode_b200.h -- C-ABI of libnbkode_b200.so, the B200-native batched ODE-ensemble engine
 * behind the nbkode solver API (hgrecco/numbakit-ode).
 *
 * The reference has no FFI: its boundary is the Python class API.
 * Each entry point below replaces the reference call chain named in its comment (paths are
 * relative to the reference repository); the Python shell nbkode_b200/ binds them with ctypes
 * exactly as INTEGRATION.md shows.
 *
 * Conventions
 *  - plain C types only; every pointer marked "device" is a CUDA device pointer owned by the
 *    caller (the Python shell passes torch.Tensor.data_ptr()); the library owns only its handle,
 *    compiled programs and a small work-queue counter;
 *  - every function returns 0 on success, a negative NBK_E* code otherwise; the message is
 *    available from nbk_last_error() (thread-local);
 *  - kernels are launched on the caller's stream (cudaStream_t passed as void*) and never
 *    synchronise; a handle is not re-entrant, different handles may be used from different
 *    host threads / devices;
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    NBK_ECUDA.
 *
 * Ensemble state layout (struct nbk_state; B = batch, n = state size, trajectory index
 * fastest so that consecutive lanes touch consecutive addresses):
 *    ts [L][B]  ys [L][n][B]  fs [L][n][B]   history window, oldest -> newest
 *                                            (nbkode/buffer.py:5-56, AlignedBuffer)
 *    K  [S+1][n][B]                          stage derivatives of the last accepted step
 *                                            (nbkode/runge_kutta/core.py:25-27), RK only
 *    h  [B]                                  next step size (nbkode/core.py:706)
 *    params [np][B] or [np]                  nbkode/core.py:118-119
 *    n_accepted, n_rejected [B], status [B], n_out [B], work_counter [2]   (no reference counterpart)
 *  L and S+1 come from nbk_layout().
 */
#ifndef NBKODE_B200_H
#define NBKODE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define NBK_VERSION 1

/* error codes */
#define NBK_OK 0
#define NBK_EINVAL (-1)   /* bad argument (maps to ValueError) */
#define NBK_ECUDA (-2)    /* CUDA runtime / driver failure, or no device */
#define NBK_ECOMPILE (-3) /* NVRTC failed; nbk_last_error() carries the compiler log */
#define NBK_ENOTFOUND (-4)/* unknown built-in right-hand side */

/* methods: nbkode class names -> ids */
#define NBK_RK23 23 /* RungeKutta23   nbkode/runge_kutta/explicit.py:132-164 */
#define NBK_RK45 45 /* RungeKutta45   nbkode/runge_kutta/explicit.py:167-253 */
#define NBK_DOP853 853 /* DOP853      nbkode/runge_kutta/explicit.py:282-351 (per-thread regime) */
#define NBK_AB1 1   /* AdamsBashforth1 / ForwardEuler  multistep/adams_bashforth.py:17-20, euler.py:19-27 */
#define NBK_AB2 2   /* AdamsBashforth2..5              multistep/adams_bashforth.py:23-61 */
#define NBK_AB3 3
#define NBK_AB4 4
#define NBK_AB5 5
/* implicit linear multistep (per-thread regime, n <= 8, kernels compiled by NVRTC):
   AdamsMoulton1..5 = 11..15 (BackwardEuler == AdamsMoulton1), nbkode/multistep/adams_moulton.py:31-61, euler.py:30-36;
   BDF1..6 = 31..36, nbkode/multistep/bdf.py:16-44; implicit step nbkode/multistep/core.py:115-170 */
#define NBK_AM1 11
#define NBK_AM5 15
#define NBK_BDF1 31
#define NBK_BDF6 36
/* fixed-step explicit Runge-Kutta classes, nbkode/runge_kutta/explicit.py:102-129 */
#define NBK_RUNGE2 202
#define NBK_RUNGE3 203
#define NBK_HEUN3 213
#define NBK_RK4 204
#define NBK_RK38 238

/* per-trajectory status */
#define NBK_TRAJ_OK 0
#define NBK_TRAJ_BOUND 1       /* stopped by the t_bound / upto_t guard, nbkode/core.py:176-184 */
#define NBK_TRAJ_NONFINITE 2   /* non-finite error norm or collapsed step (the reference spins forever) */
#define NBK_TRAJ_MAXATTEMPTS 3 /* attempt budget of a call exhausted; terminal (sticky) like every status >= 2 */
#define NBK_TRAJ_EVBRACKET 4   /* nbk_run_events: an event function does not change sign on the step's interpolant
                                  (the reference's bisect raises ValueError, nbkode/nbcompat/zeros.py:516-520) */
#define NBK_TRAJ_NEWTON 5      /* implicit multistep: Newton / secant iteration did not converge (the reference raises
                                  RuntimeError, nbkode/nbcompat/zeros.py:432-434, :503-505) */
#define NBK_TRAJ_EVENT (-1)    /* nbk_run_events: stopped by a terminal event (clean stop) */
#define NBK_MAX_EVENTS 8

typedef struct nbk_solver nbk_solver;

typedef struct {
    int method;           /* NBK_RK23, NBK_RK45, NBK_AB1..5, NBK_RUNGE2 ... NBK_RK38 */
    int n;                /* state size */
    int n_params;         /* parameters per trajectory (0 allowed) */
    int params_shared;    /* 1: one parameter vector for the whole ensemble */
    long long batch;      /* B */
    const char *rhs;      /* built-in right-hand side: "lorenz", "vdp", "decay" (per-thread regime),
                             "rd1d", "linear" (CTA regime); or NULL */
    const char *rhs_src;  /* CUDA C source compiled by NVRTC into the integrator kernels; or NULL.  Defines
                               NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy)
                             (per-thread regime, y/dy in registers) or, for large systems,
                               NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n)
                             (CTA regime: returns component i; y[j] reads component j of the whole vector, which lives in
                             shared memory).  Shape of the CTA: a right-hand side whose source contains a loop (cost per
                             component growing with n, e.g. a matrix-vector product) runs with one component per thread
                             as long as the block allows; loop-free (stencil-like) source with four components per thread
                             from n = 128 on (measured crossover).
                             Replaces numba.njit(rhs) of nbkode/core.py:125-132. */
    /* VariableStepOptions, nbkode/core.py:661-679 (min_step/max_step are never enforced by
       the reference and therefore absent) */
    double rtol, atol, min_factor, max_factor, safety_factor;
    int first_stepper;    /* multistep start-up (nbkode/multistep/core.py:37-65):
                             0 none, NBK_RK45 ("auto"), NBK_RK23, NBK_AB1..5 */
    int device;           /* CUDA device ordinal */
} nbk_config;

typedef struct {          /* all device pointers, see the layout above */
    double *ts, *ys, *fs, *K, *h, *params;
    long long *n_accepted, *n_rejected;
    int *status, *n_out;
    unsigned long long *work_counter; /* 16 bytes, reset by every call: [0] head of the trajectory queue of the
                                         persistent adaptive kernels, [1] number of trajectories that ended the
                                         call with a failure status (>= NBK_TRAJ_NONFINITE): one 8-byte copy tells
                                         the host whether anything failed */
} nbk_state;

int nbk_version(void);
const char *nbk_last_error(void);

/* Replaces class construction + numba JIT of the step closures
   (nbkode/core.py:143-184 __init_subclass__, :125-132 rhs jitting). Built-in right-hand sides
   use kernels compiled ahead of time by nvcc; rhs_src goes through NVRTC (sm_100a). */
int nbk_create(const nbk_config *cfg, nbk_solver **out);
void nbk_destroy(nbk_solver *s);

/* History length L and number of stage rows S+1 (0 for multistep) the caller must allocate. */
int nbk_layout(const nbk_solver *s, int *len_history, int *stage_rows);
/* 1 if the kernels came from NVRTC, 0 if from the ahead-of-time catalogue. */
int nbk_is_jit(const nbk_solver *s);
/* Execution regime, which also fixes the layout of nbk_state:
     0  one trajectory per thread (n <= 64):   ts [L][B], ys/fs [L][n][B], K [S+1][n][B], params [np][B]
     1  one CTA per trajectory (large n; chosen by the right-hand side: built-ins "rd1d", "linear" or
        user source defining nbk_rhs_i): ts [B][2], ys/fs [B][2][n], K [B][S+1][n], params [B][np]
        (a trajectory's n values are contiguous).  h, counters and status are [B] in both. */
int nbk_regime(const nbk_solver *s);
/* Regime 1 only: lanes of a warp that integrate one trajectory in the advance kernels (sub-warp regime for small and
   mid-size component-wise systems, n <= 127: 2, 4, 8, 16 or 32 lanes -- dedicated kernels for RungeKutta23 / RungeKutta45,
   the CTA regime's source with sub-warp teams for DOP853 (n <= 64), Adams-Bashforth and the fixed-step Runge-Kutta classes;
   the reference has one code path for every n, nbkode/runge_kutta/explicit.py:60-79); 0 = the whole CTA (or regime 0). */
int nbk_group_lanes(const nbk_solver *s);

/* Replaces Solver.__init__ / VariableStep.__init__ / Multistep.__init__
   (nbkode/core.py:106-141, :687-706, nbkode/multistep/core.py:37-65): f0 = rhs(t0, y0), history
   filled with (t0, y0, f0), h = select_initial_step(...) when h is NaN (adaptive), multistep
   start-up.  y0 element (b, c) is read from y0[b*y0_sb + c*y0_sc] (device); params likewise
   (ignored when n_params == 0; for params_shared the np values are read from params[c*p_sc]).
   h_per_traj (device, [B]) overrides h when not NULL. */
int nbk_init(nbk_solver *s, const nbk_state *st, double t0, const double *y0, long long y0_sb,
             long long y0_sc, const double *params, long long p_sb, long long p_sc, double h,
             const double *h_per_traj, void *stream);

/* Replaces Solver.run -> _run_eval (nbkode/core.py:315-332, :598-619) including the past-point
   branch of run_events (:401-408): for each sorted grid time ts[k] (device, [m]) advance every
   trajectory while t < ts[k] (steps are never clipped) and write the dense output
   (RkInterpolator.evaluate explicit.py:371-403 / Solver._interpolate core.py:577-596) to
   y_out[b*o_sb + k*o_sk + c*o_sc] (device).  A step starts only if t + h <= t_bound;
   trajectories stopped by that guard report NBK_TRAJ_BOUND and n_out < m. */
int nbk_run(nbk_solver *s, const nbk_state *st, const double *ts, int m, double t_bound,
            double *y_out, long long o_sb, long long o_sk, long long o_sc, long long max_attempts,
            void *stream);

/* Replaces Solver.step / Solver.skip -> _nsteps/_steps/_nskip/_skip (nbkode/core.py:214-313,
   :483-575): up to n_steps accepted steps per trajectory, each starting only if
   t + h <= bound (bound = upto_t or t_bound).  With emit != 0 step j of trajectory b is written
   to t_out[b*to_sb + j] and y_out[b*o_sb + j*o_sk + c*o_sc]; st->n_out[b] receives the number
   of steps taken. */
int nbk_step(nbk_solver *s, const nbk_state *st, long long n_steps, double bound, int emit,
             double *t_out, long long to_sb, double *y_out, long long o_sb, long long o_sk,
             long long o_sc, long long max_attempts, void *stream);

/* Events, replacing event_handler.build_handler + Solver._run_eval_events (nbkode/event_handler.py,
   nbkode/core.py:334-447, :621-647).  nbk_set_events compiles (NVRTC) a variant of the run kernel with the
   n_events event functions inlined; event_src is CUDA C defining
       NBK_EVENT nbk_event(int k, double t, const double* y, const double* p)     (returns g_k(t, y))
   Every regime takes the same source: per-thread, and -- for RungeKutta23 / RungeKutta45 / DOP853 -- the sub-warp groups and
   one CTA per trajectory, where the team's threads publish the whole state vector in shared memory for the event
   functions.  nbk_run_events is nbk_run plus, after every accepted step, sign-change detection per
   event (direction[k] in {-1, 0, +1}), bisection of the root on the step's dense output (|g| <= 1.48e-8, <= 50
   halvings: nbkode/nbcompat/zeros.py:508-533) and recording of (t, y(t)); a terminal event (terminal[k] != 0) stops
   its trajectory: output row n_out-1 holds the event state, t_terminal[b] its time, status NBK_TRAJ_EVENT. */
typedef struct {
    int n_events;          /* must equal the value given to nbk_set_events */
    int capacity;          /* events stored per (trajectory, event function) */
    const int *direction;  /* host, [n_events] */
    const int *terminal;   /* host, [n_events] */
    double *t_events;      /* device, [B][n_events][capacity] */
    double *y_events;      /* device, [B][n_events][capacity][n] */
    int *counts;           /* device, [B][n_events]: events detected (> capacity: the excess was not stored) */
    double *t_terminal;    /* device, [B] */
} nbk_events;
int nbk_set_events(nbk_solver *s, int n_events, const char *event_src);
int nbk_run_events(nbk_solver *s, const nbk_state *st, const double *ts, int m, double t_bound, double *y_out,
                   long long o_sb, long long o_sk, long long o_sc, const nbk_events *ev, long long max_attempts,
                   void *stream);

/* Launch geometry of the most recent nbk_run / nbk_step (for reporting). */
int nbk_last_launch(const nbk_solver *s, int *grid, int *block, int *regs_per_thread);

/* NVRTC front half only (no device needed): compiles the kernels of cfg for sm_100a and returns
   the cubin size; used by the CPU test-suite and to pre-warm the program cache. */
int nbk_jit_compile_only(const nbk_config *cfg, long long *cubin_bytes);
/* The same for the event variant of the program (see nbk_set_events). */
int nbk_jit_compile_events_only(const nbk_config *cfg, int n_events, const char *event_src, long long *cubin_bytes);

/* FP64 vector-pipe roofline denominator: runs a register-resident DFMA chain kernel for
   `iters` iterations and returns the achieved TFLOP/s (2 FLOP per DFMA) and its duration. */
int nbk_fp64_peak(int device, int iters, double *tflops, double *millis);
/* The same for the FP64 tensor pipe (mma.sync m8n8k4 f64): denominator for shared-matrix linear systems. */
int nbk_fp64_mma_peak(int device, int iters, double *tflops, double *millis);

#ifdef __cplusplus
}
#endif
#endif
