// nbk_kargs.h -- kernel argument block shared by the host launcher (nbk_api.cu) and the
// device programs (nbk_device.cuh, compiled by nvcc for the built-in catalogue and by NVRTC
// for user right-hand sides).  Plain C types only: NVRTC compiles this without host headers.
#ifndef NBK_KARGS_H
#define NBK_KARGS_H

// status codes per trajectory (also exported through include/nbkode_b200.h)
#define NBK_TRAJ_OK 0          // reached the end of the request
#define NBK_TRAJ_BOUND 1       // stopped by the t_bound / upto_t guard (nbkode/core.py:176-184)
#define NBK_TRAJ_NONFINITE 2   // non-finite error norm or step size collapsed to 0
#define NBK_TRAJ_MAXATTEMPTS 3 // attempt budget of a call exhausted -- terminal like the other failures (status >= 2 is sticky:
                               // the stage rows of the interrupted step are overwritten, so the trajectory is not resumed)
#define NBK_TRAJ_EVBRACKET 4   // run_events: the event function has the same sign at both ends of the step on the
                               // interpolant (the reference's bisect raises ValueError, nbcompat/zeros.py:516-520)
#define NBK_TRAJ_NEWTON 5      // implicit step: Newton / secant iteration did not converge (the reference raises RuntimeError)
#define NBK_TRAJ_EVENT (-1)    // run_events: stopped by a terminal event (clean stop, not sticky)
#define NBK_MAX_EVENTS 8

// advance modes
#define NBK_MODE_RUN 0   // run(ts): emit dense output at the shared grid tgrid[0..m)
#define NBK_MODE_STEP 1  // step(n, upto_t) / skip: emit (t, y) after every accepted step

struct nbk_kargs {
    // ---- ensemble state, SoA with the trajectory index fastest (see DESIGN.md "HBM layout")
    long long B;          // trajectories
    double *ts;           // [L][B]      history times, oldest -> newest
    double *ys;           // [L][n][B]   history states
    double *fs;           // [L][n][B]   history derivatives
    double *K;            // [S+1][n][B] stage derivatives of the last accepted step (RK only)
    double *h;            // [B]         next step size
    double *params;       // [np][B] (or [np] when params_shared)
    long long *n_acc;     // [B] accepted steps so far
    long long *n_rej;     // [B] rejected attempts so far
    int *status;          // [B] NBK_TRAJ_*
    int params_shared;
    int first_stepper;    // multistep start-up: 0 none, 23/45 adaptive RK, 1..5 Adams-Bashforth-k
    int pristine;         // the state is exactly what the init kernel wrote (no step taken yet):
                          // the advance kernel then skips loading the K rows / old history slot
    int n, n_params;      // run-time sizes (CTA-per-trajectory regime; the per-thread regime has them
                          // as template parameters)
    // ---- init inputs (user layout: element (b, c) at base[b*sb + c*sc])
    const double *y0;
    long long y0_sb, y0_sc;
    const double *p0;
    long long p0_sb, p0_sc;
    const double *h0;     // optional per-trajectory first step / fixed step ([B]); NULL -> h_init
    double t0;
    double h_init;        // NaN -> select_initial_step (adaptive); fixed step for multistep
    // ---- controller options (nbkode/core.py:661-679)
    double rtol, atol, min_factor, max_factor, safety;
    double x_lo, x_hi;    // clamp of the squared error norm outside which the factor clip decides
    // ---- advance request
    double bound;         // t_bound or upto_t: a step starts only if t + h <= bound
    const double *tgrid;  // [m] sorted output times shared by all trajectories (RUN)
    int m;
    int emit;             // STEP: 0 = skip (no outputs)
    long long n_steps;    // STEP: accepted steps to take (per trajectory)
    double *y_out;        // RUN: (b,k,c) at y_out[b*o_sb + k*o_sk + c*o_sc]; STEP: k = step index
    long long o_sb, o_sk, o_sc;
    double *t_out;        // STEP: (b,k) at t_out[b*to_sb + k]
    long long to_sb;
    int *n_out;           // [B] grid points emitted (RUN) / steps taken (STEP) in this call
    long long max_attempts;
    unsigned long long *work_counter;  // [2]: [0] persistent-kernel trajectory queue head, [1] trajectories that
                                       // ended this call with a failure status (>= NBK_TRAJ_NONFINITE)
    // ---- run_events (nbkode/core.py:334-447, event_handler.py); read only by programs compiled with NBK_NEVENTS
    int ev_cap;           // capacity per (trajectory, event) of ev_t / ev_y
    int ev_terminal;      // bit k set: event k is terminal
    int ev_dir[8];        // direction of event k: -1, 0, +1 (sign of event.direction)
    double *ev_t;         // [B][NE][cap]    event times in order of detection
    double *ev_y;         // [B][NE][cap][n] interpolated states at the event times
    int *ev_count;        // [B][NE] events detected in this call (only the first ev_cap are stored)
    double *ev_tterm;     // [B] time of the terminal event that stopped the trajectory: replaces the time of
                          //     output row n_out-1 (core.py:641-644); untouched otherwise
};

#endif
