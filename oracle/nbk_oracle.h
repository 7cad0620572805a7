/* nbk_oracle.h -- C-ABI of the CPU oracle (TEST INFRASTRUCTURE, see nbk_oracle.c). */
#ifndef NBK_ORACLE_H
#define NBK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RHS_DECAY = 0, ORC_RHS_LORENZ = 1, ORC_RHS_VDP = 2, ORC_RHS_LINEAR = 3, ORC_RHS_RD1D = 4 };

typedef struct {
    int method;        /* 23 / 45 = RungeKutta23 / RungeKutta45; 1..5 = AdamsBashforth-k (1 == ForwardEuler) */
    int rhs;           /* ORC_RHS_* */
    int n, np;         /* state size, parameter count */
    int params_shared; /* 1: params[np] shared by all trajectories; 0: params[B][np] */
    int first_stepper; /* AB start-up: 0 none, 45 / 23 adaptive (default tolerances), 1..5 AdamsBashforth-k */
    long max_rejects;  /* guard: give up (status 2) after this many rejections in one trajectory */
    double t0;
    double rtol, atol, min_factor, max_factor, safety;
    double first_step; /* NaN: select_initial_step */
    double h;          /* fixed step (Adams-Bashforth) */
    double t_bound;
} orc_config;

/* any pointer may be NULL */
typedef struct {
    double *ys;        /* [B][m][n] dense output at ts (NaN where not reached) */
    double *t_end;     /* [B] */
    double *y_end;     /* [B][n] */
    double *f_end;     /* [B][n] */
    double *h0;        /* [B] first step size (after start-up) */
    double *h_end;     /* [B] next proposed step size */
    long *n_accepted;  /* [B] */
    long *n_rejected;  /* [B] */
    int *n_out;        /* [B] number of ts points emitted */
    int *status;       /* [B] 0 ok, 1 stopped by t_bound, 2 failed */
} orc_result;

/* y0: [B][n] row-major; ts: [m] sorted.  Returns 0 on success. */
int orc_run_ensemble(const orc_config *cfg, long B, const double *y0, const double *params,
                     const double *ts, int m, const orc_result *res, int nthreads);
int orc_version(void);

#ifdef __cplusplus
}
#endif
#endif
