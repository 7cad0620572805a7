/*
 * nbk_oracle.c -- plain-C restatement of nbkode's explicit hot path for whole ensembles.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * CPU-baseline legs of bench.py may load this library.  The product (libnbkode_b200.so)
 * never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_c.py checks this file against golden vectors
 * produced by the live reference (tests/golden/make_golden.py) and against the bit-exact
 * NumPy restatement (oracle/nbk_oracle_np.py).  Tolerance of the pin: identical accepted-step
 * counts; dense output at fixed times within 1e-12 relative on non-chaotic cases (1e-9 for
 * Lorenz to t=10); fixed-step results within 1e-12; the adaptive step sequence itself
 * (t_end, h_end) within 1e-8 (summation order inside BLAS/NumPy is not reproducible in C and
 * the controller feeds ulp-level changes of E@K back into h; everything else follows the
 * reference's association order).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, IEEE double).
 *
 * Reference lines restated (paths under /root/reference):
 *   history window            nbkode/buffer.py:12-56
 *   ctor, f0                  nbkode/core.py:106-141
 *   initial step              nbkode/core.py:695-705 -> scipy _ivp.common.select_initial_step
 *   FSAL stage loop           nbkode/runge_kutta/explicit.py:60-79
 *   error / norm / controller nbkode/runge_kutta/core.py:73-99
 *   accept loop               nbkode/runge_kutta/explicit.py:81-99
 *   t_bound guard             nbkode/core.py:176-184
 *   RK dense output           nbkode/runge_kutta/explicit.py:354-403
 *   AB start-up               nbkode/multistep/core.py:37-65
 *   AB step                   nbkode/multistep/core.py:71-112, multistep/adams_bashforth.py
 *   window Hermite            nbkode/core.py:577-596
 *   run(ts) driver            nbkode/core.py:598-619
 */
#include "nbk_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RHS catalogue */

static void rhs_decay(double t, const double *y, const double *p, double *dy, int n, int np) {
    (void)t;
    for (int i = 0; i < n; ++i) dy[i] = -p[np == 1 ? 0 : i] * y[i];
}
static void rhs_lorenz(double t, const double *y, const double *p, double *dy, int n, int np) {
    (void)t; (void)n; (void)np;
    dy[0] = p[0] * (y[1] - y[0]);
    dy[1] = y[0] * (p[1] - y[2]) - y[1];
    dy[2] = y[0] * y[1] - p[2] * y[2];
}
static void rhs_vdp(double t, const double *y, const double *p, double *dy, int n, int np) {
    (void)t; (void)n; (void)np;
    dy[0] = y[1];
    dy[1] = p[0] * (1 - y[0] * y[0]) * y[1] - y[0];
}
static void rhs_linear(double t, const double *y, const double *A, double *dy, int n, int np) {
    (void)t; (void)np;
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s += A[(size_t)i * n + j] * y[j];
        dy[i] = s;
    }
}
static void rhs_rd1d(double t, const double *y, const double *p, double *dy, int n, int np) {
    (void)t; (void)np;
    const double d = p[0], r = p[1];
    for (int i = 0; i < n; ++i) {
        const double yr = y[i + 1 == n ? 0 : i + 1], yl = y[i == 0 ? n - 1 : i - 1];
        dy[i] = d * (yr - 2 * y[i] + yl) + r * y[i] * (1 - y[i]);
    }
}

typedef void (*rhs_fn)(double, const double *, const double *, double *, int, int);
static rhs_fn rhs_table[] = {rhs_decay, rhs_lorenz, rhs_vdp, rhs_linear, rhs_rd1d};

/* ------------------------------------------------------------------ tableaux */

typedef struct {
    int S;      /* stages without the FSAL one */
    int q;      /* error estimator order */
    int m;      /* dense-output polynomial degree (columns of P) */
    double A[6][6], B[6], C[6], E[7], P[7][4];
} tableau;

static tableau BS23, DP45;
static pthread_once_t tab_once = PTHREAD_ONCE_INIT;

static void init_tableaux(void) {
    memset(&BS23, 0, sizeof BS23);
    memset(&DP45, 0, sizeof DP45);
    /* Bogacki-Shampine 3(2): explicit.py:146-151, E via FSAL.E explicit.py:54-58 */
    BS23.S = 3; BS23.q = 2; BS23.m = 3;
    BS23.C[1] = 1.0 / 2; BS23.C[2] = 3.0 / 4;
    BS23.A[1][0] = 1.0 / 2; BS23.A[2][1] = 3.0 / 4;
    BS23.B[0] = 2.0 / 9; BS23.B[1] = 1.0 / 3; BS23.B[2] = 4.0 / 9;
    {
        const double B2[4] = {7.0 / 24, 1.0 / 4, 1.0 / 3, 1.0 / 8};
        for (int j = 0; j < 4; ++j) BS23.E[j] = -B2[j];
        for (int j = 0; j < 3; ++j) BS23.E[j] += BS23.B[j];
    }
    {
        const double P[4][3] = {{1, -4.0 / 3, 5.0 / 9}, {0, 1, -2.0 / 3}, {0, 4.0 / 3, -8.0 / 9}, {0, -1, 1}};
        for (int j = 0; j < 4; ++j) for (int k = 0; k < 3; ++k) BS23.P[j][k] = P[j][k];
    }
    /* Dormand-Prince 5(4): explicit.py:185-239 */
    DP45.S = 6; DP45.q = 4; DP45.m = 4;
    {
        const double C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
        const double B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
        const double E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
        const double A[6][6] = {
            {0},
            {1.0 / 5},
            {3.0 / 40, 9.0 / 40},
            {44.0 / 45, -56.0 / 15, 32.0 / 9},
            {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729},
            {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
        const double P[7][4] = {
            {1, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
            {0, 0, 0, 0},
            {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
            {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
            {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
            {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
            {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};
        memcpy(DP45.C, C, sizeof C); memcpy(DP45.B, B, sizeof B); memcpy(DP45.E, E, sizeof E);
        memcpy(DP45.A, A, sizeof A); memcpy(DP45.P, P, sizeof P);
    }
}

/* Adams-Bashforth weights as listed by the reference; index 0 multiplies the OLDEST f. */
static const double AB_B[6][5] = {
    {0},
    {1.0},
    {3.0 / 2, -1.0 / 2},
    {23.0 / 12, -16.0 / 12, 5.0 / 12},
    {55.0 / 24, -59.0 / 24, 37.0 / 24, -9.0 / 24},
    {1901.0 / 720, -2774.0 / 720, 2616.0 / 720, -1274.0 / 720, 251.0 / 720}};

/* Fixed-step explicit RK tableaux (explicit.py:102-129): Runge2 202, Runge3 203, Heun3 213,
   RungeKutta4 204, RungeKutta3_8 238 */
typedef struct { int id, S; double A[4][4], B[4], C[4]; } erk_tab;
static const erk_tab ERK_TABS[5] = {
    {202, 2, {{0}, {1.0 / 2}}, {0, 1.0}, {0, 1.0 / 2}},
    {203, 4, {{0}, {1.0 / 2}, {0, 1}, {0, 0, 1}}, {1.0 / 6, 4.0 / 6, 0, 1.0 / 6}, {0, 1.0 / 2, 1, 1}},
    {213, 3, {{0}, {1.0 / 3}, {0, 2.0 / 3}}, {1.0 / 4, 0, 3.0 / 4}, {0, 1.0 / 3, 2.0 / 3}},
    {204, 4, {{0}, {1.0 / 2}, {0, 1.0 / 2}, {0, 0, 1}}, {1.0 / 6, 2.0 / 6, 2.0 / 6, 1.0 / 6}, {0, 1.0 / 2, 1.0 / 2, 1}},
    {238, 4, {{0}, {1.0 / 3}, {-1.0 / 3, 1}, {1, -1, 1}}, {1.0 / 8, 3.0 / 8, 3.0 / 8, 1.0 / 8}, {0, 1.0 / 3, 2.0 / 3, 1}}};
static const erk_tab *erk_lookup(int id) {
    for (int i = 0; i < 5; ++i) if (ERK_TABS[i].id == id) return &ERK_TABS[i];
    return 0;
}

/* ------------------------------------------------------------------ single-trajectory solver */

typedef struct {
    const orc_config *cfg;
    const tableau *tab;  /* NULL for Adams-Bashforth / fixed-step RK */
    const erk_tab *erk;  /* fixed-step explicit RK, else NULL */
    rhs_fn rhs;
    const double *p;
    int n, L, order;
    double *ts, *ys, *fs;     /* [L], [L][n], [L][n] oldest -> newest */
    double *K;                /* [S+1][n] */
    double *ytmp, *ynew, *work;
    double h;
    long n_acc, n_rej;
    double rtol, atol;
} solver;

static void push(solver *s, double t, const double *y, const double *f) {
    const int n = s->n, L = s->L;
    memmove(s->ts, s->ts + 1, sizeof(double) * (L - 1));
    memmove(s->ys, s->ys + n, sizeof(double) * (size_t)n * (L - 1));
    memmove(s->fs, s->fs + n, sizeof(double) * (size_t)n * (L - 1));
    s->ts[L - 1] = t;
    memcpy(s->ys + (size_t)n * (L - 1), y, sizeof(double) * n);
    memcpy(s->fs + (size_t)n * (L - 1), f, sizeof(double) * n);
}

#define CUR_T(s) ((s)->ts[(s)->L - 1])
#define CUR_Y(s) ((s)->ys + (size_t)(s)->n * ((s)->L - 1))
#define CUR_F(s) ((s)->fs + (size_t)(s)->n * ((s)->L - 1))

static double rms(const double *x, const double *scale, int n, int divide) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const double v = divide ? x[i] / scale[i] : x[i];
        acc += v * v;
    }
    return sqrt(acc) / sqrt((double)n);
}

static double initial_step(solver *s, int q) {
    const int n = s->n;
    const double *y0 = CUR_Y(s), *f0 = CUR_F(s);
    const double t0 = CUR_T(s);
    double *scale = s->work, *y1 = s->ytmp, *f1 = s->ynew;
    for (int i = 0; i < n; ++i) scale[i] = s->atol + fabs(y0[i]) * s->rtol;
    const double d0 = rms(y0, scale, n, 1), d1 = rms(f0, scale, n, 1);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    for (int i = 0; i < n; ++i) y1[i] = y0[i] + h0 * 1 * f0[i];
    s->rhs(t0 + h0 * 1, y1, s->p, f1, n, s->cfg->np);
    for (int i = 0; i < n; ++i) f1[i] = f1[i] - f0[i];
    const double d2 = rms(f1, scale, n, 1) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / (q + 1));
    return fmin(100 * h0, h1);
}

static void solver_alloc(solver *s, const orc_config *cfg, const tableau *tab, int order,
                         double rtol, double atol, const double *p) {
    memset(s, 0, sizeof *s);
    s->cfg = cfg; s->tab = tab; s->rhs = rhs_table[cfg->rhs]; s->p = p;
    s->n = cfg->n; s->order = order;
    s->L = tab ? 2 : (order > 2 ? order : 2);
    s->rtol = rtol; s->atol = atol;
    const size_t n = (size_t)s->n;
    s->ts = malloc(sizeof(double) * s->L);
    s->ys = malloc(sizeof(double) * n * s->L);
    s->fs = malloc(sizeof(double) * n * s->L);
    s->K = calloc(n * 8, sizeof(double));
    s->ytmp = malloc(sizeof(double) * n);
    s->ynew = malloc(sizeof(double) * n);
    s->work = malloc(sizeof(double) * n * 5);
}

static void solver_free(solver *s) {
    free(s->ts); free(s->ys); free(s->fs); free(s->K); free(s->ytmp); free(s->ynew); free(s->work);
}

static void solver_start(solver *s, double t0, const double *y0) {
    double *f0 = s->work;
    s->rhs(t0, y0, s->p, f0, s->n, s->cfg->np);
    /* AlignedBuffer ctor: every slot holds (t0, y0, f0) */
    for (int l = 0; l < s->L; ++l) {
        s->ts[l] = t0;
        memcpy(s->ys + (size_t)s->n * l, y0, sizeof(double) * s->n);
        memcpy(s->fs + (size_t)s->n * l, f0, sizeof(double) * s->n);
    }
    s->n_acc = s->n_rej = 0;
}

/* one accepted adaptive step; returns 0 without stepping if t + h > bound */
static int rk_step(solver *s, double bound) {
    const tableau *T = s->tab;
    const int n = s->n, S = T->S;
    const orc_config *c = s->cfg;
    if (CUR_T(s) + s->h > bound) return 0;
    double *K = s->K;
    for (;;) {
        const double t = CUR_T(s), h = s->h;
        const double *y = CUR_Y(s);
        memcpy(K, CUR_F(s), sizeof(double) * n);
        for (int st = 1; st < S; ++st) {
            for (int i = 0; i < n; ++i) {
                double dy = T->A[st][0] * K[i];
                for (int j = 1; j < st; ++j) dy += T->A[st][j] * K[(size_t)j * n + i];
                s->ytmp[i] = y[i] + h * dy;
            }
            s->rhs(t + h * T->C[st], s->ytmp, s->p, K + (size_t)st * n, n, c->np);
        }
        const double t_new = t + h;
        for (int i = 0; i < n; ++i) {
            double acc = (h * T->B[0]) * K[i];
            for (int j = 1; j < S; ++j) acc += (h * T->B[j]) * K[(size_t)j * n + i];
            s->ynew[i] = y[i] + acc;
        }
        s->rhs(t_new, s->ynew, s->p, K + (size_t)S * n, n, c->np);
        double acc2 = 0.0;
        for (int i = 0; i < n; ++i) {
            double e = T->E[0] * K[i];
            for (int j = 1; j <= S; ++j) e += T->E[j] * K[(size_t)j * n + i];
            e = h * e;
            const double sc = s->atol + s->rtol * fmax(fabs(s->ynew[i]), fabs(y[i]));
            const double r = e / sc;
            acc2 += r * r;
        }
        const double norm = sqrt(acc2 / n);
        int ok;
        if (norm == 0) {
            s->h *= c->max_factor;
            ok = 1;
        } else {
            const double factor = c->safety / pow(norm, 1.0 / (T->q + 1));
            s->h *= fmin(c->max_factor, fmax(c->min_factor, factor));
            ok = norm < 1;
        }
        if (ok) {
            push(s, t_new, s->ynew, K + (size_t)S * n);
            s->n_acc++;
            return 1;
        }
        s->n_rej++;
        if (!(s->h > 0) || s->n_rej > c->max_rejects) return -1; /* non-finite guard (reference would spin) */
    }
}

static void rk_interp(const solver *s, double te, double *out) {
    const tableau *T = s->tab;
    const int n = s->n, S = T->S, m = T->m;
    if (te == s->ts[1]) { memcpy(out, s->ys + n, sizeof(double) * n); return; }
    if (te == s->ts[0]) { memcpy(out, s->ys, sizeof(double) * n); return; }
    const double t_old = s->ts[0], H = s->ts[1] - t_old;
    const double x = (te - t_old) / H;
    double pw[4];
    pw[0] = x;
    for (int k = 1; k < m; ++k) pw[k] = pw[k - 1] * x;
    for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int k = 0; k < m; ++k) {
            double q = 0.0;
            for (int j = 0; j <= S; ++j) q += s->K[(size_t)j * n + i] * T->P[j][k];
            acc += q * pw[k];
        }
        out[i] = H * acc + s->ys[i];
    }
}

static int ab_step(solver *s, double bound) {
    const int n = s->n, k = s->order, L = s->L, off = L - k;
    if (CUR_T(s) + s->h > bound) return 0;
    const double h = s->h, t_new = CUR_T(s) + h;
    const double *Bc = AB_B[k];
    for (int i = 0; i < n; ++i) {
        double acc = (h * Bc[0]) * s->fs[(size_t)off * n + i];
        double ay = (k == 1 ? -1.0 : 0.0) * s->ys[(size_t)off * n + i];
        for (int j = 1; j < k; ++j) {
            acc += (h * Bc[j]) * s->fs[(size_t)(off + j) * n + i];
            ay += (j == k - 1 ? -1.0 : 0.0) * s->ys[(size_t)(off + j) * n + i];
        }
        s->ynew[i] = acc - ay;
    }
    s->rhs(t_new, s->ynew, s->p, s->ytmp, n, s->cfg->np);
    push(s, t_new, s->ynew, s->ytmp);
    s->n_acc++;
    return 1;
}

static void hermite_interp(const solver *s, double te, double *out) {
    const int n = s->n, L = s->L;
    const double t0 = s->ts[0];
    if (te == t0) { memcpy(out, s->ys, sizeof(double) * n); return; }
    const double dt = s->ts[L - 1] - t0;
    const double T = (te - t0) / dt;
    const double *y0 = s->ys, *y1 = CUR_Y(s), *f0 = s->fs, *f1 = CUR_F(s);
    for (int i = 0; i < n; ++i) {
        const double dy = y1[i] - y0[i];
        out[i] = y0[i] + T * dy + T * (T - 1) * ((1 - 2 * T) * dy + dt * ((T - 1) * f0[i] + T * f1[i]));
    }
}

/* fixed-step explicit RK: explicit.py:28-46 + runge_kutta/core.py:33-42 */
static int erk_step(solver *s, double bound) {
    const erk_tab *T = s->erk;
    const int n = s->n, S = T->S;
    if (CUR_T(s) + s->h > bound) return 0;
    const double t = CUR_T(s), h = s->h;
    const double *y = CUR_Y(s);
    double *K = s->K;
    memcpy(K, CUR_F(s), sizeof(double) * n);
    for (int st = 1; st < S; ++st) {
        for (int i = 0; i < n; ++i) {
            double dy = T->A[st][0] * K[i];
            for (int j = 1; j < st; ++j) dy += T->A[st][j] * K[(size_t)j * n + i];
            s->ytmp[i] = y[i] + h * dy;
        }
        s->rhs(t + h * T->C[st], s->ytmp, s->p, K + (size_t)st * n, n, s->cfg->np);
    }
    const double t_new = t + h;
    for (int i = 0; i < n; ++i) {
        double acc = (h * T->B[0]) * K[i];
        for (int j = 1; j < S; ++j) acc += (h * T->B[j]) * K[(size_t)j * n + i];
        s->ynew[i] = y[i] + acc;
    }
    s->rhs(t_new, s->ynew, s->p, s->ytmp, n, s->cfg->np);
    push(s, t_new, s->ynew, s->ytmp);
    s->n_acc++;
    return 1;
}

static int any_step(solver *s, double bound) {
    return s->tab ? rk_step(s, bound) : (s->erk ? erk_step(s, bound) : ab_step(s, bound));
}
static void any_interp(const solver *s, double te, double *out) {
    if (s->tab) rk_interp(s, te, out); else hermite_interp(s, te, out);
}

/* AB start-up, multistep/core.py:37-65 */
static int ab_startup(solver *s, int first_stepper) {
    const orc_config *c = s->cfg;
    const int k = s->order, n = s->n;
    if (first_stepper == 0 || k == 1) return 0;
    solver sub;
    if (first_stepper == 45 || first_stepper == 23) {
        /* adaptive first stepper with DEFAULT tolerances; dense output at j*h (absolute times) */
        orc_config dc = *c;
        dc.min_factor = 0.2; dc.max_factor = 10.0; dc.safety = 0.9;
        const tableau *T = first_stepper == 45 ? &DP45 : &BS23;
        solver_alloc(&sub, c, T, 0, 1e-3, 1e-6, s->p);
        sub.cfg = &dc;
        solver_start(&sub, CUR_T(s), CUR_Y(s));
        sub.h = initial_step(&sub, T->q);
        double *yi = malloc(sizeof(double) * n * 2), *fi = yi + n;
        for (int j = 1; j < k; ++j) {
            const double ti = j * s->h;
            if (ti > sub.ts[1]) {
                while (sub.ts[1] < ti)
                    if (rk_step(&sub, INFINITY) <= 0) { free(yi); solver_free(&sub); return -1; }
            }
            rk_interp(&sub, ti, yi);
            s->rhs(ti, yi, s->p, fi, n, c->np);
            push(s, ti, yi, fi);
        }
        free(yi);
    } else {
        /* fixed-step first stepper: built with ITS default start-up ("auto" = RungeKutta45),
           multistep/core.py:54-59 constructs first_stepper_cls(rhs, t, y, h=h) */
        solver_alloc(&sub, c, NULL, first_stepper, 0, 0, s->p);
        solver_start(&sub, CUR_T(s), CUR_Y(s));
        sub.h = s->h;
        if (ab_startup(&sub, 45) != 0) { solver_free(&sub); return -1; }
        for (int j = 1; j < k; ++j) {
            ab_step(&sub, INFINITY);
            push(s, CUR_T(&sub), CUR_Y(&sub), CUR_F(&sub));
        }
    }
    solver_free(&sub);
    return 0;
}

/* ------------------------------------------------------------------ ensemble driver */

typedef struct {
    const orc_config *cfg;
    long B;
    const double *y0, *params, *ts;
    int m;
    orc_result res;
    long next;          /* dynamic work counter */
    long chunk;
    pthread_mutex_t mu;
} job;

static void run_one(const job *J, long i) {
    const orc_config *c = J->cfg;
    const int n = c->n;
    const double *p = c->params_shared ? J->params : J->params + (size_t)i * c->np;
    const int adaptive = c->method == 23 || c->method == 45;
    const tableau *T = c->method == 45 ? &DP45 : (c->method == 23 ? &BS23 : NULL);
    solver s;
    const erk_tab *erk = erk_lookup(c->method);
    solver_alloc(&s, c, T, (adaptive || erk) ? 0 : c->method, c->rtol, c->atol, p);
    s.erk = erk;
    solver_start(&s, c->t0, J->y0 + (size_t)i * n);
    int status = 0;
    if (adaptive) {
        s.h = isnan(c->first_step) ? initial_step(&s, T->q) : c->first_step;
    } else if (erk) {
        s.h = c->h;
    } else {
        s.h = c->h;
        if (ab_startup(&s, c->first_stepper) != 0) status = 2;
    }
    if (J->res.h0) J->res.h0[i] = s.h;
    int k = 0;
    /* run(ts): nbkode/core.py:401-408 (past points) + :598-619 (_run_eval) */
    for (; k < J->m && status == 0; ++k) {
        const double ti = J->ts[k];
        if (ti > CUR_T(&s)) {
            while (CUR_T(&s) < ti) {
                const int r = any_step(&s, c->t_bound);
                if (r == 0) { status = 1; break; }   /* t_bound guard fired: prefix only */
                if (r < 0) { status = 2; break; }    /* non-finite / reject storm */
            }
            if (status) break;
        }
        if (J->res.ys) any_interp(&s, ti, J->res.ys + ((size_t)i * J->m + k) * n);
    }
    if (J->res.ys)
        for (int kk = k; kk < J->m; ++kk)
            for (int q = 0; q < n; ++q) J->res.ys[((size_t)i * J->m + kk) * n + q] = NAN;
    if (J->res.t_end) J->res.t_end[i] = CUR_T(&s);
    if (J->res.y_end) memcpy(J->res.y_end + (size_t)i * n, CUR_Y(&s), sizeof(double) * n);
    if (J->res.f_end) memcpy(J->res.f_end + (size_t)i * n, CUR_F(&s), sizeof(double) * n);
    if (J->res.h_end) J->res.h_end[i] = s.h;
    if (J->res.n_accepted) J->res.n_accepted[i] = s.n_acc;
    if (J->res.n_rejected) J->res.n_rejected[i] = s.n_rej;
    if (J->res.n_out) J->res.n_out[i] = k;
    if (J->res.status) J->res.status[i] = status;
    solver_free(&s);
}

static void *worker(void *arg) {
    job *J = arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const long lo = J->next;
        J->next += J->chunk;
        pthread_mutex_unlock(&J->mu);
        if (lo >= J->B) break;
        const long hi = lo + J->chunk < J->B ? lo + J->chunk : J->B;
        for (long i = lo; i < hi; ++i) run_one(J, i);
    }
    return NULL;
}

int orc_run_ensemble(const orc_config *cfg, long B, const double *y0, const double *params,
                     const double *ts, int m, const orc_result *res, int nthreads) {
    if (!cfg || cfg->rhs < 0 || cfg->rhs > 4 || cfg->n < 1) return -1;
    if (!(cfg->method == 23 || cfg->method == 45 || (cfg->method >= 1 && cfg->method <= 5) || erk_lookup(cfg->method)))
        return -1;
    pthread_once(&tab_once, init_tableaux);
    job J;
    J.cfg = cfg; J.B = B; J.y0 = y0; J.params = params; J.ts = ts; J.m = m; J.res = *res;
    J.next = 0;
    if (nthreads < 1) nthreads = 1;
    J.chunk = B / (nthreads * 16L) + 1;
    if (J.chunk > 256) J.chunk = 256;
    pthread_mutex_init(&J.mu, NULL);
    if (nthreads == 1) {
        worker(&J);
    } else {
        pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
        for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker, &J);
        for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&J.mu);
    return 0;
}

int orc_version(void) { return 1; }
