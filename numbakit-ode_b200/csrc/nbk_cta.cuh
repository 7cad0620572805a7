// nbk_cta.cuh -- CTA-per-trajectory regime (K3) for large systems (method-of-lines grids,
// n up to a few thousand): one thread block integrates one trajectory.
//
//  * every thread owns E consecutive state components: its slice of y, of the stage derivatives
//    K0..K_S and of y_new live in registers (all loops over E are unrolled);
//  * the stage ARGUMENT vector y + h*sum(a_sj K_j) is staged in shared memory (two buffers used
//    alternately: one __syncthreads per stage) so that the right-hand side can read neighbours /
//    the whole vector;
//  * the scaled error norm is a warp-shuffle reduction followed by one shared-memory hop; every
//    thread then computes the same accept/reject decision and step-size factor, so a CTA never
//    diverges -- only CTAs differ in their number of steps, and they pull trajectories from the
//    same global queue as the per-thread kernels;
//  * state layout is trajectory-major (a trajectory's n values are contiguous, so the CTA's loads
//    and stores are coalesced): ts [B][2], ys/fs [B][2][n], K [B][S+1][n], h [B],
//    params [B][np] or shared [np].
//
// Reference rows restated: the same as nbk_device.cuh (FSAL stage loop explicit.py:60-79, error
// norm / controller runge_kutta/core.py:73-99, accept loop explicit.py:81-99, t_bound guard
// core.py:176-184, RkInterpolator explicit.py:354-403, select_initial_step core.py:695-705,
// drivers core.py:483-537, 598-619).
#ifndef NBK_CTA_CUH
#define NBK_CTA_CUH

#include "nbk_device.cuh"

namespace nbk {

// ---------------------------------------------------------------------------------------------------------------------
// The TEAM: the kernels of this file are written for T threads that integrate one trajectory together.  By default the
// team is the whole CTA.  With NBK_TEAM_G = G (2 .. 32) the team is G lanes of a warp and a block of NBK_GRP_BLOCK threads
// holds NBK_GRP_BLOCK / G teams -- the sub-warp regime for the solver classes that have no dedicated kernel in
// nbk_group.cuh (DOP853, Adams-Bashforth, the fixed-step Runge-Kutta classes on mid-size systems): the same source, with
// the barrier a __syncwarp of the team's lanes, the reduction a butterfly over them, and the team's shared memory a
// slice of the block's.  Teams of one warp run independently (different trajectories, different trip counts); every
// barrier names exactly the team's lanes.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef NBK_TEAM_G
#define NBK_TEAM_G 0
#endif
#ifndef NBK_GRP_BLOCK
#define NBK_GRP_BLOCK 128
#endif

#ifndef NBK_HOST_EMUL  // (tests/emul/nbk_emul_group.cpp provides these for a warp of 32 POSIX threads)
NBK_DEV void group_sync(unsigned gmask) { __syncwarp(gmask); }
NBK_DEV double group_shfl_xor(unsigned gmask, double v, int o) { return __shfl_xor_sync(gmask, v, o); }
NBK_DEV long long group_bcast(unsigned gmask, long long v, int src_lane) { return __shfl_sync(gmask, v, src_lane); }
NBK_DEV bool warp_all(bool p) { return __all_sync(0xffffffffu, p); }
#endif

#if NBK_TEAM_G > 0
NBK_DEV unsigned team_mask() {
    return NBK_TEAM_G == 32 ? 0xffffffffu : (((1u << NBK_TEAM_G) - 1u) << ((threadIdx.x & 31) & ~(NBK_TEAM_G - 1)));
}
NBK_DEV int team_tid() { return (int)(threadIdx.x & (NBK_TEAM_G - 1)); }
NBK_DEV int team_size() { return NBK_TEAM_G; }
NBK_DEV void team_sync() { group_sync(team_mask()); }
NBK_DEV long long team_index() { return (long long)blockIdx.x * (NBK_GRP_BLOCK / NBK_TEAM_G) + threadIdx.x / NBK_TEAM_G; }
NBK_DEV long long team_count() { return (long long)gridDim.x * (NBK_GRP_BLOCK / NBK_TEAM_G); }
NBK_DEV double *team_smem(double *sm, int doubles_per_team) { return sm + (size_t)(threadIdx.x / NBK_TEAM_G) * doubles_per_team; }
#else
NBK_DEV int team_tid() { return (int)threadIdx.x; }
NBK_DEV int team_size() { return (int)blockDim.x; }
NBK_DEV void team_sync() { __syncthreads(); }
NBK_DEV long long team_index() { return (long long)blockIdx.x; }
NBK_DEV long long team_count() { return (long long)gridDim.x; }
NBK_DEV double *team_smem(double *sm, int) { return sm; }
#endif

// Sum over the team; every thread returns the same value (same summation order everywhere).
// The caller guarantees a team barrier between two uses of `red` (the stage barriers do).
NBK_DEV double block_sum(double v, double *red, int nwarps) {
#if NBK_TEAM_G > 0
    (void)red;
    (void)nwarps;
    const unsigned gmask = team_mask();
#pragma unroll
    for (int o = NBK_TEAM_G / 2; o > 0; o >>= 1) v += group_shfl_xor(gmask, v, o);  // (the same pair on both sides: equal totals)
    return v;
#else
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31;
    if (lane == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (nwarps == 16) {
        // the 512-thread blocks of the large systems: every thread reads the 16 warp partials as broadcast 16-byte
        // loads and adds them as a tree (one shared round trip + 4 dependent additions instead of 5 shuffle rounds)
        const double2 *r2 = reinterpret_cast<const double2 *>(red);
        double2 q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = r2[i];
        const double a = (q[0].x + q[0].y) + (q[1].x + q[1].y), b = (q[2].x + q[2].y) + (q[3].x + q[3].y);
        const double c = (q[4].x + q[4].y) + (q[5].x + q[5].y), d = (q[6].x + q[6].y) + (q[7].x + q[7].y);
        return (a + b) + (c + d);
    }
    // every warp reduces the (<= 32) warp partials with the same butterfly: identical total everywhere
    double tot = lane < nwarps ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    return tot;
#endif
}

// The stage-argument vector in shared memory.  Thread `tid` owns the E consecutive components
// i = tid*E + e, but they are stored TRANSPOSED, component (tid, e) at base[e*T + tid]: for a fixed e
// the 32 lanes of a warp touch 32 consecutive doubles, so neither the stores of the owners nor the
// halo reads of a stencil have bank conflicts (the natural order would be a 4-way conflict for E = 4;
// measured 65 % of all shared wavefronts were replays).  operator[] takes the natural index.
template <int E>
struct smem_vec {
    const double *base;
    int T;
    NBK_DEV double operator[](int i) const { return base[(i % E) * T + i / E]; }
    NBK_DEV double at(int tid, int e) const { return base[e * T + tid]; }
};

// =====================================================================================
// Right-hand sides of the CTA regime:
//   struct prm;  static prm load(const double *p)      -- a trajectory's parameters, fetched ONCE per trajectory
//                                                         (a pointer for parameter blocks that do not fit registers)
//   template <int E, bool FULL> static void eval(t, yown[E], ysm (whole vector, shared memory), prm, i0, n, dy[E])
// `yown` is the thread's slice of the vector that is also in `ysm`.  FULL: n == team_size() * E, every thread owns E
// valid components -- the kernels then compile without any per-component bounds check (branch-free stages whose
// E dependency chains interleave; with the checks the chains ran one after the other).
// =====================================================================================

template <bool FULL>
NBK_DEV bool cta_inb(int i, int n) { return FULL || i < n; }

struct rhs_rd1d_cta {  // periodic 1-D reaction-diffusion: d (y[i+1] - 2 y[i] + y[i-1]) + r y (1 - y)
    static constexpr int NP = 2;
    struct prm { double d, r; };
    static NBK_DEV prm load(const double *p) { return prm{p[0], p[1]}; }
    // 5 FP64 instructions.  The association of the stencil matters: (right - 2 y) + left cancels exactly, whereas
    // (left + right) - 2 y rounds the sum first and feeds 2-4x more noise into the highest grid mode, which a
    // stability-limited explicit pair amplifies up to the tolerance (measured on the C5 ensemble with a CPU restatement:
    // 4.1x (rtol |y| + atol) between the two associations, 0.07x between this one and the reference's expression).
    static NBK_DEV double point(const prm &q, double left, double y, double right) {
        const double ry = q.r * y;
        return fma(q.d, fma(-2.0, y, right) + left, fma(-ry, y, ry));
    }
    template <int E, bool FULL = false>
    static NBK_DEV void eval(double, const double (&yown)[E], smem_vec<E> ysm, const prm &q, int i0, int n,
                             double (&dy)[E]) {
        // only the two halo values come from shared memory, the interior from registers
        if constexpr (FULL) {
            const int tid = team_tid(), T = ysm.T;
            double left = ysm.at(tid == 0 ? T - 1 : tid - 1, E - 1);
            const double halo_right = ysm.at(tid + 1 == T ? 0 : tid + 1, 0);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                dy[e] = point(q, left, yown[e], e + 1 < E ? yown[e + 1 < E ? e + 1 : e] : halo_right);
                left = yown[e];
            }
        } else {
            // ragged tail (n < team_size() * E): no branches either -- slots beyond n are computed from whatever
            // finite values sit there and never reach a reduction or a store
            double left = ysm[i0 == 0 ? n - 1 : i0 - 1];
            const int ilast = i0 + E - 1 < n ? i0 + E - 1 : n - 1;
            const double halo_right = ysm[ilast + 1 == n ? 0 : ilast + 1];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double right = (i0 + e == ilast) ? halo_right : yown[e + 1 < E ? e + 1 : e];
                dy[e] = point(q, left, yown[e], right);
                left = yown[e];
            }
        }
    }
};

struct rhs_linear_cta {  // y' = A y with a shared dense A; params holds A TRANSPOSED ([j][i]) so that
                         // consecutive threads read consecutive addresses; y[j] is a broadcast smem read
    static constexpr int NP = -1;  // n*n, run-time
    struct prm { const double *At; };
    static NBK_DEV prm load(const double *p) { return prm{p}; }
    template <int E, bool FULL = false>
    static NBK_DEV void eval(double, const double (&)[E], smem_vec<E> ysm, const prm &q, int i0, int n,
                             double (&dy)[E]) {
        const double *At = q.At;
        double acc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = 0.0;
        int col[E];  // slots beyond n recompute row n-1 (never stored): no branch in the inner loop
#pragma unroll
        for (int e = 0; e < E; ++e) col[e] = cta_inb<FULL>(i0 + e, n) ? i0 + e : n - 1;
        for (int j = 0; j < n; ++j) {
            const double yj = ysm[j];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = fma(__ldg(At + (long long)j * n + col[e]), yj, acc[e]);
        }
#pragma unroll
        for (int e = 0; e < E; ++e) dy[e] = acc[e];
    }
};

#ifdef NBK_USER_RHS_CTA
// User right-hand side of the CTA regime, compiled by NVRTC: component-wise
//   NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n)
// where y[j] reads component j of the whole stage-argument vector (shared memory).
struct rhs_user_cta {
    static constexpr int NP = -1;
#if defined(NBK_CTA_NP) && NBK_CTA_NP > 0
    // a small per-trajectory parameter block (NBK_CTA_NP <= 16 doubles, known when NVRTC compiles the program) is
    // copied into registers once per trajectory: the user's p[k] with constant k never touches memory again (it was
    // a global load behind every barrier)
    struct prm { double v[NBK_CTA_NP]; };
    static NBK_DEV prm load(const double *p) {
        prm q;
#pragma unroll
        for (int k = 0; k < NBK_CTA_NP; ++k) q.v[k] = p[k];
        return q;
    }
    static NBK_DEV const double *ptr(const prm &q) { return q.v; }
#else
    struct prm { const double *p; };
    static NBK_DEV prm load(const double *p) { return prm{p}; }
    static NBK_DEV const double *ptr(const prm &q) { return q.p; }
#endif
    template <int E, bool FULL = false>
    static NBK_DEV void eval(double t, const double (&)[E], smem_vec<E> ysm, const prm &q, int i0, int n,
                             double (&dy)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e)  // slots beyond n recompute component n-1 (never stored): no branches
            dy[e] = nbk_rhs_i(t, ysm, ptr(q), cta_inb<FULL>(i0 + e, n) ? i0 + e : n - 1, n);
    }
};
#endif

// =====================================================================================
// One FSAL attempt of the whole CTA.  K[0] holds this thread's slice of f(t, y).
// =====================================================================================
template <class Tab, class Rhs, int E, bool FULL = false>
NBK_DEV void cta_attempt(double t, double h, const double (&y)[E], double (&K)[Tab::S + 1][E], const typename Rhs::prm &p,
                         int i0, int n, double inv_n, double *buf0, double *buf1, double *red, int nwarps,
                         double (&ynew)[E], double &err2, double rtol, double atol) {
    constexpr int S = Tab::S;
#pragma unroll
    for (int s = 1; s <= S; ++s) {
        double ys[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < s; ++j) {
                const bool nz = s < S ? Tab::a(s < S ? s : 0, j) != 0.0 : Tab::b(j) != 0.0;
                if (nz) {
                    const double c = s < S ? Tab::cb().a[s < S ? s : 0][j] : Tab::cb().b[j];
                    acc = first ? c * K[j][e] : fma(c, K[j][e], acc);
                    first = false;
                }
            }
            ys[e] = fma(h, acc, y[e]);
        }
        double *buf = (s & 1) ? buf1 : buf0;
        const int T = team_size();
#pragma unroll
        for (int e = 0; e < E; ++e) buf[e * T + team_tid()] = ys[e];  // transposed, conflict-free (see smem_vec)
        team_sync();
        Rhs::template eval<E, FULL>(s < S ? fma(h, Tab::cb().c[s < S ? s : 0], t) : t + h, ys, smem_vec<E>{buf, T}, p, i0,
                                    n, K[s]);
        if (s == S) {
#pragma unroll
            for (int e = 0; e < E; ++e) ynew[e] = ys[e];
        }
    }
    if constexpr (Tab::DOP) {
        // DOP853: err5 = h E5.K, err3 = h E3.K, norm^2 = a^2 / (a + b / 100) with a = mean((err5/sc)^2), b = mean((err3/sc)^2)
        // (explicit.py:325-336; same squared form as the per-thread kernel)
        double p5 = 0.0, p3 = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            double e5 = 0.0, e3 = 0.0;
            bool f5 = true, f3 = true;
#pragma unroll
            for (int j = 0; j <= S; ++j) {
                if (Tab::e5(j) != 0.0) {
                    e5 = f5 ? Tab::cb().e5[j] * K[j][e] : fma(Tab::cb().e5[j], K[j][e], e5);
                    f5 = false;
                }
                if (Tab::e3(j) != 0.0) {
                    e3 = f3 ? Tab::cb().e3[j] * K[j][e] : fma(Tab::cb().e3[j], K[j][e], e3);
                    f3 = false;
                }
            }
            const double rs = fast_rcp(pos_max(fma(rtol, fabs(ynew[e]), atol), fma(rtol, fabs(y[e]), atol)));
            const double r5 = e5 * rs, r3 = e3 * rs;
            p5 = cta_inb<FULL>(i0 + e, n) ? fma(r5, r5, p5) : p5;
            p3 = cta_inb<FULL>(i0 + e, n) ? fma(r3, r3, p3) : p3;
        }
        const double hh = h * h * inv_n;
        const double a5 = block_sum(p5, red, nwarps) * hh;
        team_sync();  // `red` is reused
        const double a3 = block_sum(p3, red, nwarps) * hh;
        const double den = fma(0.01, a3, a5);
        err2 = den > 0.0 ? a5 * a5 / den : a5;  // exact step: the reference divides 0/0 here
    } else {
        double part = 0.0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            double ee = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j <= S; ++j) {
                if (Tab::e(j) != 0.0) {
                    ee = first ? Tab::cb().e[j] * K[j][e] : fma(Tab::cb().e[j], K[j][e], ee);
                    first = false;
                }
            }
            const double sc = pos_max(fma(rtol, fabs(ynew[e]), atol), fma(rtol, fabs(y[e]), atol));
            const double r = ee * fast_rcp(sc);
            part = cta_inb<FULL>(i0 + e, n) ? fma(r, r, part) : part;  // a select, not a branch
        }
        err2 = block_sum(part, red, nwarps) * (h * h * inv_n);  // inv_n = 1/n, computed once per kernel
    }
}

// f(t, y) of the whole CTA: stage the vector (transposed, see smem_vec), one barrier, evaluate the slice
template <class Rhs, int E>
NBK_DEV void cta_eval(double t, const double (&y)[E], const typename Rhs::prm &p, int i0, int n, double *buf,
                      double (&f)[E]) {
    const int T = team_size();
#pragma unroll
    for (int e = 0; e < E; ++e) buf[e * T + team_tid()] = y[e];
    team_sync();
#pragma unroll
    for (int e = 0; e < E; ++e) f[e] = 0.0;
    Rhs::template eval<E>(t, y, smem_vec<E>{buf, T}, p, i0, n, f);
}

// Dense output of one step for the whole CTA (RkInterpolator / DOP853Interpolator): prepare() builds the step's
// interpolant once -- for DOP853 that takes three more right-hand-side evaluations of the whole vector (rows 13..15 of
// the extended tableau, explicit.py:406-481), i.e. three barriers every thread of the CTA must reach -- and
// eval() evaluates a thread's slice at a time inside the step.
template <class Tab, class Rhs, int E>
struct cta_dense {
    double F[Tab::DOP ? 7 : E][Tab::DOP ? E : Tab::M];
    NBK_DEV void prepare(double t_old, double t, const double (&y_old)[E], const double (&y)[E],
                         const double (&K)[Tab::S + 1][E], const typename Rhs::prm &p, int i0, int n, double *buf0,
                         double *buf1) {
        if constexpr (Tab::DOP) {
            const double h = t - t_old;
            double Ke[3][E];
#pragma unroll
            for (int s = 13; s < 16; ++s) {
                double ys[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < s; ++j)
                        if (Tab::a(s, j) != 0.0)
                            acc = fma(Tab::cb().a[s][j], j < 13 ? K[j < 13 ? j : 0][e] : Ke[j >= 13 ? j - 13 : 0][e], acc);
                    ys[e] = fma(acc, h, y_old[e]);
                }
                team_sync();  // the previous users of the buffer are done
                cta_eval<Rhs, E>(fma(Tab::cb().c[s], h, t_old), ys, p, i0, n, (s & 1) ? buf1 : buf0, Ke[s - 13]);
            }
            // the last evaluation read buf1, which the first stage of the NEXT attempt writes without a barrier of its
            // own in between (stages rely on the alternation of the two buffers): close the window here.  Found by the
            // host emulation of the block (tests/test_cta_logic_host.py); real warps run close enough to lock-step
            // that the GPU tests never saw it.
            team_sync();
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double dy = y[e] - y_old[e];
                F[0][e] = dy;
                F[1][e] = h * K[0][e] - dy;
                F[2][e] = 2 * dy - h * (K[12][e] + K[0][e]);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (Tab::d(r, j) != 0.0)
                            acc = fma(Tab::cb().d[r][j], j < 13 ? K[j < 13 ? j : 0][e] : Ke[j >= 13 ? j - 13 : 0][e], acc);
                    F[3 + r][e] = h * acc;
                }
            }
        } else {
            rk_dense_coeffs<Tab, E>(K, F);
        }
    }
    NBK_DEV void eval(double te, double t_old, double t, const double (&y_old)[E], const double (&y)[E],
                      double (&out)[E]) const {
        if constexpr (Tab::DOP) dop853_eval<E>(te, t_old, t, y_old, y, F, out);
        else rk_dense_eval<Tab, E>(te, t_old, t, y_old, y, F, out);
    }
};

// rms(x / scale) over the whole vector (cold path)
template <int E>
NBK_DEV double cta_rms_scaled(const double (&x)[E], const double (&sc)[E], int i0, int n, double *red, int nwarps) {
    double part = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (i0 + e < n) {
            const double v = x[e] / sc[e];
            part = fma(v, v, part);
        }
    const double tot = block_sum(part, red, nwarps);
    team_sync();  // `red` is reused by the next norm
    return sqrt(tot) / sqrt((double)n);
}

// =====================================================================================
// Construction: f0, history, initial step (select_initial_step with block-wide norms)
// =====================================================================================
template <class Tab, class Rhs, int E>
NBK_DEV void cta_init_adaptive(const nbk_kargs &a, double *sm) {
    constexpr int S = Tab::S;
    const int n = a.n, tid = team_tid(), nwarps = team_size() >> 5, i0 = tid * E, npad = team_size() * E;
    sm = team_smem(sm, 2 * npad + 34);
    double *buf0 = sm, *buf1 = sm + npad, *red = sm + 2 * npad;
    for (long long idx = team_index(); idx < a.B; idx += team_count()) {
        team_sync();
        // per-trajectory parameters: copy the user's row into the state's [B][np] array
        if (!a.params_shared) {
            for (int c = tid; c < a.n_params; c += team_size())
                a.params[idx * a.n_params + c] = a.p0[idx * a.p0_sb + c * a.p0_sc];
            team_sync();  // made visible to the block before the first right-hand side call
        }
        const typename Rhs::prm p = Rhs::load(a.params_shared ? a.params : a.params + idx * a.n_params);
        double y[E], f[E], sc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            y[e] = i0 + e < n ? a.y0[idx * a.y0_sb + (long long)(i0 + e) * a.y0_sc] : 0.0;
            f[e] = 0.0;
            buf0[e * team_size() + tid] = y[e];
        }
        team_sync();
        Rhs::template eval<E>(a.t0, y, smem_vec<E>{buf0, team_size()}, p, i0, n, f);
        double h = a.h0 ? a.h0[idx] : a.h_init;
        if (!(h == h)) {
            // scipy select_initial_step (direction +1, unbounded), SURVEY.md row a4
#pragma unroll
            for (int e = 0; e < E; ++e) sc[e] = a.atol + fabs(y[e]) * a.rtol;
            const double d0 = cta_rms_scaled<E>(y, sc, i0, n, red, nwarps);
            const double d1 = cta_rms_scaled<E>(f, sc, i0, n, red, nwarps);
            const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
            double y1[E], f1[E];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                y1[e] = y[e] + h0 * f[e];
                f1[e] = 0.0;
                buf1[e * team_size() + tid] = y1[e];
            }
            team_sync();
            Rhs::template eval<E>(a.t0 + h0, y1, smem_vec<E>{buf1, team_size()}, p, i0, n, f1);
#pragma unroll
            for (int e = 0; e < E; ++e) f1[e] -= f[e];
            const double d2 = cta_rms_scaled<E>(f1, sc, i0, n, red, nwarps) / h0;
            double h1;
            if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
            else h1 = pow(0.01 / fmax(d1, d2), 1.0 / (Tab::Q + 1));
            h = fmin(100 * h0, h1);
        }
        const long long nn = n;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            if (i < n) {
                a.ys[(idx * 2 + 0) * nn + i] = y[e];
                a.ys[(idx * 2 + 1) * nn + i] = y[e];
                a.fs[(idx * 2 + 0) * nn + i] = f[e];
                a.fs[(idx * 2 + 1) * nn + i] = f[e];
#pragma unroll
                for (int s = 0; s <= S; ++s) a.K[(idx * (S + 1) + s) * nn + i] = 0.0;
            }
        }
        if (tid == 0) {
            a.ts[idx * 2] = a.t0;
            a.ts[idx * 2 + 1] = a.t0;
            a.h[idx] = h;
            a.n_acc[idx] = 0;
            a.n_rej[idx] = 0;
            a.status[idx] = NBK_TRAJ_OK;
        }
    }
}

#ifdef NBK_NEVENTS
// =====================================================================================
// run_events for a TEAM of threads that integrates one trajectory together -- the whole CTA here, a sub-warp group in
// nbk_group.cuh (nbkode/core.py:334-447, :621-647, event_handler.py, nbcompat/zeros.py:508-533).
// The event functions take the WHOLE state vector -- nbk_event(k, t, y, p) with plain pointers, the same source as in the
// per-thread regime -- so the team's threads publish their slices in natural order into `evbuf` (n doubles of shared
// memory) and every thread then evaluates the event functions redundantly: all threads see the same values, take the same
// branches, and the team stays converged through detection, bisection and recording.
//   Barrier: functor, the team's barrier.   Dense: functor (te, out[E]) -> the thread's slice of the step's dense output
//   (it may itself contain team barriers: every thread of the team calls it at the same places).
// =====================================================================================
template <int E, class Barrier>
NBK_DEV void team_publish(double *evbuf, const double (&v)[E], int i0, int n, Barrier sync) {
    sync();  // the readers of the previous vector are done
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (i0 + e < n) evbuf[i0 + e] = v[e];
    sync();
}

// build_handler(events, self.t, self.y): last_value = func(t, y), empty lists (event_handler.py:127-135)
template <int E, class Barrier>
NBK_DEV void team_events_start(const nbk_kargs &a, long long idx, double t, const double (&y)[E], const double *pp,
                               double (&ev_last)[NBK_NEVENTS], double *evbuf, bool lead, int i0, int n, Barrier sync) {
    team_publish<E>(evbuf, y, i0, n, sync);
#pragma unroll
    for (int k = 0; k < NBK_NEVENTS; ++k) {
        ev_last[k] = nbk_event(k, t, evbuf, pp);
        if (lead) a.ev_count[idx * NBK_NEVENTS + k] = 0;
    }
    sync();
}

// After an accepted step [t_old, t_new]: sign changes, bisection of each root on the step's dense output, recording.
// Returns true if the trajectory stops (terminal event -> t_term, or status = NBK_TRAJ_EVBRACKET).
template <int E, class Barrier, class Dense>
NBK_DEV bool team_events_after_step(const nbk_kargs &a, long long idx, double t_old, double t_new, const double (&y_new)[E],
                                    const double *pp, double (&ev_last)[NBK_NEVENTS], double *evbuf, bool lead, int i0, int n,
                                    Barrier sync, Dense &dense, double &t_term, int &status) {
    // every event function at the new point first (the bisections below reuse the buffer)
    double value[NBK_NEVENTS];
    team_publish<E>(evbuf, y_new, i0, n, sync);
#pragma unroll
    for (int k = 0; k < NBK_NEVENTS; ++k) value[k] = nbk_event(k, t_new, evbuf, pp);
    bool terminate = false;
#pragma unroll
    for (int k = 0; k < NBK_NEVENTS; ++k) {
        const double last = ev_last[k];
        const bool up = last <= 0.0 && value[k] >= 0.0, down = last >= 0.0 && value[k] <= 0.0;
        const int dir = a.ev_dir[k];
        const bool trigger = (up && dir > 0) || (down && dir < 0) || ((up || down) && dir == 0);
        ev_last[k] = value[k];
        if (!trigger) continue;
        double root = t_new, out[E];
        if (value[k] != 0.0) {
            double left = t_old, right = t_new;
            dense(left, out);
            team_publish<E>(evbuf, out, i0, n, sync);
            double yl = nbk_event(k, left, evbuf, pp);
            if (yl == 0.0) {
                root = left;
            } else if (!(yl * value[k] < 0.0)) {
                status = NBK_TRAJ_EVBRACKET;
                return true;
            } else {
                double mid = left;
                for (int it = 0; it < 50; ++it) {
                    mid = (left + right) / 2;
                    dense(mid, out);
                    team_publish<E>(evbuf, out, i0, n, sync);
                    const double ym = nbk_event(k, mid, evbuf, pp);
                    if (fabs(ym) <= 1.48e-8) break;
                    if ((ym > 0.0) == (yl > 0.0)) {
                        left = mid;
                        yl = ym;
                    } else {
                        right = mid;
                    }
                }
                root = mid;
            }
        }
        const long long slot = idx * NBK_NEVENTS + k;
        const int cnt = a.ev_count[slot];  // written by this team only; the barrier below orders it
        // "This is required to avoid duplicates" (event_handler.py:158-160)
        const bool dup = cnt > 0 && cnt <= a.ev_cap && a.ev_t[slot * a.ev_cap + cnt - 1] == root;
        sync();  // every thread has read the counter and the last time before the leader updates them
        if (!dup) {
            if (cnt < a.ev_cap) {
                dense(root, out);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (i0 + e < n) a.ev_y[(slot * a.ev_cap + cnt) * n + i0 + e] = out[e];
                if (lead) a.ev_t[slot * a.ev_cap + cnt] = root;
            }
            if (lead) a.ev_count[slot] = cnt + 1;
        }
        if ((a.ev_terminal >> k) & 1) {
            // EventHandler.evaluate: min_t is never raised, so the LAST terminal event of the list wins
            terminate = true;
            t_term = root;
        }
    }
    return terminate;
}

struct cta_barrier {
    NBK_DEV void operator()() const { team_sync(); }
};
// dense output of the step just accepted, for the whole CTA (prepared lazily; for DOP853 the preparation evaluates three
// more stages with block barriers -- every thread reaches it at the same call)
template <class Tab, class Rhs, int E>
struct cta_step_dense {
    double t_old, t_new;
    const double (&y_old)[E];
    const double (&y_new)[E];
    const double (&K)[Tab::S + 1][E];
    const typename Rhs::prm &p;
    int i0, n;
    double *buf0, *buf1;
    cta_dense<Tab, Rhs, E> dn;
    bool prepared;
    NBK_DEV cta_step_dense(double t0, double t1, const double (&y0)[E], const double (&y1)[E], const double (&K_)[Tab::S + 1][E],
                           const typename Rhs::prm &p_, int i0_, int n_, double *b0, double *b1)
        : t_old(t0), t_new(t1), y_old(y0), y_new(y1), K(K_), p(p_), i0(i0_), n(n_), buf0(b0), buf1(b1), prepared(false) {}
    NBK_DEV void operator()(double te, double (&out)[E]) {
        if (!prepared) {
            dn.prepare(t_old, t_new, y_old, y_new, K, p, i0, n, buf0, buf1);
            prepared = true;
        }
        dn.eval(te, t_old, t_new, y_old, y_new, out);
    }
};
#endif  // NBK_NEVENTS

// =====================================================================================
// Advance: persistent CTAs pulling trajectories from the global queue.
// =====================================================================================
template <class Tab, class Rhs, int E, int MODE, bool FULL = false, bool EV = false>
NBK_DEV void cta_advance(const nbk_kargs &a, double *sm) {
    constexpr int S = Tab::S;
    const int n = a.n, tid = team_tid(), nwarps = team_size() >> 5, i0 = tid * E, npad = team_size() * E;
    const long long nn = n;
    const double inv_n = 1.0 / (double)n;
    sm = team_smem(sm, (EV ? 3 : 2) * npad + 34);
    double *buf0 = sm, *buf1 = sm + npad, *red = sm + 2 * npad;
    long long *s_idx = (long long *)(red + 32);
#ifdef NBK_NEVENTS
    double *evbuf = sm + 2 * npad + 34;  // EV: the whole vector in natural order for the event functions (npad doubles)
    double ev_last[NBK_NEVENTS];
#endif
    const double rtol = a.rtol, atol = a.atol, bound = a.bound;
    const bool bounded = (__double2hiint(a.bound) & 0x7fffffff) < 0x7ff00000;  // finite (an integer test: stays off the FP64 pipe)  // (usually not: skips the guard of every accepted step)
    const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;

    for (;;) {
        team_sync();  // the previous trajectory's last readers of s_idx / the buffers are done
        if (tid == 0) *s_idx = (long long)atomicAdd(a.work_counter, 1ULL);
        team_sync();
        const long long idx = *s_idx;
        if (idx >= a.B) break;

        const typename Rhs::prm p = Rhs::load(a.params_shared ? a.params : a.params + idx * a.n_params);
        double t = a.ts[idx * 2 + 1], h = a.h[idx];
        double y[E], K[S + 1][E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            y[e] = cta_inb<FULL>(i, n) ? a.ys[(idx * 2 + 1) * nn + i] : 0.0;
#pragma unroll
            for (int s = 0; s < S; ++s) K[s][e] = 0.0;
            K[S][e] = cta_inb<FULL>(i, n) ? a.fs[(idx * 2 + 1) * nn + i] : 0.0;  // f(t, y): K[0] of the first attempt
        }
        int status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
        int nacc = 0, natt = 0, kout = 0;
        double te_next = d_inf();
        bool running = status == NBK_TRAJ_OK;
        if (running && MODE == NBK_MODE_RUN) {
            if (kout < a.m && __ldg(a.tgrid + kout) <= t) {
                // grid points inside the stored last step: answered from history (core.py:401-408)
                const double t_old = a.ts[idx * 2];
                double y_old[E], Kh[S + 1][E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = i0 + e;
                    y_old[e] = cta_inb<FULL>(i, n) ? a.ys[(idx * 2) * nn + i] : 0.0;
#pragma unroll
                    for (int s = 0; s <= S; ++s) Kh[s][e] = (cta_inb<FULL>(i, n) && !a.pristine) ? a.K[(idx * (S + 1) + s) * nn + i] : 0.0;
                }
                cta_dense<Tab, Rhs, E> dn;
                dn.prepare(t_old, t, y_old, y, Kh, p, i0, n, buf0, buf1);
                while (kout < a.m && __ldg(a.tgrid + kout) <= t) {
                    double out[E];
                    dn.eval(__ldg(a.tgrid + kout), t_old, t, y_old, y, out);
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                    ++kout;
                }
            }
            running = kout < a.m;
            if (running) te_next = __ldg(a.tgrid + kout);
        } else if (running) {
            running = a.n_steps > 0;
        }
#ifdef NBK_NEVENTS
        const double *pp = a.params_shared ? a.params : a.params + idx * a.n_params;
        if constexpr (EV && MODE == NBK_MODE_RUN) {
            if (status == NBK_TRAJ_OK) team_events_start<E>(a, idx, t, y, pp, ev_last, evbuf, tid == 0, i0, n, cta_barrier{});
        }
#endif
        if (status >= NBK_TRAJ_NONFINITE && MODE == NBK_MODE_RUN) {
            for (int k = 0; k < a.m; ++k)
                for (int e = 0; e < E; ++e)
                    if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + k * a.o_sk + (long long)(i0 + e) * a.o_sc] = d_nan();
        }
        if (running && (t + h > bound)) {  // guard of the first step (core.py:176-184)
            status = NBK_TRAJ_BOUND;
            running = false;
        }
        bool fresh = true;
        while (running) {  // every quantity steering this loop is identical in all threads of the CTA
            if (fresh) {
#pragma unroll
                for (int e = 0; e < E; ++e) K[0][e] = K[S][e];
            }
            double ynew[E], err2;
            cta_attempt<Tab, Rhs, E, FULL>(t, h, y, K, p, i0, n, inv_n, buf0, buf1, red, nwarps, ynew, err2, rtol, atol);
            const bool accept = (unsigned)__double2hiint(err2) < 0x3ff00000u;
            const double t_new = t + h;
            h *= step_factor<Tab>(err2, a.safety, a.min_factor, a.max_factor, a.x_lo, a.x_hi);
            ++natt;
            bool done = false;
            bool stopped_by_event = false;
#ifdef NBK_NEVENTS
            if constexpr (EV && MODE == NBK_MODE_RUN) {
                if (accept) {
                    // events first, then the grid points inside the step (core.py:638-645)
                    cta_step_dense<Tab, Rhs, E> dense(t, t_new, y, ynew, K, p, i0, n, buf0, buf1);
                    double t_term = 0.0;
                    if (team_events_after_step<E>(a, idx, t, t_new, ynew, pp, ev_last, evbuf, tid == 0, i0, n, cta_barrier{}, dense,
                                                  t_term, status)) {
                        if (status != NBK_TRAJ_EVBRACKET) {
                            double out[E];
                            dense(t_term, out);
#pragma unroll
                            for (int e = 0; e < E; ++e)
                                if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                            if (tid == 0) a.ev_tterm[idx] = t_term;
                            ++kout;
                            status = NBK_TRAJ_EVENT;
                        }
                        done = true;
                        stopped_by_event = true;
                    }
                    team_sync();  // the event path's last readers of the stage buffers (DOP853 dense output) are done
                }
            }
#endif
            if (accept && !stopped_by_event) {
                if (MODE == NBK_MODE_RUN) {
                    if (t_new >= te_next) {  // the same for every thread of the CTA: prepare() may use barriers
                        cta_dense<Tab, Rhs, E> dn;
                        dn.prepare(t, t_new, y, ynew, K, p, i0, n, buf0, buf1);
                        do {
                            double out[E];
                            dn.eval(te_next, t, t_new, y, ynew, out);
#pragma unroll
                            for (int e = 0; e < E; ++e)
                                if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                            ++kout;
                            te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                        } while (t_new >= te_next);
                    }
                    done = kout >= a.m;
                } else {
                    if (a.emit) {
                        if (tid == 0) a.t_out[idx * a.to_sb + nacc] = t_new;
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + nacc * a.o_sk + (long long)(i0 + e) * a.o_sc] = ynew[e];
                    }
                    done = nacc + 1 >= a.n_steps;
                }
            }
            if (d_nonfinite(err2) || d_collapsed(h)) {
                status = NBK_TRAJ_NONFINITE;
                done = true;
            } else if (!done && natt >= max_att) {
                status = NBK_TRAJ_MAXATTEMPTS;
                done = true;
            }
            if (bounded && accept && !done && (t_new + h > bound)) {  // guard of the next step
                status = NBK_TRAJ_BOUND;
                done = true;
            }
            if (done) {
                // history slot 0 = start of the last accepted step (see advance_adaptive)
                if (tid == 0) a.ts[idx * 2] = t;
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (cta_inb<FULL>(i0 + e, n)) a.ys[(idx * 2) * nn + i0 + e] = y[e];
                running = false;
            }
            if (accept) {
                t = t_new;
#pragma unroll
                for (int e = 0; e < E; ++e) y[e] = ynew[e];
                ++nacc;
            }
            fresh = accept;
        }
        // ---- retire
        if (nacc > 0) {
            const bool clean = status != NBK_TRAJ_NONFINITE && status != NBK_TRAJ_MAXATTEMPTS;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = i0 + e;
                if (cta_inb<FULL>(i, n)) {
                    a.ys[(idx * 2 + 1) * nn + i] = y[e];
                    a.fs[(idx * 2) * nn + i] = K[0][e];
                    a.fs[(idx * 2 + 1) * nn + i] = clean ? K[S][e] : K[0][e];
#pragma unroll
                    for (int s = 0; s <= S; ++s) a.K[(idx * (S + 1) + s) * nn + i] = K[s][e];
                }
            }
        }
        if (tid == 0) {
            if (nacc > 0) a.ts[idx * 2 + 1] = t;
            a.h[idx] = h;
            if (a.pristine) {
                a.n_acc[idx] = nacc;
                a.n_rej[idx] = natt - nacc;
            } else {
                a.n_acc[idx] += nacc;
                a.n_rej[idx] += natt - nacc;
            }
            a.status[idx] = status;
            if (status >= NBK_TRAJ_NONFINITE) atomicAdd(a.work_counter + 1, 1ULL);
            a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : nacc;
        }
    }
}

// =====================================================================================
// Fixed-step solvers in the CTA regime: Adams-Bashforth-k / ForwardEuler (multistep/core.py:90-112) and the
// fixed-step explicit Runge-Kutta classes (explicit.py:28-46, 102-129) for large systems.  One trajectory per
// CTA, the history window / stage vectors of a thread's E components in registers, one barrier per right-hand
// side evaluation.  State layout: ts [B][L], ys/fs [B][L][n], K [B][S][n] (Runge-Kutta), h [B].
// =====================================================================================

// window Hermite on a thread's slice (nbkode/core.py:577-596)
template <int L, int E>
NBK_DEV void cta_hermite(double te, const double (&ts)[L], const double (&ys)[L][E], const double (&fs)[L][E], double (&out)[E]) {
    const double t0 = ts[0];
    if (te == t0) {
#pragma unroll
        for (int e = 0; e < E; ++e) out[e] = ys[0][e];
        return;
    }
    const double dt = ts[L - 1] - t0;
    const double T = (te - t0) / dt;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double dy = ys[L - 1][e] - ys[0][e];
        out[e] = ys[0][e] + T * dy + T * (T - 1) * ((1 - 2 * T) * dy + dt * ((T - 1) * fs[0][e] + T * fs[L - 1][e]));
    }
}

// Start-up points of an adaptive first stepper with DEFAULT tolerances (rtol 1e-3, atol 1e-6): dense output at the
// absolute times hfix, 2 hfix, ... pushed into the history window (multistep/core.py:60-65).  Returns false on failure.
template <class Tab, class Rhs, int E, int L>
NBK_DEV bool cta_rk_startup(double t0, const double (&y0)[E], const double (&f0)[E], const typename Rhs::prm &p, int i0, int n,
                            double hfix, int npts, double *buf0, double *buf1, double *red, int nwarps, double (&ts)[L],
                            double (&ys)[L][E], double (&fs)[L][E]) {
    constexpr int S = Tab::S;
    const double rtol = 1e-3, atol = 1e-6;
    double t = t0, t_old = t0, y[E], y_old[E], K[S + 1][E], sc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        y[e] = y_old[e] = y0[e];
#pragma unroll
        for (int s = 0; s <= S; ++s) K[s][e] = 0.0;
        K[S][e] = f0[e];
        sc[e] = atol + fabs(y0[e]) * rtol;
    }
    // select_initial_step with the default tolerances (same code path as cta_init_adaptive)
    const double d0 = cta_rms_scaled<E>(y0, sc, i0, n, red, nwarps);
    const double d1 = cta_rms_scaled<E>(f0, sc, i0, n, red, nwarps);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    double y1[E], f1[E];
#pragma unroll
    for (int e = 0; e < E; ++e) y1[e] = y0[e] + h0 * f0[e];
    team_sync();
    cta_eval<Rhs, E>(t0 + h0, y1, p, i0, n, buf1, f1);
#pragma unroll
    for (int e = 0; e < E; ++e) f1[e] -= f0[e];
    const double d2 = cta_rms_scaled<E>(f1, sc, i0, n, red, nwarps) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / (Tab::Q + 1));
    double h = fmin(100 * h0, h1);
    bool fresh = true;
    int budget = 1000000;
    const double inv_n = 1.0 / (double)n;
    for (int j = 1; j <= npts; ++j) {
        const double ti = j * hfix;
        while (t < ti) {
            if (fresh) {
#pragma unroll
                for (int e = 0; e < E; ++e) K[0][e] = K[S][e];
            }
            double ynew[E], err2;
            team_sync();
            cta_attempt<Tab, Rhs, E>(t, h, y, K, p, i0, n, inv_n, buf0, buf1, red, nwarps, ynew, err2, rtol, atol);
            const bool accept = pos_less(err2, 1.0);
            const double t_new = t + h;
            h *= step_factor<Tab>(err2, 0.9, 0.2, 10.0, 1e-13, 1e8);
            if (accept) {
                t_old = t;
                t = t_new;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    y_old[e] = y[e];
                    y[e] = ynew[e];
                }
            }
            fresh = accept;
            if (d_nonfinite(err2) || d_collapsed(h) || --budget < 0) return false;
        }
        double out[E], fo[E];
        rk_dense<Tab, E>(ti, t_old, t, y_old, y, K, out);
        team_sync();
        cta_eval<Rhs, E>(ti, out, p, i0, n, (j & 1) ? buf0 : buf1, fo);
        // push(ti, out, fo)
#pragma unroll
        for (int l = 0; l < L - 1; ++l) {
            ts[l] = ts[l + 1];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                ys[l][e] = ys[l + 1][e];
                fs[l][e] = fs[l + 1][e];
            }
        }
        ts[L - 1] = ti;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            ys[L - 1][e] = out[e];
            fs[L - 1][e] = fo[e];
        }
    }
    return true;
}

// METHOD: 1..5 Adams-Bashforth-k, 202/203/213/204/238 fixed-step Runge-Kutta
template <int METHOD>
struct cta_fixed_traits {
    static constexpr bool ERK = METHOD >= 200;
    static constexpr int L = ERK ? 2 : (METHOD > 2 ? METHOD : 2);
    static constexpr int S = ERK ? tab_erk<ERK ? METHOD : 204>::S : 0;
};

template <int METHOD, class Rhs, int E>
NBK_DEV void cta_init_fixed(const nbk_kargs &a, double *sm) {
    using TR = cta_fixed_traits<METHOD>;
    constexpr int L = TR::L, S = TR::S;
    const int n = a.n, tid = team_tid(), nwarps = team_size() >> 5, i0 = tid * E, npad = team_size() * E;
    const long long nn = n;
    sm = team_smem(sm, 2 * npad + 34);
    double *buf0 = sm, *buf1 = sm + npad, *red = sm + 2 * npad;
    for (long long idx = team_index(); idx < a.B; idx += team_count()) {
        team_sync();
        if (!a.params_shared) {
            for (int c = tid; c < a.n_params; c += team_size())
                a.params[idx * a.n_params + c] = a.p0[idx * a.p0_sb + c * a.p0_sc];
            team_sync();
        }
        const typename Rhs::prm p = Rhs::load(a.params_shared ? a.params : a.params + idx * a.n_params);
        double y[E], f[E];
#pragma unroll
        for (int e = 0; e < E; ++e) y[e] = i0 + e < n ? a.y0[idx * a.y0_sb + (long long)(i0 + e) * a.y0_sc] : 0.0;
        cta_eval<Rhs, E>(a.t0, y, p, i0, n, buf0, f);
        const double h = a.h0 ? a.h0[idx] : a.h_init;
        double ts[L], ys[L][E], fs[L][E];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            ts[l] = a.t0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                ys[l][e] = y[e];
                fs[l][e] = f[e];
            }
        }
        int status = NBK_TRAJ_OK;
        if (!TR::ERK && METHOD > 1 && a.first_stepper != 0) {
            bool ok = true;
            if (a.first_stepper == 45)
                ok = cta_rk_startup<tab_rk45, Rhs, E, L>(a.t0, y, f, p, i0, n, h, METHOD - 1, buf0, buf1, red, nwarps, ts, ys, fs);
            else
                ok = cta_rk_startup<tab_rk23, Rhs, E, L>(a.t0, y, f, p, i0, n, h, METHOD - 1, buf0, buf1, red, nwarps, ts, ys, fs);
            if (!ok) status = NBK_TRAJ_NONFINITE;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            if (i < n) {
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    a.ys[(idx * L + l) * nn + i] = ys[l][e];
                    a.fs[(idx * L + l) * nn + i] = fs[l][e];
                }
#pragma unroll
                for (int s = 0; s < S; ++s) a.K[(idx * S + s) * nn + i] = 0.0;
            }
        }
        if (tid == 0) {
#pragma unroll
            for (int l = 0; l < L; ++l) a.ts[idx * L + l] = ts[l];
            a.h[idx] = h;
            a.n_acc[idx] = 0;
            a.n_rej[idx] = 0;
            a.status[idx] = status;
            if (status >= NBK_TRAJ_NONFINITE) atomicAdd(a.work_counter + 1, 1ULL);
        }
    }
}

// window Hermite with the two ends given explicitly, for te behind the window start (the emission loop of the
// ring-buffered Adams-Bashforth path below)
template <int E>
NBK_DEV void cta_hermite_ends(double te, double t0, const double (&y0)[E], const double (&f0)[E], double t1,
                              const double (&y1)[E], const double (&f1)[E], double (&out)[E]) {
    const double dt = t1 - t0;
    const double T = (te - t0) / dt;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double dy = y1[e] - y0[e];
        out[e] = y0[e] + T * dy + T * (T - 1) * ((1 - 2 * T) * dy + dt * ((T - 1) * f0[e] + T * f1[e]));
    }
}

// history window of a CTA -> HBM, logical slot l living in physical slot (l + ROT) % L
template <int L, int E, int ROT>
NBK_DEV void cta_store_window(const nbk_kargs &a, long long idx, int n, int i0, const double (&ts)[L],
                              const double (&ys)[L][E], const double (&fs)[L][E]) {
    const long long nn = n;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = i0 + e;
        if (i < n) {
#pragma unroll
            for (int l = 0; l < L; ++l) {
                a.ys[(idx * L + l) * nn + i] = ys[(l + ROT) % L][e];
                a.fs[(idx * L + l) * nn + i] = fs[(l + ROT) % L][e];
            }
        }
    }
    if (team_tid() == 0) {
#pragma unroll
        for (int l = 0; l < L; ++l) a.ts[idx * L + l] = ts[(l + ROT) % L];
    }
}

#ifdef NBK_NEVENTS
// dense output of the fixed-step classes (the window Hermite of cta_hermite) as the functor team_events_after_step takes
template <int L, int E>
struct cta_window_dense {
    const double (&ts)[L];
    const double (&ys)[L][E];
    const double (&fs)[L][E];
    NBK_DEV cta_window_dense(const double (&t_)[L], const double (&y_)[L][E], const double (&f_)[L][E]) : ts(t_), ys(y_), fs(f_) {}
    NBK_DEV void operator()(double te, double (&out)[E]) { cta_hermite<L, E>(te, ts, ys, fs, out); }
};
#endif

// EV: the run_events variant (run mode only) -- after every step the team's event routine (team_events_after_step, shared
// with the adaptive kernels) brackets the previous point and the new one on the window Hermite, like the per-thread
// kernels (advance_ab / advance_erk in nbk_device.cuh) and the reference's run_events (core.py:621-647) do; the
// Adams-Bashforth classes of order >= 3 then take the shifting window below instead of the register ring.
template <int METHOD, class Rhs, int E, int MODE, bool EV = false>
NBK_DEV void cta_advance_fixed(const nbk_kargs &a, double *sm) {
    using TR = cta_fixed_traits<METHOD>;
    constexpr int L = TR::L, S = TR::S;
    constexpr int KS = S > 0 ? S : 1;
    const int n = a.n, tid = team_tid(), i0 = tid * E, npad = team_size() * E;
    const long long nn = n;
    sm = team_smem(sm, (EV ? 3 : 2) * npad + 34);
    double *buf0 = sm, *buf1 = sm + npad;
#ifdef NBK_NEVENTS
    double *evbuf = sm + 2 * npad + 34;  // EV: the whole vector in natural order for the event functions (npad doubles)
    double ev_last[NBK_NEVENTS];
#endif
    for (long long idx = team_index(); idx < a.B; idx += team_count()) {
        team_sync();
        const typename Rhs::prm p = Rhs::load(a.params_shared ? a.params : a.params + idx * a.n_params);
        double ts[L], ys[L][E], fs[L][E], K[KS][E];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            ts[l] = a.ts[idx * L + l];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = i0 + e;
                ys[l][e] = i < n ? a.ys[(idx * L + l) * nn + i] : 0.0;
                fs[l][e] = i < n ? a.fs[(idx * L + l) * nn + i] : 0.0;
            }
        }
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int e = 0; e < E; ++e) K[s][e] = 0.0;
        const double h = a.h[idx];
        int kout = 0, status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
        long long nsteps = 0;
        double te_next = d_inf();
        bool running;
        if (status != NBK_TRAJ_OK) {
            running = false;
            if (MODE == NBK_MODE_RUN)
                for (int k = 0; k < a.m; ++k)
                    for (int e = 0; e < E; ++e)
                        if (i0 + e < n) a.y_out[idx * a.o_sb + k * a.o_sk + (long long)(i0 + e) * a.o_sc] = d_nan();
        } else if (MODE == NBK_MODE_RUN) {
            while (kout < a.m && __ldg(a.tgrid + kout) <= ts[L - 1]) {
                double out[E];
                cta_hermite<L, E>(__ldg(a.tgrid + kout), ts, ys, fs, out);
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (i0 + e < n) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                ++kout;
            }
            running = kout < a.m;
            if (running) te_next = __ldg(a.tgrid + kout);
        } else {
            running = a.n_steps > 0;
        }
#ifdef NBK_NEVENTS
        const double *pp = a.params_shared ? a.params : a.params + idx * a.n_params;
        if constexpr (EV && MODE == NBK_MODE_RUN) {
            if (status == NBK_TRAJ_OK) team_events_start<E>(a, idx, ts[L - 1], ys[L - 1], pp, ev_last, evbuf, tid == 0, i0, n, cta_barrier{});
        }
#endif
        int flip = 0;
        if constexpr (!TR::ERK && (L > 2) && !EV) {
            // AdamsBashforth3..5: the window is a ring in registers like in the per-thread kernel (advance_ab) -- L
            // steps unrolled, a push overwrites the oldest slot (the shifting loop below moved (L-1)(2E+1) doubles per
            // step), a trajectory that stops inside the block skips the rest of it.  Control flow is the same in every
            // thread of the CTA, so the barriers inside cta_eval are safe.
            constexpr int ORDER = METHOD;
            int nst = 0, rot = 0;
            const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;
            const int n_steps = a.n_steps > 2147483647LL ? 2147483647 : (int)a.n_steps;
            while (running) {
#pragma unroll
                for (int r = 0; r < L; ++r) {
                    const int newest = (L - 1 + r) % L, oldest = r % L;
                    const double t_new = ts[newest] + h;
                    if (running && t_new > a.bound) {
                        status = NBK_TRAJ_BOUND;
                        running = false;
                    }
                    if (running) {
                        double ynew[E], fnew[E];
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            double acc = (h * tab_ab<ORDER>::b(0)) * fs[(L - ORDER + r) % L][e];
#pragma unroll
                            for (int j = 1; j < ORDER; ++j) acc = fma(h * tab_ab<ORDER>::b(j), fs[(L - ORDER + j + r) % L][e], acc);
                            ynew[e] = acc + ys[newest][e];
                        }
                        cta_eval<Rhs, E>(t_new, ynew, p, i0, n, (flip ^= 1) ? buf1 : buf0, fnew);
                        ts[oldest] = t_new;
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            ys[oldest][e] = ynew[e];
                            fs[oldest][e] = fnew[e];
                        }
                        ++nst;
                        rot = (r + 1) % L;
                        if (MODE == NBK_MODE_RUN) {
                            while (t_new >= te_next) {
                                double out[E];
                                cta_hermite_ends<E>(te_next, ts[(r + 1) % L], ys[(r + 1) % L], fs[(r + 1) % L], t_new, ynew, fnew, out);
#pragma unroll
                                for (int e = 0; e < E; ++e)
                                    if (i0 + e < n) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                                ++kout;
                                te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                            }
                            running = kout < a.m;
                        } else {
                            if (a.emit) {
                                if (tid == 0) a.t_out[idx * a.to_sb + (nst - 1)] = t_new;
#pragma unroll
                                for (int e = 0; e < E; ++e)
                                    if (i0 + e < n) a.y_out[idx * a.o_sb + (nst - 1) * a.o_sk + (long long)(i0 + e) * a.o_sc] = ynew[e];
                            }
                            running = nst < n_steps;
                        }
                        if (running && nst >= max_att) {
                            status = NBK_TRAJ_MAXATTEMPTS;
                            running = false;
                        }
                    }
                }
            }
            if (nst > 0) {
                switch (rot) {
                    case 0: cta_store_window<L, E, 0>(a, idx, n, i0, ts, ys, fs); break;
                    case 1: cta_store_window<L, E, 1 % L>(a, idx, n, i0, ts, ys, fs); break;
                    case 2: cta_store_window<L, E, 2 % L>(a, idx, n, i0, ts, ys, fs); break;
                    case 3: cta_store_window<L, E, 3 % L>(a, idx, n, i0, ts, ys, fs); break;
                    default: cta_store_window<L, E, 4 % L>(a, idx, n, i0, ts, ys, fs); break;
                }
            }
            if (tid == 0) {
                if (nst > 0) a.n_acc[idx] += nst;
                a.status[idx] = status;
                if (status >= NBK_TRAJ_NONFINITE) atomicAdd(a.work_counter + 1, 1ULL);
                a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : nst;
            }
            continue;  // next trajectory of this CTA
        }
        while (running) {  // identical control flow in every thread of the CTA
            if (ts[L - 1] + h > a.bound) {
                status = NBK_TRAJ_BOUND;
                break;
            }
            const double t = ts[L - 1], t_new = t + h;
            double ynew[E], fnew[E];
            if constexpr (TR::ERK) {
                using Tab = tab_erk<TR::ERK ? METHOD : 204>;
#pragma unroll
                for (int e = 0; e < E; ++e) K[0][e] = fs[L - 1][e];
#pragma unroll
                for (int s = 1; s < S; ++s) {
                    double ystage[E];
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        double acc = 0.0;
                        bool first = true;
#pragma unroll
                        for (int j = 0; j < s; ++j) {
                            if (Tab::a(s, j) != 0.0) {
                                acc = first ? Tab::a(s, j) * K[j][e] : fma(Tab::a(s, j), K[j][e], acc);
                                first = false;
                            }
                        }
                        ystage[e] = fma(h, acc, ys[L - 1][e]);
                    }
                    cta_eval<Rhs, E>(fma(h, Tab::c(s), t), ystage, p, i0, n, (flip ^= 1) ? buf1 : buf0, K[s]);
                }
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    double acc = 0.0;
                    bool first = true;
#pragma unroll
                    for (int j = 0; j < S; ++j) {
                        if (Tab::b(j) != 0.0) {
                            acc = first ? (h * Tab::b(j)) * K[j][e] : fma(h * Tab::b(j), K[j][e], acc);
                            first = false;
                        }
                    }
                    ynew[e] = ys[L - 1][e] + acc;
                }
            } else {
                constexpr int ORDER = TR::ERK ? 1 : METHOD;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    double acc = (h * tab_ab<ORDER>::b(0)) * fs[L - ORDER][e];
#pragma unroll
                    for (int j = 1; j < ORDER; ++j) acc = fma(h * tab_ab<ORDER>::b(j), fs[L - ORDER + j][e], acc);
                    ynew[e] = acc + ys[L - 1][e];
                }
            }
            cta_eval<Rhs, E>(t_new, ynew, p, i0, n, (flip ^= 1) ? buf1 : buf0, fnew);
#pragma unroll
            for (int l = 0; l < L - 1; ++l) {
                ts[l] = ts[l + 1];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ys[l][e] = ys[l + 1][e];
                    fs[l][e] = fs[l + 1][e];
                }
            }
            ts[L - 1] = t_new;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                ys[L - 1][e] = ynew[e];
                fs[L - 1][e] = fnew[e];
            }
            ++nsteps;
#ifdef NBK_NEVENTS
            if constexpr (EV && MODE == NBK_MODE_RUN) {
                // events first, then the grid points inside the step (core.py:638-645); every quantity that steers this
                // block is identical in all threads of the team
                cta_window_dense<L, E> dense(ts, ys, fs);
                double t_term = 0.0;
                if (team_events_after_step<E>(a, idx, ts[L - 2], t_new, ynew, pp, ev_last, evbuf, tid == 0, i0, n, cta_barrier{}, dense,
                                              t_term, status)) {
                    if (status != NBK_TRAJ_EVBRACKET) {
                        double out[E];
                        dense(t_term, out);
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            if (i0 + e < n) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                        if (tid == 0) a.ev_tterm[idx] = t_term;
                        ++kout;
                        status = NBK_TRAJ_EVENT;
                    }
                    break;
                }
            }
#endif
            if (MODE == NBK_MODE_RUN) {
                while (t_new >= te_next) {
                    double out[E];
                    cta_hermite<L, E>(te_next, ts, ys, fs, out);
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (i0 + e < n) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                    ++kout;
                    te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                }
                running = kout < a.m;
            } else {
                if (a.emit) {
                    if (tid == 0) a.t_out[idx * a.to_sb + (nsteps - 1)] = t_new;
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (i0 + e < n) a.y_out[idx * a.o_sb + (nsteps - 1) * a.o_sk + (long long)(i0 + e) * a.o_sc] = ynew[e];
                }
                running = nsteps < a.n_steps;
            }
            if (running && nsteps >= a.max_attempts) {
                status = NBK_TRAJ_MAXATTEMPTS;
                break;
            }
        }
        if (nsteps > 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = i0 + e;
                if (i < n) {
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        a.ys[(idx * L + l) * nn + i] = ys[l][e];
                        a.fs[(idx * L + l) * nn + i] = fs[l][e];
                    }
#pragma unroll
                    for (int s = 0; s < S; ++s) a.K[(idx * S + s) * nn + i] = K[s][e];
                }
            }
        }
        if (tid == 0) {
            if (nsteps > 0) {
#pragma unroll
                for (int l = 0; l < L; ++l) a.ts[idx * L + l] = ts[l];
                a.n_acc[idx] += nsteps;
            }
            a.status[idx] = status;
            if (status >= NBK_TRAJ_NONFINITE) atomicAdd(a.work_counter + 1, 1ULL);
            a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : (int)nsteps;
        }
    }
}

}  // namespace nbk
#endif
