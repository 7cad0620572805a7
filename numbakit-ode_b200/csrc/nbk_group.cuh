// nbk_group.cuh -- sub-warp regime for mid-size systems (n ~ 13..64): G lanes of a warp (G = 4, 8, 16 or 32)
// integrate one trajectory, 32 / G trajectories per warp.
//
// Between the two regimes of the north star there is a band where neither fits: one trajectory per thread needs
// 7n + 2n doubles of registers (spills from n ~ 13 on), and one CTA (>= one warp) per trajectory pays the whole
// controller -- error norm, step-size factor, accept / reject, guards, ~100 instructions -- per warp for 32 or fewer
// components.  Here a lane owns E = ceil(n / G) consecutive components, so a warp's controller instructions are shared by
// 32 / G trajectories x E components per lane:
//  * y, K0..K_S, y_new of a lane's slice live in registers (7E + 2E doubles: 36 for E = 4);
//  * the stage ARGUMENT vector is staged in the group's own 2 x (G E) doubles of shared memory (same transposed layout
//    and the same component-wise right-hand-side interface as the CTA regime, nbk_cta.cuh: built-ins, NVRTC nbk_rhs_i,
//    traced Python callables); the barrier is a __syncwarp of the group's lanes;
//  * the error norm is a log2(G)-round shuffle butterfly inside the group: every lane of the group gets the same sum,
//    hence the same accept / reject decision and step-size factor -- a group never diverges internally;
//  * groups of one warp differ in their decisions: like the per-thread kernel (advance_adaptive) every loop iteration is
//    one ATTEMPT for every running group, accept / reject is committed with selects, and a group that finishes takes the
//    next trajectory from the global queue without waiting for its neighbours;
//  * state layout = CTA regime (trajectory-major: ts [B][2], ys/fs [B][2][n], K [B][S+1][n], params [B][np]); the
//    construction kernel is the CTA regime's (cta_init_adaptive, one small block per trajectory).
//
// Reference rows restated: as nbk_cta.cuh (FSAL stage loop explicit.py:60-79, error norm / controller
// runge_kutta/core.py:73-99, accept loop explicit.py:81-99, t_bound guard core.py:176-184, RkInterpolator
// explicit.py:354-403, drivers core.py:483-537, 598-619).
#ifndef NBK_GROUP_CUH
#define NBK_GROUP_CUH

#include "nbk_cta.cuh"

#ifndef NBK_GRP_INNER
#define NBK_GRP_INNER 4  // attempts between two looks at the queue / retire code
#endif

namespace nbk {

// (group_sync / group_shfl_xor / group_bcast / warp_all: nbk_cta.cuh; tests/emul/nbk_emul_group.cpp on the host)

// Sum over the G lanes of a group; the xor butterfly gives every lane the same value (each round adds the same pair
// of numbers on both sides).
template <int G>
NBK_DEV double group_sum(double v, unsigned gmask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += group_shfl_xor(gmask, v, o);
    return v;
}

// Element-wise decay y' = -k y for the component-wise regimes (k scalar or per component): needs no neighbour, so the
// stage argument is not staged at all (LOCAL).
template <bool SCALAR>
struct rhs_decay_cta {
    static constexpr int NP = SCALAR ? 1 : -1;  // one rate for all components, or n of them (run-time)
    static constexpr bool LOCAL = true;
    struct prm { const double *p; };
    static NBK_DEV prm load(const double *p) { return prm{p}; }
    template <int E, bool FULL = false>
    static NBK_DEV void eval(double, const double (&yown)[E], smem_vec<E>, const prm &q, int i0, int n, double (&dy)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e < n ? i0 + e : n - 1;
            dy[e] = -q.p[SCALAR ? 0 : i] * yown[e];
        }
    }
    // sub-warp regime: a lane's rates are fetched once per trajectory and stay in registers (see group_prm)
    template <int E>
    struct lane_prm { double k[E]; };
    template <int E>
    static NBK_DEV lane_prm<E> load_lane(const double *p, int i0, int n) {
        lane_prm<E> q;
#pragma unroll
        for (int e = 0; e < E; ++e) q.k[e] = p[SCALAR ? 0 : (i0 + e < n ? i0 + e : n - 1)];
        return q;
    }
    template <int E, bool FULL = false>
    static NBK_DEV void eval(double, const double (&yown)[E], smem_vec<E>, const lane_prm<E> &q, int, int, double (&dy)[E]) {
#pragma unroll
        for (int e = 0; e < E; ++e) dy[e] = -q.k[e] * yown[e];
    }
};

// Parameter block of a group's lane: the CTA regime's Rhs::prm, or -- when the right-hand side offers it -- a per-lane
// block Rhs::lane_prm<E> loaded by Rhs::load_lane<E>(p, i0, n) (parameters that belong to the lane's own components).
template <class Rhs, int E, class = void>
struct group_prm {
    using type = typename Rhs::prm;
    static NBK_DEV type load(const double *p, int, int) { return Rhs::load(p); }
};
template <class Rhs, int E>
struct group_prm<Rhs, E, decltype((void)sizeof(typename Rhs::template lane_prm<E>))> {
    using type = typename Rhs::template lane_prm<E>;
    static NBK_DEV type load(const double *p, int i0, int n) { return Rhs::template load_lane<E>(p, i0, n); }
};

template <class Rhs, class = void>
struct rhs_is_local { static constexpr bool value = false; };
template <class Rhs>
struct rhs_is_local<Rhs, decltype((void)Rhs::LOCAL)> { static constexpr bool value = Rhs::LOCAL; };

// One FSAL attempt of a group.  K[0] holds the lane's slice of f(t, y).
template <class Tab, class Rhs, int G, int E, bool FULL, class P>
NBK_DEV void group_attempt(double t, double h, const double (&y)[E], double (&K)[Tab::S + 1][E], const P &p,
                           int gl, int i0, int n, double inv_n, double *buf0, double *buf1, unsigned gmask,
                           double (&ynew)[E], double &err2, double rtol, double atol) {
    constexpr int S = Tab::S;
    constexpr bool LOCAL = rhs_is_local<Rhs>::value;
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
        if constexpr (!LOCAL) {
#pragma unroll
            for (int e = 0; e < E; ++e) buf[e * G + gl] = ys[e];
            group_sync(gmask);
        }
        Rhs::template eval<E, false>(s < S ? fma(h, Tab::cb().c[s < S ? s : 0], t) : t + h, ys, smem_vec<E>{buf, G}, p, i0, n,
                                     K[s]);
        if (s == S) {
#pragma unroll
            for (int e = 0; e < E; ++e) ynew[e] = ys[e];
        }
    }
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
        part = cta_inb<FULL>(i0 + e, n) ? fma(r, r, part) : part;  // a select, not a branch (nothing at all in a full group)
    }
    err2 = group_sum<G>(part, gmask) * (h * h * inv_n);
    // closes the window between this attempt's last reads of a stage buffer and the next attempt's first writes (for
    // RungeKutta23 the last stage and the next first stage use the same buffer)
    if constexpr (!LOCAL) group_sync(gmask);
}

#ifdef NBK_NEVENTS
// run_events in the sub-warp regime: the team routine of nbk_cta.cuh (team_events_after_step) with the group's barrier and
// the RkInterpolator of the lane's slice.
struct group_barrier {
    unsigned gmask;
    NBK_DEV void operator()() const { group_sync(gmask); }
};
template <class Tab, int E>
struct group_step_dense {
    double t_old, t_new;
    const double (&y_old)[E];
    const double (&y_new)[E];
    const double (&K)[Tab::S + 1][E];
    double Q[E][Tab::M];
    bool have_q;
    NBK_DEV group_step_dense(double t0, double t1, const double (&y0)[E], const double (&y1)[E], const double (&K_)[Tab::S + 1][E])
        : t_old(t0), t_new(t1), y_old(y0), y_new(y1), K(K_), have_q(false) {}
    NBK_DEV void operator()(double te, double (&out)[E]) {
        if (!have_q) {
            rk_dense_coeffs<Tab, E>(K, Q);
            have_q = true;
        }
        rk_dense_eval<Tab, E>(te, t_old, t_new, y_old, y_new, Q, out);
    }
};
#endif  // NBK_NEVENTS

// =====================================================================================
// Advance: persistent warps, every group pulls trajectories from the global queue.
// MODE RUN / STEP as in advance_adaptive (nbk_device.cuh) and cta_advance (nbk_cta.cuh).
// =====================================================================================
template <class Tab, class Rhs, int G, int E, int MODE, bool EV = false, bool FULL = false>
NBK_DEV void group_advance(const nbk_kargs &a, double *sm) {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
    static_assert(!Tab::DOP, "the sub-warp regime implements RungeKutta23 / RungeKutta45");
    constexpr int S = Tab::S;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gbase = lane - gl;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const int n = a.n, i0 = gl * E;
    const long long nn = n, B = a.B;
    const double inv_n = 1.0 / (double)n;
    double *buf0 = sm + (size_t)(threadIdx.x / G) * (2 * G * E), *buf1 = buf0 + G * E;
#ifdef NBK_NEVENTS
    // EV: the whole-vector buffer of the group's event functions sits behind the stage buffers of the block
    double *evbuf = sm + (size_t)(NBK_GRP_BLOCK / G) * (2 * G * E) + (size_t)(threadIdx.x / G) * (G * E);
    double ev_last[NBK_NEVENTS];
    const double *pp = a.params;
#endif
    const double rtol = a.rtol, atol = a.atol, bound = a.bound;
    const bool bounded = (__double2hiint(a.bound) & 0x7fffffff) < 0x7ff00000;  // finite (an integer test: stays off the FP64 pipe)  // (usually not: skips the guard of every accepted step)
    const double safety = a.safety, min_factor = a.min_factor, max_factor = a.max_factor;
    const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;

    // loop-carried state of the group's trajectory; the scalars are replicated in every lane of the group
    long long idx = -1;
    bool running = false, finishing = false, queue_empty = false, fresh = true;
    double t = 0, h = 0, te_next = 0;
    double y[E], K[S + 1][E];
    using PRM = group_prm<Rhs, E>;
    typename PRM::type p{};
    int nacc = 0, natt = 0, kout = 0, status = NBK_TRAJ_OK;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        y[e] = 0.0;
#pragma unroll
        for (int s = 0; s <= S; ++s) K[s][e] = 0.0;
    }

    for (;;) {
        // ---------------------------------------------------------------- (1) retire
        if (finishing) {
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
            if (gl == 0) {
                if (nacc > 0) a.ts[idx * 2 + 1] = t;
                a.h[idx] = h;
                if (a.pristine) {
                    a.n_acc[idx] = nacc;
                    a.n_rej[idx] = natt - nacc;
                } else {
                    a.n_acc[idx] += nacc;
                    a.n_rej[idx] += natt - nacc;
                }
                store_status(a, idx, status);
                a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : nacc;
            }
            finishing = false;
            idx = -1;
        }
        // ---------------------------------------------------------------- (2) refill (group-uniform branch)
        if (!running && !queue_empty) {
            long long mine = 0;
            if (gl == 0) mine = (long long)atomicAdd(a.work_counter, 1ULL);
            mine = group_bcast(gmask, mine, gbase);
            if (mine >= B) {
                queue_empty = true;
            } else {
                idx = mine;
                p = PRM::load(a.params_shared ? a.params : a.params + idx * a.n_params, i0, n);
                t = a.ts[idx * 2 + 1];
                h = a.h[idx];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i = i0 + e;
                    y[e] = cta_inb<FULL>(i, n) ? a.ys[(idx * 2 + 1) * nn + i] : 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) K[s][e] = 0.0;
                    K[S][e] = cta_inb<FULL>(i, n) ? a.fs[(idx * 2 + 1) * nn + i] : 0.0;  // f(t, y): K[0] of the first attempt
                }
                status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
                nacc = natt = kout = 0;
                fresh = true;
                te_next = d_inf();
                running = status == NBK_TRAJ_OK;
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
                        double Q[E][Tab::M];
                        rk_dense_coeffs<Tab, E>(Kh, Q);
                        while (kout < a.m && __ldg(a.tgrid + kout) <= t) {
                            double out[E];
                            rk_dense_eval<Tab, E>(__ldg(a.tgrid + kout), t_old, t, y_old, y, Q, out);
#pragma unroll
                            for (int e = 0; e < E; ++e)
                                if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                            ++kout;
                        }
                    }
                    running = kout < a.m;
                    if (running) te_next = __ldg(a.tgrid + kout);
#ifdef NBK_NEVENTS
                    if constexpr (EV) {
                        // build_handler(events, self.t, self.y): last_value = func(t, y), empty lists (event_handler.py:127-135)
                        pp = a.params_shared ? a.params : a.params + idx * a.n_params;
                        team_events_start<E>(a, idx, t, y, pp, ev_last, evbuf, gl == 0, i0, n, group_barrier{gmask});
                    }
#endif
                } else if (running) {
                    running = a.n_steps > 0;
                }
                if (status >= NBK_TRAJ_NONFINITE && MODE == NBK_MODE_RUN) {
                    for (int k = 0; k < a.m; ++k)
                        for (int e = 0; e < E; ++e)
                            if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + k * a.o_sk + (long long)(i0 + e) * a.o_sc] = d_nan();
                }
                if (running && (t + h > bound)) {  // guard of the first step (core.py:176-184)
                    status = NBK_TRAJ_BOUND;
                    running = false;
                }
                finishing = !running;
            }
        }
        // the warp leaves when the queue is empty and none of its groups has anything left
        if (warp_all(!running && !finishing && queue_empty)) break;
        // ---------------------------------------------------------------- (3) attempts
#pragma unroll 1
        for (int rep = 0; rep < NBK_GRP_INNER; ++rep) {
            if (running) {
                if (fresh) {  // FSAL
#pragma unroll
                    for (int e = 0; e < E; ++e) K[0][e] = K[S][e];
                }
                double ynew[E], err2;
                group_attempt<Tab, Rhs, G, E, FULL>(t, h, y, K, p, gl, i0, n, inv_n, buf0, buf1, gmask, ynew, err2, rtol, atol);
                const bool accept = (unsigned)__double2hiint(err2) < 0x3ff00000u;
                const double t_new = t + h;
                h *= step_factor<Tab>(err2, safety, min_factor, max_factor, a.x_lo, a.x_hi);
                ++natt;
                bool done = false;
                if (accept) {
                    bool stopped_by_event = false;
#ifdef NBK_NEVENTS
                    if constexpr (MODE == NBK_MODE_RUN && EV) {
                        // events first, then the grid points inside the step (core.py:638-645)
                        double t_term = 0.0;
                        group_step_dense<Tab, E> dense(t, t_new, y, ynew, K);
                        if (team_events_after_step<E>(a, idx, t, t_new, ynew, pp, ev_last, evbuf, gl == 0, i0, n, group_barrier{gmask},
                                                      dense, t_term, status)) {
                            if (status != NBK_TRAJ_EVBRACKET) {
                                double out[E];
                                dense(t_term, out);
#pragma unroll
                                for (int e = 0; e < E; ++e)
                                    if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                                if (gl == 0) a.ev_tterm[idx] = t_term;
                                ++kout;
                                status = NBK_TRAJ_EVENT;
                            }
                            done = true;
                            stopped_by_event = true;
                        }
                    }
#endif
                    if (stopped_by_event) {
                    } else if (MODE == NBK_MODE_RUN) {
                        if (t_new >= te_next) {
                            double Q[E][Tab::M];
                            rk_dense_coeffs<Tab, E>(K, Q);
                            const double H = t_new - t, r0 = fast_rcp(H);
                            const double inv_H = fma(r0, fma(-H, r0, 1.0), r0);
                            do {
                                double out[E];
                                rk_dense_eval_in_step<Tab, E>(te_next, t, t_new, y, ynew, Q, out, inv_H);
#pragma unroll
                                for (int e = 0; e < E; ++e)
                                    if (cta_inb<FULL>(i0 + e, n)) a.y_out[idx * a.o_sb + kout * a.o_sk + (long long)(i0 + e) * a.o_sc] = out[e];
                                ++kout;
                                te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                            } while (t_new >= te_next);
                            done = kout >= a.m;
                        }
                    } else {
                        if (a.emit) {
                            if (gl == 0) a.t_out[idx * a.to_sb + nacc] = t_new;
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
                    if (gl == 0) a.ts[idx * 2] = t;
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (cta_inb<FULL>(i0 + e, n)) a.ys[(idx * 2) * nn + i0 + e] = y[e];
                    running = false;
                    finishing = true;
                }
                if (accept) {  // commit (predicated moves)
                    t = t_new;
#pragma unroll
                    for (int e = 0; e < E; ++e) y[e] = ynew[e];
                    ++nacc;
                }
                fresh = accept;
            }
        }
    }
}

}  // namespace nbk

#endif
