// nbk_mma.cuh -- K4: adaptive Runge-Kutta for LINEAR systems y' = A y with ONE dense matrix A
// shared by the whole ensemble (BASELINE.json configs[3]: n = 128, 262 144 trajectories).
//
// With a shared A the right-hand side of a TILE of trajectories is a dense GEMM,
//      K_s[n x NT] = A[n x n] . Ystage_s[n x NT],
// which runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; Blackwell's tcgen05
// has no FP64 kind).  Measured DMMA peak on B200: 36.9 TFLOP/s, the same roof as the vector pipe
// (34.1), so the gain is not a higher ceiling but 8x fewer instructions per FLOP and operands that
// stay in registers -- the generic CTA kernel (one DFMA + one global + one shared load per
// multiply-add) reaches 15 % of peak on this configuration.
//
// One CTA of 8 warps integrates a tile of NT = 16 trajectories in lock-step ATTEMPTS; every
// trajectory (column) keeps its own t, h, accept/reject decision and output cursor, so results are
// identical to integrating each trajectory alone.
//   * y, y_new and the stage derivatives K0..K_S of the tile live in registers in the DMMA
//     accumulator layout: warp w owns rows [16w, 16w+16) of all 16 columns = 4 accumulator tiles,
//     i.e. 8 doubles per thread and vector;
//   * the stage argument Y_s is staged in shared memory (two buffers, padded rows) as the B operand;
//   * A is read as A-fragments from its transpose in global memory (L1/L2 resident: 128 KB);
//   * the error norm of a column is a shuffle reduction over the lanes that share the column plus
//     one shared-memory hop over the 8 warps; 16 "column leader" threads run the controller.
// State layout: trajectory-major, exactly the CTA regime's (nbk_cta.cuh), so the same
// construction kernel is used.
#ifndef NBK_MMA_CUH
#define NBK_MMA_CUH

#include "nbk_cta.cuh"

namespace nbk {

// Shape of the tile kernel.  NBK_MMA_GROUPS = 1 (shipped): one group of 8 warps, tiles of 16 trajectories (a warp owns 16
// rows x 16 columns = 4 accumulator tiles), everything between two CTA-wide barriers in lock-step.  NBK_MMA_GROUPS = 2: the
// CTA's 8 warps form TWO independent groups of 4; each group integrates its own tiles of 8 trajectories (a warp owns 32
// rows x 8 columns) behind its own named barrier, so that one group's element-wise stage algebra, error norm and serial
// controller section overlap the other group's DMMA stream -- both share the one copy of the matrix in shared memory.
// Measured on C4 (262 144 x n = 128): one group 110.4 ms, two groups 113.2 ms; 16 warps of 8 rows 112.8 ms; fragments
// fetched as 16-byte pairs 113.1-115.3 ms -- every shape lands within 4 %: with 8 warps per SM a DMMA stream fed from
// shared memory tops out at 27 of the 37 TFLOP/s the pipe delivers from registers (tools/micro/dmma_occ.cu), and the
// kernel runs at 24.4 including everything that is not the GEMM.
#ifndef NBK_MMA_GROUPS
#define NBK_MMA_GROUPS 1
#endif
#ifndef NBK_MMA_MT
#define NBK_MMA_MT (NBK_MMA_GROUPS == 2 ? 4 : 2)  // 8-row accumulator tiles per warp
#endif
#ifndef NBK_MMA_KSPLIT
#define NBK_MMA_KSPLIT (NBK_MMA_GROUPS == 2 ? 2 : 4)  // interleaved partial sums per accumulator tile (independent DMMA chains)
#endif
constexpr int MMA_GROUPS = NBK_MMA_GROUPS;
constexpr int MMA_KSPLIT = NBK_MMA_KSPLIT;
constexpr int MMA_NTN = MMA_GROUPS == 2 ? 1 : 2;  // 8-column accumulator tiles across a tile of trajectories
constexpr int MMA_NT = 8 * MMA_NTN;       // trajectories per tile
constexpr int MMA_MT = NBK_MMA_MT;
constexpr int MMA_ROWS = 8 * MMA_MT;      // rows of the tile a warp owns: warp w owns rows [MMA_ROWS w, MMA_ROWS (w + 1))
constexpr int MMA_WARPS = 128 / MMA_ROWS; // warps per group
constexpr int MMA_GTHREADS = 32 * MMA_WARPS;            // threads per group
constexpr int MMA_THREADS = MMA_GTHREADS * MMA_GROUPS;  // threads per CTA
constexpr int MMA_LD = MMA_NT + 4;  // padded row length of the staged Y (doubles)
constexpr int MMA_LDA = 128 + 4;    // padded row length of the matrix staged in shared memory: a fragment load touches
                                    // (k, row) = (k0 + lane % 4, r0 + lane / 4); LDA = 4 mod 16 puts the 16 lanes of a
                                    // half-warp into 16 different 8-byte banks

NBK_DEV void dmma_8x8x4(double &c0, double &c1, double a, double b) {
#ifdef NBK_HOST_EMUL
    nbk_emul_dmma(c0, c1, a, b);  // tests/emul: the fragment exchange of one warp through a scratch array
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

// The group's barrier: the whole CTA when there is one group, a named barrier of the group's threads otherwise (ids 1, 2;
// id 0 stays the CTA-wide one).  mma_bar_or also ORs a predicate over the group (bar.red).
NBK_DEV void mma_bar(int grp) {
#ifdef NBK_HOST_EMUL
    nbk_emul_bar(grp);
#else
    if (MMA_GROUPS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(MMA_GTHREADS) : "memory");
#endif
}
NBK_DEV int mma_bar_or(int grp, int pred) {
#ifdef NBK_HOST_EMUL
    return nbk_emul_bar_or(grp, pred);
#else
    if (MMA_GROUPS == 1) return __syncthreads_or(pred);
    unsigned r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r)
                 : "r"((unsigned)pred), "r"(grp + 1), "r"(MMA_GTHREADS)
                 : "memory");
    return (int)r;
#endif
}

// per-column bookkeeping kept in shared memory (written by the column leaders only)
struct mma_cols {
    double t[MMA_NT], h[MMA_NT], t_new[MMA_NT], te_next[MMA_NT], err2[MMA_NT];
    long long idx[MMA_NT];
    int nacc[MMA_NT], natt[MMA_NT], kout[MMA_NT], status[MMA_NT];
    int running[MMA_NT], fresh[MMA_NT], accept[MMA_NT], done[MMA_NT], emit_from[MMA_NT];
    // what the element threads need of the attempt that was just decided, after the leaders have already moved the
    // column on (one barrier per attempt tail instead of four): was the column active, its start time and step count
    double t_prev[MMA_NT];
    int act[MMA_NT], nacc_prev[MMA_NT];
    long long tile;
};

// Element (i, c) of a thread's values v[mt][nt][q]: row = MMA_ROWS w + 8 mt + lane/4, column = 8 nt + 2 (lane%4) + q
#define NBK_MMA_FOR_ELEMS(...)                                                    \
    _Pragma("unroll") for (int mt = 0; mt < MMA_MT; ++mt)                         \
    _Pragma("unroll") for (int nt = 0; nt < MMA_NTN; ++nt)                        \
    _Pragma("unroll") for (int q = 0; q < 2; ++q) {                               \
        const int row = MMA_ROWS * warp + 8 * mt + (lane >> 2);                   \
        const int col = 8 * nt + 2 * (lane & 3) + q;                              \
        (void)row; (void)col;                                                     \
        __VA_ARGS__                                                               \
    }

template <class Tab, int NPAD, int MODE>
NBK_DEV void mma_advance(const nbk_kargs &a, double *sm) {
    constexpr int S = Tab::S;
    static_assert(NPAD == MMA_ROWS * MMA_WARPS, "the warps' row blocks cover the padded matrix");
    // tid / warp are relative to the thread's GROUP from here on: leaders, row blocks and the reduction scratch are per group
    const int n = a.n, grp = (int)threadIdx.x / MMA_GTHREADS, tid = (int)threadIdx.x - grp * MMA_GTHREADS, warp = tid >> 5, lane = tid & 31;
    const long long nn = n;
    constexpr int COLS_WORDS = (int)((sizeof(mma_cols) + 7) / 8);
    constexpr int GROUP_WORDS = 2 * NPAD * MMA_LD + MMA_WARPS * MMA_NT + COLS_WORDS + 2;  // a group's slice of shared memory (even)
    double *gsm = sm + (size_t)grp * GROUP_WORDS;
    double *ybuf0 = gsm, *ybuf1 = gsm + NPAD * MMA_LD;
    double *red = gsm + 2 * NPAD * MMA_LD;                     // [MMA_WARPS][MMA_NT]
    mma_cols *cs = (mma_cols *)(red + MMA_WARPS * MMA_NT);
    const double *At = a.params;  // A transposed: At[k*n + i] = A[i][k]
    // The matrix lives in shared memory for the whole (persistent) kernel -- one copy for all groups --, zero-padded to
    // NPAD x NPAD: fragment loads need no bounds checks, and a shared-memory load of 8 rows x 4 k costs 2 wavefronts where
    // the same load from global memory touched 4 cache lines (the L1 data path was 83 % busy, ahead of the DMMA pipe at 61 %).
    double *As = (double *)(((unsigned long long)(sm + (size_t)MMA_GROUPS * GROUP_WORDS) + 15ULL) & ~15ULL);
    for (int i = (int)threadIdx.x; i < NPAD * NPAD; i += (int)blockDim.x) {
        const int k = i / NPAD, r = i % NPAD;
        As[k * MMA_LDA + r] = (k < n && r < n) ? At[(long long)k * n + r] : 0.0;
    }
    __syncthreads();  // the last CTA-wide barrier: from here on the groups run on their own
    const double rtol = a.rtol, atol = a.atol, bound = a.bound;
    const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;
    const long long n_tiles = (a.B + MMA_NT - 1) / MMA_NT;

    for (;;) {
        mma_bar(grp);
        if (tid == 0) cs->tile = (long long)atomicAdd(a.work_counter, 1ULL);
        mma_bar(grp);
        const long long tile = cs->tile;
        if (tile >= n_tiles) break;

        // ---------------------------------------------------------------- load the tile
        if (tid < MMA_NT) {
            const int c = tid;
            const long long idx = tile * MMA_NT + c;
            cs->idx[c] = idx < a.B ? idx : -1;
            cs->nacc[c] = cs->natt[c] = cs->kout[c] = 0;
            cs->fresh[c] = 1;
            cs->te_next[c] = d_inf();
            if (idx < a.B) {
                cs->t[c] = a.ts[idx * 2 + 1];
                cs->h[c] = a.h[idx];
                const int st = a.status[idx];
                cs->status[c] = st >= NBK_TRAJ_NONFINITE ? st : NBK_TRAJ_OK;
                int run = cs->status[c] == NBK_TRAJ_OK;
                if (run) {
                    if (MODE == NBK_MODE_RUN) {
                        int k = 0;
                        while (k < a.m && __ldg(a.tgrid + k) <= cs->t[c]) ++k;
                        cs->emit_from[c] = k;  // [0, k) lie in the stored last step: emitted below from history
                        cs->kout[c] = k;
                        run = k < a.m;
                        if (run) cs->te_next[c] = __ldg(a.tgrid + k);
                    } else {
                        cs->emit_from[c] = 0;
                        run = a.n_steps > 0;
                    }
                    if (run && (cs->t[c] + cs->h[c] > bound)) {
                        cs->status[c] = NBK_TRAJ_BOUND;
                        run = 0;
                    }
                } else {
                    cs->emit_from[c] = 0;
                }
                cs->running[c] = run;
            } else {
                cs->t[c] = cs->h[c] = 0.0;
                cs->status[c] = NBK_TRAJ_OK;
                cs->running[c] = 0;
                cs->emit_from[c] = 0;
            }
        }
        mma_bar(grp);
        double y[MMA_MT][MMA_NTN][2], K[S + 1][MMA_MT][MMA_NTN][2];
        NBK_MMA_FOR_ELEMS({
            const long long idx = cs->idx[col];
            const bool ok = idx >= 0 && row < n;
            y[mt][nt][q] = ok ? a.ys[(idx * 2 + 1) * nn + row] : 0.0;
_Pragma("unroll")
            for (int s = 0; s < S; ++s) K[s][mt][nt][q] = 0.0;
            K[S][mt][nt][q] = ok ? a.fs[(idx * 2 + 1) * nn + row] : 0.0;
            if (MODE == NBK_MODE_RUN && ok && cs->emit_from[col] > 0) {
                // grid points inside the stored last step are answered from history (core.py:401-408)
                double yo[1], yn[1] = {y[mt][nt][q]}, Ke[S + 1][1], out[1];
                const double t_old = a.pristine ? cs->t[col] : a.ts[idx * 2];
                yo[0] = a.pristine ? yn[0] : a.ys[(idx * 2) * nn + row];
_Pragma("unroll")
                for (int s = 0; s <= S; ++s) Ke[s][0] = a.pristine ? 0.0 : a.K[(idx * (S + 1) + s) * nn + row];
                for (int k = 0; k < cs->emit_from[col]; ++k) {
                    rk_dense<Tab, 1>(__ldg(a.tgrid + k), t_old, cs->t[col], yo, yn, Ke, out);
                    a.y_out[idx * a.o_sb + k * a.o_sk + (long long)row * a.o_sc] = out[0];
                }
            }
            if (MODE == NBK_MODE_RUN && ok && cs->status[col] >= NBK_TRAJ_NONFINITE) {
                for (int k = 0; k < a.m; ++k) a.y_out[idx * a.o_sb + k * a.o_sk + (long long)row * a.o_sc] = d_nan();
            }
        })
        // (the barrier doubles as the OR over the columns' running flags)
        int any_running = mma_bar_or(grp, tid < MMA_NT ? cs->running[tid] : 0);

        // ---------------------------------------------------------------- lock-step attempts
        while (any_running) {
            // columns that already finished keep their registers frozen until the tile retires
            double hcol[MMA_NTN][2], ynew[MMA_MT][MMA_NTN][2];
            bool runc[MMA_NTN][2];
_Pragma("unroll")
            for (int nt = 0; nt < MMA_NTN; ++nt)
_Pragma("unroll")
                for (int q = 0; q < 2; ++q) {
                    hcol[nt][q] = cs->h[8 * nt + 2 * (lane & 3) + q];
                    runc[nt][q] = cs->running[8 * nt + 2 * (lane & 3) + q] != 0;
                }
            // FSAL per column
            NBK_MMA_FOR_ELEMS({
                if (cs->fresh[col] && runc[nt][q]) K[0][mt][nt][q] = K[S][mt][nt][q];
            })
_Pragma("unroll")
            for (int s = 1; s <= S; ++s) {
                double *ybuf = (s & 1) ? ybuf1 : ybuf0;
                // stage argument (elementwise, accumulator layout) -> shared memory [row][col]
                NBK_MMA_FOR_ELEMS({
                    double acc = 0.0;
                    bool first = true;
_Pragma("unroll")
                    for (int j = 0; j < s; ++j) {
                        const bool nz = s < S ? Tab::a(s < S ? s : 0, j) != 0.0 : Tab::b(j) != 0.0;
                        if (nz) {
                            const double cf = s < S ? Tab::cb().a[s < S ? s : 0][j] : Tab::cb().b[j];
                            acc = first ? cf * K[j][mt][nt][q] : fma(cf, K[j][mt][nt][q], acc);
                            first = false;
                        }
                    }
                    const double v = fma(hcol[nt][q], acc, y[mt][nt][q]);
                    if (s == S) ynew[mt][nt][q] = v;
                    ybuf[row * MMA_LD + col] = v;
                })
                mma_bar(grp);
                // K_s = A . Y_s on the FP64 tensor pipe: 2 MMA_MT accumulator tiles per warp, each summed as MMA_KSPLIT
                // interleaved partial sums over k (a DMMA into an accumulator waits for the previous one into it: with
                // 2 warps per scheduler and 4 accumulators per warp only 8 are in flight; two partial sums per
                // accumulator double that -- see DESIGN.md section 9, K4)
                double acc[MMA_KSPLIT][MMA_MT][MMA_NTN][2];
_Pragma("unroll")
                for (int ks = 0; ks < MMA_KSPLIT; ++ks)
_Pragma("unroll")
                    for (int mt = 0; mt < MMA_MT; ++mt)
_Pragma("unroll")
                        for (int nt = 0; nt < MMA_NTN; ++nt) acc[ks][mt][nt][0] = acc[ks][mt][nt][1] = 0.0;
                const int arow = MMA_ROWS * warp + (lane >> 2), kk = lane & 3;
                const double *ap = As + kk * MMA_LDA + arow, *bp = ybuf + kk * MMA_LD + (lane >> 2);
_Pragma("unroll 4")
                for (int k0 = 0; k0 < NPAD; k0 += 4 * MMA_KSPLIT) {
_Pragma("unroll")
                    for (int ks = 0; ks < MMA_KSPLIT; ++ks) {
                        const int k1 = k0 + 4 * ks;
                        // B fragments (4x8, "col"): element (k, col) = Y[k][col]
                        double bf[MMA_NTN];
_Pragma("unroll")
                        for (int nt = 0; nt < MMA_NTN; ++nt) bf[nt] = bp[k1 * MMA_LD + 8 * nt];
_Pragma("unroll")
                        for (int mt = 0; mt < MMA_MT; ++mt) {
                            // A fragment (8x4, row-major): element (row, k) = As[k][row] (zero beyond n)
                            const double am = ap[k1 * MMA_LDA + 8 * mt];
_Pragma("unroll")
                            for (int nt = 0; nt < MMA_NTN; ++nt) dmma_8x8x4(acc[ks][mt][nt][0], acc[ks][mt][nt][1], am, bf[nt]);
                        }
                    }
                }
_Pragma("unroll")
                for (int ks = 1; ks < MMA_KSPLIT; ++ks)
_Pragma("unroll")
                    for (int mt = 0; mt < MMA_MT; ++mt)
_Pragma("unroll")
                        for (int nt = 0; nt < MMA_NTN; ++nt) {
                            acc[0][mt][nt][0] += acc[ks][mt][nt][0];
                            acc[0][mt][nt][1] += acc[ks][mt][nt][1];
                        }
                NBK_MMA_FOR_ELEMS({ if (runc[nt][q]) K[s][mt][nt][q] = acc[0][mt][nt][q]; })
            }
            // ---- scaled error per column: partial over this thread's 2 rows, then lanes sharing the column
            double part[MMA_NTN][2];
_Pragma("unroll")
            for (int nt = 0; nt < MMA_NTN; ++nt)
_Pragma("unroll")
                for (int q = 0; q < 2; ++q) part[nt][q] = 0.0;
            NBK_MMA_FOR_ELEMS({
                double ee = 0.0;
                bool first = true;
_Pragma("unroll")
                for (int j = 0; j <= S; ++j) {
                    if (Tab::e(j) != 0.0) {
                        ee = first ? Tab::cb().e[j] * K[j][mt][nt][q] : fma(Tab::cb().e[j], K[j][mt][nt][q], ee);
                        first = false;
                    }
                }
                if (row < n) {
                    const double sc = pos_max(fma(rtol, fabs(ynew[mt][nt][q]), atol), fma(rtol, fabs(y[mt][nt][q]), atol));
                    const double r = ee * fast_rcp(sc);
                    part[nt][q] = fma(r, r, part[nt][q]);
                }
            })
_Pragma("unroll")
            for (int nt = 0; nt < MMA_NTN; ++nt)
_Pragma("unroll")
                for (int q = 0; q < 2; ++q) {
                    double v = part[nt][q];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (lane < 4) red[warp * MMA_NT + 8 * nt + 2 * lane + q] = v;
                }
            mma_bar(grp);
            // ---- controller: one leader thread per column
            if (tid < MMA_NT) {
                const int c = tid;
                if (cs->running[c]) {
                    double tot = 0.0;
_Pragma("unroll")
                    for (int w = 0; w < MMA_WARPS; ++w) tot += red[w * MMA_NT + c];
                    const double h = cs->h[c];
                    const double err2 = tot * (h * h / (double)n);
                    const bool accept = (unsigned)__double2hiint(err2) < 0x3ff00000u;
                    const double t_new = cs->t[c] + h;
                    const double h_new = h * step_factor<Tab>(err2, a.safety, a.min_factor, a.max_factor, a.x_lo, a.x_hi);
                    const int natt = ++cs->natt[c];
                    int done = 0, status = cs->status[c];
                    int kout = cs->kout[c];
                    cs->emit_from[c] = kout;
                    if (accept) {
                        if (MODE == NBK_MODE_RUN) {
                            double te = cs->te_next[c];
                            while (t_new >= te) {
                                ++kout;
                                te = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                            }
                            cs->te_next[c] = te;
                            done = kout >= a.m;
                        } else {
                            if (a.emit) a.t_out[cs->idx[c] * a.to_sb + cs->nacc[c]] = t_new;
                            done = cs->nacc[c] + 1 >= a.n_steps;
                        }
                    }
                    if (d_nonfinite(err2) || d_collapsed(h_new)) {
                        status = NBK_TRAJ_NONFINITE;
                        done = 1;
                    } else if (!done && natt >= max_att) {
                        status = NBK_TRAJ_MAXATTEMPTS;
                        done = 1;
                    }
                    if (accept && !done && (t_new + h_new > bound)) {
                        status = NBK_TRAJ_BOUND;
                        done = 1;
                    }
                    cs->kout[c] = kout;
                    cs->status[c] = status;
                    cs->accept[c] = accept;
                    cs->done[c] = done;
                    cs->t_new[c] = t_new;
                    cs->h[c] = h_new;
                    // the column moves on here; the element threads finish the attempt below from act / t_prev / nacc_prev
                    cs->act[c] = 1;
                    cs->t_prev[c] = cs->t[c];
                    cs->nacc_prev[c] = cs->nacc[c];
                    if (done) {
                        a.ts[cs->idx[c] * 2] = cs->t[c];  // history slot 0 (see nbk_cta.cuh)
                        cs->running[c] = 0;
                    }
                    if (accept) {
                        cs->t[c] = t_new;
                        ++cs->nacc[c];
                    }
                    cs->fresh[c] = accept;
                } else {
                    cs->act[c] = 0;
                    cs->accept[c] = 0;
                    cs->done[c] = 0;
                }
            }
            any_running = mma_bar_or(grp, tid < MMA_NT ? cs->running[tid] : 0);
            // ---- emission, history stash, commit (every thread for its own elements)
            NBK_MMA_FOR_ELEMS({
                const long long idx = cs->idx[col];
                if (cs->act[col] && row < n) {
                    const bool accept = cs->accept[col];
                    if (accept) {
                        if (MODE == NBK_MODE_RUN) {
                            const int k1 = cs->kout[col];
                            if (cs->emit_from[col] < k1) {
                                double yo[1] = {y[mt][nt][q]}, yn[1] = {ynew[mt][nt][q]}, Ke[S + 1][1], out[1];
_Pragma("unroll")
                                for (int s = 0; s <= S; ++s) Ke[s][0] = K[s][mt][nt][q];
                                for (int k = cs->emit_from[col]; k < k1; ++k) {
                                    rk_dense<Tab, 1>(__ldg(a.tgrid + k), cs->t_prev[col], cs->t_new[col], yo, yn, Ke, out);
                                    a.y_out[idx * a.o_sb + k * a.o_sk + (long long)row * a.o_sc] = out[0];
                                }
                            }
                        } else if (a.emit) {
                            a.y_out[idx * a.o_sb + cs->nacc_prev[col] * a.o_sk + (long long)row * a.o_sc] = ynew[mt][nt][q];
                        }
                    }
                    if (cs->done[col]) a.ys[(idx * 2) * nn + row] = y[mt][nt][q];  // history slot 0 (see nbk_cta.cuh)
                    if (accept) y[mt][nt][q] = ynew[mt][nt][q];
                }
            })
            // no barrier here: the next attempt's first writes to the fields read above (act, accept, kout, emit_from,
            // t_prev, t_new, done, nacc_prev) come from its leaders, S + 1 barriers further on
        }
        // ---------------------------------------------------------------- retire the tile
        NBK_MMA_FOR_ELEMS({
            const long long idx = cs->idx[col];
            if (idx >= 0 && row < n && cs->nacc[col] > 0) {
                const bool clean = cs->status[col] == NBK_TRAJ_OK || cs->status[col] == NBK_TRAJ_BOUND;
                a.ys[(idx * 2 + 1) * nn + row] = y[mt][nt][q];
                a.fs[(idx * 2) * nn + row] = K[0][mt][nt][q];
                a.fs[(idx * 2 + 1) * nn + row] = clean ? K[S][mt][nt][q] : K[0][mt][nt][q];
_Pragma("unroll")
                for (int s = 0; s <= S; ++s) a.K[(idx * (S + 1) + s) * nn + row] = K[s][mt][nt][q];
            }
        })
        if (tid < MMA_NT && cs->idx[tid] >= 0) {
            const int c = tid;
            const long long idx = cs->idx[c];
            if (cs->nacc[c] > 0) a.ts[idx * 2 + 1] = cs->t[c];
            a.h[idx] = cs->h[c];
            if (a.pristine) {
                a.n_acc[idx] = cs->nacc[c];
                a.n_rej[idx] = cs->natt[c] - cs->nacc[c];
            } else {
                a.n_acc[idx] += cs->nacc[c];
                a.n_rej[idx] += cs->natt[c] - cs->nacc[c];
            }
            a.status[idx] = cs->status[c];
            if (cs->status[c] >= NBK_TRAJ_NONFINITE) atomicAdd(a.work_counter + 1, 1ULL);
            a.n_out[idx] = MODE == NBK_MODE_RUN ? cs->kout[c] : cs->nacc[c];
        }
    }
}

}  // namespace nbk
#endif
