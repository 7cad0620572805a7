// nbk_device.cuh -- device side of the B200 ODE-ensemble engine (sm_100a, FP64).
//
// One trajectory per thread; the state vector, the stage derivatives K0..K_S and the
// parameters live in registers (every loop below has a compile-time trip count and is fully
// unrolled, so all array indices are static).  HBM is touched only to load a trajectory,
// to emit dense output and to retire the trajectory.  The adaptive kernels are persistent:
// a lane that finishes its trajectory pulls the next one from a global queue so that
// per-trajectory step-count variance (adaptive-step divergence) does not idle the warp.
//
// The same file is compiled by nvcc (built-in right-hand sides, nbk_program.cu) and by
// NVRTC at run time (user right-hand sides), so it must not include host headers.
//
// Reference behaviour restated here (paths under /root/reference):
//   FSAL stage loop            nbkode/runge_kutta/explicit.py:60-79
//   error estimate, norm       nbkode/runge_kutta/core.py:73-83
//   I-controller               nbkode/runge_kutta/core.py:85-99
//   accept/retry loop          nbkode/runge_kutta/explicit.py:81-99
//   t_bound guard              nbkode/core.py:176-184
//   RK dense output            nbkode/runge_kutta/explicit.py:354-403
//   initial step               nbkode/core.py:695-705 (scipy select_initial_step)
//   history window             nbkode/buffer.py:12-56
//   Adams-Bashforth step       nbkode/multistep/core.py:90-112, multistep/adams_bashforth.py
//   multistep start-up         nbkode/multistep/core.py:37-65
//   window Hermite             nbkode/core.py:577-596
//   run(ts) / step driver      nbkode/core.py:483-537, 598-619
#ifndef NBK_DEVICE_CUH
#define NBK_DEVICE_CUH

#include "nbk_kargs.h"
#include "nbk_dop853_coeffs.h"

#define NBK_DEV __device__ __forceinline__
#define NBK_HD __host__ __device__ constexpr

#ifndef NBK_REFILL_MIN
#define NBK_REFILL_MIN 2  // idle lanes needed before a warp stops to refill from the queue
#endif
#ifndef NBK_BLOCK
#define NBK_BLOCK 128
#endif
#ifndef NBK_INNER
#define NBK_INNER 16  // attempts between two looks at the trajectory queue (measured, round 1: 1 -> 7.48 ms, 4 -> 7.31, 8 -> 7.26;
                      // round 2, with the tail compaction: 4 -> 6.95, 8 -> 6.85, 12 -> 6.82, 16 -> 6.80, 24 -> 6.79, 32 -> 6.80;
                      // 125 000 trajectories: 8 -> 1.117, 16 -> 1.106, 32 -> 1.127)
#endif
#ifndef NBK_INNER_ADAPT
#define NBK_INNER_ADAPT 1
#endif
#ifdef NBK_HOST_EMUL
#define NBK_LOG2F(x) log2f(x)
#define NBK_EXP2F(x) exp2f(x)
#else
// single MUFU.LG2 / MUFU.EX2 (the argument is a normal float well inside the range)
static __device__ __forceinline__ float nbk_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
static __device__ __forceinline__ float nbk_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#define NBK_LOG2F(x) nbk_lg2(x)
#define NBK_EXP2F(x) nbk_ex2(x)
#endif
#ifndef NBK_FULL_MASK
#define NBK_FULL_MASK 0xffffffffu  // tests/emul compiles the kernels for the host as 1-lane warps
#endif
// Tail compaction of the persistent adaptive kernels (advance_adaptive): once the trajectory queue is empty, the warps
// of a CTA merge their surviving trajectories through a shared-memory pool so that whole warps retire instead of
// running on with a few live lanes each.  Off in the host emulation (one-lane warps, no shared memory).
#ifndef NBK_COMPACT
#ifdef NBK_HOST_EMUL
#define NBK_COMPACT 0
#else
#define NBK_COMPACT 1
#endif
#endif
#ifndef NBK_POOL_BYTES
#define NBK_POOL_BYTES (36 * 1024)  // shared memory a CTA may spend on the pool
#endif

namespace nbk {

// max/min as one DSETP + selects.  fmax()/fmin() cost ~7 instructions each for their NaN rules;
// every caller below tests finiteness separately, so a NaN may fall through either way.
NBK_DEV double dmax(double a, double b) { return a > b ? a : b; }
NBK_DEV double dmin(double a, double b) { return a < b ? a : b; }

// The same for operands known to be non-negative: IEEE doubles >= 0 order like their bit patterns,
// so the comparison runs on the integer pipe and leaves the FP64 pipe (the bottleneck) alone.
NBK_DEV bool pos_less(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }
NBK_DEV double pos_max(double a, double b) { return pos_less(a, b) ? b : a; }
NBK_DEV double pos_min(double a, double b) { return pos_less(a, b) ? a : b; }
NBK_DEV double abs_max(double a, double b) {
    const long long ia = __double_as_longlong(a) & 0x7fffffffffffffffLL, ib = __double_as_longlong(b) & 0x7fffffffffffffffLL;
    return __longlong_as_double(ia < ib ? ib : ia);
}

NBK_DEV double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
NBK_DEV double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
// inf/NaN and "h <= 0 or denormal" tests on the integer pipe (the FP64 pipe is the busy one)
NBK_DEV bool d_nonfinite(double x) { return (__double2hiint(x) & 0x7ff00000) == 0x7ff00000; }
NBK_DEV bool d_collapsed(double h) { return __double2hiint(h) < 0x00100000; }

// =====================================================================================
// Tableaux.  constexpr accessors: after unrolling every coefficient is an immediate /
// constant-bank operand and zero entries cost nothing.
// =====================================================================================

struct cbank_rk23;
struct cbank_rk45;
struct tab_rk23 {  // Bogacki-Shampine 3(2), nbkode/runge_kutta/explicit.py:146-151
    static constexpr int S = 3, Q = 2, M = 3, ROOT = 6, ID = 23;
    static constexpr bool DOP = false;
    typedef cbank_rk23 bank;
    static __device__ __forceinline__ const cbank_rk23 &cb();
    static NBK_HD double a(int i, int j) {
        constexpr double A[3][3] = {{0, 0, 0}, {1.0 / 2, 0, 0}, {0, 3.0 / 4, 0}};
        return A[i][j];
    }
    static NBK_HD double b(int j) {
        constexpr double Bv[3] = {2.0 / 9, 1.0 / 3, 4.0 / 9};
        return Bv[j];
    }
    static NBK_HD double c(int j) {
        constexpr double Cv[3] = {0, 1.0 / 2, 3.0 / 4};
        return Cv[j];
    }
    static NBK_HD double e(int j) {  // FSAL.E = [B,0] - B2 (explicit.py:54-58)
        constexpr double Ev[4] = {2.0 / 9 - 7.0 / 24, 1.0 / 3 - 1.0 / 4, 4.0 / 9 - 1.0 / 3, -1.0 / 8};
        return Ev[j];
    }
    static NBK_HD double p(int j, int k) {
        constexpr double Pv[4][3] = {{1, -4.0 / 3, 5.0 / 9}, {0, 1, -2.0 / 3}, {0, 4.0 / 3, -8.0 / 9}, {0, -1, 1}};
        return Pv[j][k];
    }
};

struct tab_rk45 {  // Dormand-Prince 5(4), nbkode/runge_kutta/explicit.py:185-239
    static constexpr int S = 6, Q = 4, M = 4, ROOT = 10, ID = 45;
    static constexpr bool DOP = false;
    typedef cbank_rk45 bank;
    static __device__ __forceinline__ const cbank_rk45 &cb();
    static NBK_HD double a(int i, int j) {
        constexpr double A[6][6] = {
            {0, 0, 0, 0, 0, 0},
            {1.0 / 5, 0, 0, 0, 0, 0},
            {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
            {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
            {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
            {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0}};
        return A[i][j];
    }
    static NBK_HD double b(int j) {
        constexpr double Bv[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
        return Bv[j];
    }
    static NBK_HD double c(int j) {
        constexpr double Cv[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
        return Cv[j];
    }
    static NBK_HD double e(int j) {
        constexpr double Ev[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
        return Ev[j];
    }
    static NBK_HD double p(int j, int k) {
        constexpr double Pv[7][4] = {
            {1, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
            {0, 0, 0, 0},
            {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
            {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
            {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
            {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
            {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}};
        return Pv[j][k];
    }
};

// Constant-bank copies of the same numbers.  The constexpr accessors above decide at compile
// time WHICH terms exist (zero entries are skipped); the values themselves are read as
// c[bank][offset] operands of the DFMA/DMUL, which costs no instruction -- an immediate
// double would need two UMOVs each (measured: 61 UMOVs per RK45 attempt, profiles/r01).
struct cbank_rk23 { double a[3][3], b[3], c[3], e[4], p[4][3]; };
struct cbank_rk45 { double a[6][6], b[6], c[6], e[7], p[7][4]; };
static __constant__ cbank_rk23 nbk_c23 = {
    {{0, 0, 0}, {1.0 / 2, 0, 0}, {0, 3.0 / 4, 0}},
    {2.0 / 9, 1.0 / 3, 4.0 / 9},
    {0, 1.0 / 2, 3.0 / 4},
    {2.0 / 9 - 7.0 / 24, 1.0 / 3 - 1.0 / 4, 4.0 / 9 - 1.0 / 3, -1.0 / 8},
    {{1, -4.0 / 3, 5.0 / 9}, {0, 1, -2.0 / 3}, {0, 4.0 / 3, -8.0 / 9}, {0, -1, 1}}};
static __constant__ cbank_rk45 nbk_c45 = {
    {{0, 0, 0, 0, 0, 0},
     {1.0 / 5, 0, 0, 0, 0, 0},
     {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
     {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
     {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
     {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0}},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84},
    {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1},
    {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40},
    {{1, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},
     {0, 0, 0, 0},
     {0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},
     {0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},
     {0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},
     {0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},
     {0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}}};

// DOP853 (nbkode/runge_kutta/explicit.py:282-351): 12 stages + FSAL, error estimate from the E5 and E3
// rows, 7th-order dense output from three extra stages (rows 13..15 of A/C) and D.  Coefficients:
// nbk_dop853_coeffs.h (generated from SciPy; identical to nbkode/runge_kutta/dop853_coefficients.py).
struct cbank_dop853 { double a[16][16], b[12], c[16], e3[13], e5[13], d[4][16]; };
static __constant__ cbank_dop853 nbk_c853 = {NBK_DOP853_A_INIT, NBK_DOP853_B_INIT, NBK_DOP853_C_INIT,
                                            NBK_DOP853_E3_INIT, NBK_DOP853_E5_INIT, NBK_DOP853_D_INIT};
struct tab_dop853 {
    static constexpr int S = 12, Q = 7, M = 7, ROOT = 16, ID = 853;
    static constexpr bool DOP = true;
    static NBK_HD double a(int i, int j) {
        constexpr double A[16][16] = NBK_DOP853_A_INIT;
        return A[i][j];
    }
    static NBK_HD double b(int j) {
        constexpr double Bv[12] = NBK_DOP853_B_INIT;
        return Bv[j];
    }
    static NBK_HD double c(int j) {
        constexpr double Cv[16] = NBK_DOP853_C_INIT;
        return Cv[j];
    }
    static NBK_HD double e(int j) { return e5(j); }
    static NBK_HD double e5(int j) {
        constexpr double Ev[13] = NBK_DOP853_E5_INIT;
        return Ev[j];
    }
    static NBK_HD double e3(int j) {
        constexpr double Ev[13] = NBK_DOP853_E3_INIT;
        return Ev[j];
    }
    static NBK_HD double d(int i, int j) {
        constexpr double Dv[4][16] = NBK_DOP853_D_INIT;
        return Dv[i][j];
    }
    typedef cbank_dop853 bank;
    static __device__ __forceinline__ const cbank_dop853 &cb() { return nbk_c853; }
};

__device__ __forceinline__ const cbank_rk23 &tab_rk23::cb() { return nbk_c23; }
__device__ __forceinline__ const cbank_rk45 &tab_rk45::cb() { return nbk_c45; }

template <int ORDER>
struct tab_ab {
    static constexpr int ID = ORDER, LEN = ORDER > 2 ? ORDER : 2;
    static NBK_HD double b(int j) {
        constexpr double Bv[6][5] = {
            {0, 0, 0, 0, 0},
            {1.0, 0, 0, 0, 0},
            {3.0 / 2, -1.0 / 2, 0, 0, 0},
            {23.0 / 12, -16.0 / 12, 5.0 / 12, 0, 0},
            {55.0 / 24, -59.0 / 24, 37.0 / 24, -9.0 / 24, 0},
            {1901.0 / 720, -2774.0 / 720, 2616.0 / 720, -1274.0 / 720, 251.0 / 720}};
        return Bv[ORDER][j];
    }
};

// Implicit linear multistep classes of the reference: AdamsMoulton1..5 (ids 11..15; BackwardEuler is
// AdamsMoulton1) and BDF1..6 (ids 31..36), nbkode/multistep/adams_moulton.py:31-61, bdf.py:16-44.
// y_new solves  y - h Bn f(t_new, y) - (h B . fs[-K:] - A . ys[-K:]) = 0  (multistep/core.py:115-170); like the
// Adams-Bashforth classes, index 0 of A / B multiplies the OLDEST history entry.  K = A.size is the reference's
// ORDER (AdamsMoulton2 therefore has K = 1), LEN = max(K, 2) its LEN_HISTORY.
template <int ID>
struct tab_im {
    static constexpr bool BDF = ID >= 31;
    static constexpr int K = BDF ? ID - 30 : (ID <= 12 ? 1 : ID - 11);
    static constexpr int LEN = K > 2 ? K : 2;
    static NBK_HD double a(int j) {
        constexpr double Abdf[7][6] = {
            {0},
            {-1.0},
            {1.0 / 3, -4.0 / 3},
            {-2.0 / 11, 9.0 / 11, -18.0 / 11},
            {3.0 / 25, -16.0 / 25, 36.0 / 25, -48.0 / 25},
            {-12.0 / 137, 75.0 / 137, -200.0 / 137, 300.0 / 137, -300.0 / 137},
            {10.0 / 147, -72.0 / 147, 225.0 / 147, -400.0 / 147, 450.0 / 147, -360.0 / 147}};
        return BDF ? Abdf[K][j] : (j == K - 1 ? -1.0 : 0.0);
    }
    static NBK_HD double b(int j) {
        constexpr double Bam[6][4] = {
            {0},
            {0.0},
            {1.0 / 2},
            {-1.0 / 12, 2.0 / 3},
            {19.0 / 24, -5.0 / 24, 1.0 / 24},
            {646.0 / 720, -264.0 / 720, 106.0 / 720, -19.0 / 720}};
        return BDF ? 0.0 : Bam[ID - 10][j];
    }
    static NBK_HD double bn() {
        constexpr double BNam[6] = {0, 1.0, 1.0 / 2, 5.0 / 12, 9.0 / 24, 251.0 / 720};
        constexpr double BNbdf[7] = {0, 1.0, 2.0 / 3, 6.0 / 11, 12.0 / 25, 60.0 / 137, 60.0 / 147};
        return BDF ? BNbdf[K] : BNam[ID - 10];
    }
};

// Fixed-step explicit Runge-Kutta classes of the reference (nbkode/runge_kutta/explicit.py:102-129):
// Runge2 (202), Runge3 (203, four stages as the reference lists it), Heun3 (213), RungeKutta4 (204),
// RungeKutta3_8 (238).
template <int ID>
struct tab_erk {
    static constexpr int S = ID == 202 ? 2 : (ID == 213 ? 3 : 4);
    static NBK_HD double a(int i, int j) {
        constexpr double A202[4][4] = {{0}, {1.0 / 2}};
        constexpr double A203[4][4] = {{0}, {1.0 / 2}, {0, 1}, {0, 0, 1}};
        constexpr double A213[4][4] = {{0}, {1.0 / 3}, {0, 2.0 / 3}};
        constexpr double A204[4][4] = {{0}, {1.0 / 2}, {0, 1.0 / 2}, {0, 0, 1}};
        constexpr double A238[4][4] = {{0}, {1.0 / 3}, {-1.0 / 3, 1}, {1, -1, 1}};
        return ID == 202 ? A202[i][j] : ID == 203 ? A203[i][j] : ID == 213 ? A213[i][j] : ID == 204 ? A204[i][j] : A238[i][j];
    }
    static NBK_HD double b(int j) {
        constexpr double B202[4] = {0, 1.0};
        constexpr double B203[4] = {1.0 / 6, 4.0 / 6, 0, 1.0 / 6};
        constexpr double B213[4] = {1.0 / 4, 0, 3.0 / 4};
        constexpr double B204[4] = {1.0 / 6, 2.0 / 6, 2.0 / 6, 1.0 / 6};
        constexpr double B238[4] = {1.0 / 8, 3.0 / 8, 3.0 / 8, 1.0 / 8};
        return ID == 202 ? B202[j] : ID == 203 ? B203[j] : ID == 213 ? B213[j] : ID == 204 ? B204[j] : B238[j];
    }
    static NBK_HD double c(int j) {
        constexpr double C202[4] = {0, 1.0 / 2};
        constexpr double C203[4] = {0, 1.0 / 2, 1, 1};
        constexpr double C213[4] = {0, 1.0 / 3, 2.0 / 3};
        constexpr double C204[4] = {0, 1.0 / 2, 1.0 / 2, 1};
        constexpr double C238[4] = {0, 1.0 / 3, 2.0 / 3, 1};
        return ID == 202 ? C202[j] : ID == 203 ? C203[j] : ID == 213 ? C213[j] : ID == 204 ? C204[j] : C238[j];
    }
};

// runtime-order copy for the (cold) multistep start-up with an Adams-Bashforth first stepper
NBK_DEV double ab_weight(int order, int j) {
    const double Bv[6][5] = {
        {0, 0, 0, 0, 0},
        {1.0, 0, 0, 0, 0},
        {3.0 / 2, -1.0 / 2, 0, 0, 0},
        {23.0 / 12, -16.0 / 12, 5.0 / 12, 0, 0},
        {55.0 / 24, -59.0 / 24, 37.0 / 24, -9.0 / 24, 0},
        {1901.0 / 720, -2774.0 / 720, 2616.0 / 720, -1274.0 / 720, 251.0 / 720}};
    return Bv[order][j];
}

// =====================================================================================
// Built-in right-hand sides (the catalogue of BASELINE.json's configs that fit one thread).
// A right-hand side is a struct with N, NP and
//   static void eval(double t, const double (&y)[N], const double (&p)[NPX], double (&dy)[N]).
// =====================================================================================

template <int N_, int NP_>
struct rhs_decay {  // y' = -k*y, k scalar (NP=1) or element-wise (NP=N): README.rst:44-50
    static constexpr int N = N_, NP = NP_;
    static NBK_DEV void eval(double, const double (&y)[N_], const double (&p)[NP_], double (&dy)[N_]) {
#pragma unroll
        for (int i = 0; i < N_; ++i) dy[i] = -p[NP_ == 1 ? 0 : i] * y[i];
    }
};

typedef rhs_decay<1, 1> rhs_decay11;  // catalogue instance (nvcc -D cannot carry a comma)

struct rhs_lorenz {  // Lorenz-63: [sigma (y-x), x (rho-z) - y, x y - beta z]
    static constexpr int N = 3, NP = 3;
    static NBK_DEV void eval(double, const double (&y)[3], const double (&p)[3], double (&dy)[3]) {
        dy[0] = p[0] * (y[1] - y[0]);
        dy[1] = y[0] * (p[1] - y[2]) - y[1];
        dy[2] = y[0] * y[1] - p[2] * y[2];
    }
};

struct rhs_vdp {  // Van der Pol: [v, mu (1-x^2) v - x]
    static constexpr int N = 2, NP = 1;
    static NBK_DEV void eval(double, const double (&y)[2], const double (&p)[1], double (&dy)[2]) {
        dy[0] = y[1];
        dy[1] = p[0] * (1 - y[0] * y[0]) * y[1] - y[0];
    }
};

#ifdef NBK_USER_RHS
// User right-hand side compiled by NVRTC: the user source defines
//   __device__ void nbk_rhs(double t, const double* y, const double* p, double* dy)
// with NBK_N state variables and NBK_NP parameters.
struct rhs_user {
    static constexpr int N = NBK_N, NP = NBK_NP;
    static NBK_DEV void eval(double t, const double (&y)[NBK_N], const double (&p)[NBK_NP > 0 ? NBK_NP : 1],
                             double (&dy)[NBK_N]) {
        nbk_rhs(t, y, p, dy);
    }
};
#endif

// =====================================================================================
// Arithmetic helpers
// =====================================================================================

// 1/x for positive normal x: MUFU.RCP64H seed (relative error e <= 2^-23) and ONE Newton step
// r (1 + e): two dependent DFMA instead of the ~12 instruction IEEE division sequence.  Truncation
// error e^2 <= 2^-46 (7e-15): it only scales the error estimate, whose own cancellation noise
// (E.K loses ~5 digits) is three orders of magnitude larger.  (A third-order step, 2^-69, costs one
// more DFMA per component: measured +0.6 % kernel time.)
NBK_DEV double fast_rcp(double x) {
    double r;
#ifdef NBK_HOST_EMUL
    r = (double)(float)(1.0 / x);  // stands in for the MUFU seed
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    const double e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// a / b for positive normal b: reciprocal seed + Newton step, then one residual correction of the quotient
// (result within ~1 ulp; 6 instructions instead of the ~14 + slow path of the IEEE sequence).
NBK_DEV double fast_div(double a, double b) {
    const double r = fast_rcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// x^(-1/R) for a positive normal x inside float range ([1e-30, 1e30], guaranteed by the clamp in
// step_factor).  Seed: the top bits of x reinterpreted as a float (integer ops, no F2F), then
// lg2/ex2 MUFU: relative error eps ~ 2e-7.  Polish: ONE third-order step on g(y) = y^-R - x,
//   r = 1 - x y^R,   y <- y (1 + r/R + (R+1) r^2 / (2 R^2)),
// i.e. the Taylor series of (1-r)^(-1/R); with |r| ~ R eps ~ 2e-6 the dropped r^3 term is ~3e-19,
// so the step-size factor is accurate to double rounding (a few ulp).
template <int R>
NBK_DEV double inv_root(double x) {
    const int hi = __double2hiint(x);
    const unsigned lo = (unsigned)__double2loint(x);
    const float xf = __int_as_float(((hi - 0x38000000) << 3) | (int)(lo >> 29));
    const int sb = __float_as_int(NBK_EXP2F(NBK_LOG2F(xf) * (-1.0f / R)));
    const double y = __hiloint2double((sb >> 3) + 0x38000000, sb << 29);
    const double y2 = y * y;
    double yr;
    if (R == 10) {
        const double y4 = y2 * y2, y8 = y4 * y4;
        yr = y8 * y2;
    } else if (R == 6) {
        const double y4 = y2 * y2;
        yr = y4 * y2;
    } else if (R == 16) {
        const double y4 = y2 * y2, y8 = y4 * y4;
        yr = y8 * y8;
    } else {
        yr = 1.0;
#pragma unroll
        for (int k = 0; k < R; ++k) yr *= y;
    }
    const double r = fma(-x, yr, 1.0);
    const double q = fma(r, (R + 1.0) / (2.0 * R * R), 1.0 / R) * r;
    return fma(y, q, y);
}

// rms(x / scale) of scipy.integrate._ivp.common.norm (cold path: IEEE division and sqrt)
template <int N>
NBK_DEV double rms_scaled(const double (&x)[N], const double (&sc)[N]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double v = x[i] / sc[i];
        acc = fma(v, v, acc);
    }
    return sqrt(acc) / sqrt((double)N);
}

// =====================================================================================
// One FSAL attempt: K[0] must hold f(t, y).  Produces K[1..S], y_new and the squared
// scaled RMS error  err2 = mean((h E.K / (atol + rtol max(|y_new|,|y|)))^2).
// Algorithmic FLOPs: RK45 6R + 66N + 18, RK23 3R + 27N + 18.
// =====================================================================================
template <class Tab, class Rhs>
NBK_DEV void rk_attempt(double t, double h, const double (&y)[Rhs::N], double (&K)[Tab::S + 1][Rhs::N],
                        const double (&p)[Rhs::NP > 0 ? Rhs::NP : 1], double (&ynew)[Rhs::N], double &err2,
                        double rtol, double atol) {
    constexpr int N = Rhs::N, S = Tab::S;
    const typename Tab::bank &cb = Tab::cb();
#pragma unroll
    for (int s = 1; s < S; ++s) {
        double ys[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < s; ++j) {
                if (Tab::a(s, j) != 0.0) {
                    acc = first ? cb.a[s][j] * K[j][i] : fma(cb.a[s][j], K[j][i], acc);
                    first = false;
                }
            }
            ys[i] = fma(h, acc, y[i]);
        }
        Rhs::eval(fma(h, cb.c[s], t), ys, p, K[s]);
    }
    // y_new = y + h (B . K[:S]).  (The reference forms the vector h*B first; the difference is one
    // rounding, far inside the parity bars, and saves the S scalar multiplies.)
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            if (Tab::b(j) != 0.0) {
                acc = first ? cb.b[j] * K[j][i] : fma(cb.b[j], K[j][i], acc);
                first = false;
            }
        }
        ynew[i] = fma(h, acc, y[i]);
    }
    Rhs::eval(t + h, ynew, p, K[S]);
    if constexpr (Tab::DOP) {
        // err5 = h E5.K, err3 = h E3.K; norm = err5n / sqrt(1 + (err3n / (10 err5n))^2)  (explicit.py:325-336)
        // in squared terms with a = err5n^2, b = err3n^2:  norm^2 = a^2 / (a + b/100)
        double a5 = 0.0, a3 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double e5 = 0.0, e3 = 0.0;
            bool f5 = true, f3 = true;
#pragma unroll
            for (int j = 0; j <= S; ++j) {
                if (Tab::e5(j) != 0.0) {
                    e5 = f5 ? cb.e5[j] * K[j][i] : fma(cb.e5[j], K[j][i], e5);
                    f5 = false;
                }
                if (Tab::e3(j) != 0.0) {
                    e3 = f3 ? cb.e3[j] * K[j][i] : fma(cb.e3[j], K[j][i], e3);
                    f3 = false;
                }
            }
            const double rs = fast_rcp(pos_max(fma(rtol, fabs(ynew[i]), atol), fma(rtol, fabs(y[i]), atol)));
            const double r5 = e5 * rs, r3 = e3 * rs;
            a5 = fma(r5, r5, a5);
            a3 = fma(r3, r3, a3);
        }
        const double hh = h * h * (1.0 / N);
        a5 *= hh;
        a3 *= hh;
        const double den = fma(0.01, a3, a5);
        err2 = den > 0.0 ? a5 * a5 / den : a5;  // a5 == a3 == 0: exact step (the reference divides 0/0 here)
    } else {
        double acc2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double e = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j <= S; ++j) {
                if (Tab::e(j) != 0.0) {
                    e = first ? cb.e[j] * K[j][i] : fma(cb.e[j], K[j][i], e);
                    first = false;
                }
            }
            // atol + rtol max(|a|, |b|) == max(atol + rtol |a|, atol + rtol |b|) bit for bit (monotone); the |.| are
            // free operand modifiers of DFMA, the max of two positive doubles runs on the integer pipe
            const double sc = pos_max(fma(rtol, fabs(ynew[i]), atol), fma(rtol, fabs(y[i]), atol));
            const double r = e * fast_rcp(sc);
            acc2 = fma(r, r, acc2);
        }
        err2 = acc2 * (h * h * (1.0 / N));  // h factored out of the N components
    }
}

// I-controller of the reference: h *= clip(safety * norm^(-1/(q+1)), min_factor, max_factor)
// on accept AND reject; accept iff norm < 1 (norm == 0 -> max_factor, accept).  Works on the
// squared norm: norm^(-1/(q+1)) == err2^(-1/(2(q+1))), and sqrt(x) < 1 <=> x < 1.
// [x_lo, x_hi] (host-computed from the options, nbk_api.cu) brackets the range of err2 in which
// the clip is not already decided; clamping into it keeps the single-precision seed in range.
template <class Tab>
NBK_DEV double step_factor(double err2, double safety, double min_factor, double max_factor, double x_lo,
                           double x_hi) {
    const double xc = pos_min(pos_max(err2, x_lo), x_hi);  // err2 >= 0 (or NaN, caught by the caller)
    const double f = safety * inv_root<Tab::ROOT>(xc);
    return pos_min(max_factor, pos_max(min_factor, f));
}

// Dense output inside the last accepted step [t_old, t_new] (RkInterpolator, explicit.py:354-403):
// Q = K^T P once per step (RkInterpolator.update), then y(te) = y_old + H (Q . [x, x^2, ...]), x = (te - t_old) / H.
template <class Tab, int N>
NBK_DEV void rk_dense_coeffs(const double (&K)[Tab::S + 1][N], double (&Q)[N][Tab::M]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int k = 0; k < Tab::M; ++k) {
            double q = 0.0;
#pragma unroll
            for (int j = 0; j <= Tab::S; ++j)
                if (Tab::p(j, k) != 0.0) q = fma(K[j][i], Tab::cb().p[j][k], q);
            Q[i][k] = q;
        }
    }
}

template <class Tab, int N, bool RCP = false>
NBK_DEV void rk_dense_eval(double te, double t_old, double t_new, const double (&y_old)[N], const double (&y_new)[N],
                           const double (&Q)[N][Tab::M], double (&out)[N], double inv_H = 0.0) {
    if (te == t_new) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = y_new[i];
        return;
    }
    if (te == t_old) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = y_old[i];
        return;
    }
    const double H = t_new - t_old;
    // RCP: the caller divided 1 / H once for all grid points of the step (x within 1 ulp of the quotient)
    const double x = RCP ? (te - t_old) * inv_H : (te - t_old) / H;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = 0.0, pw = x;
#pragma unroll
        for (int k = 0; k < Tab::M; ++k) {
            acc = fma(Q[i][k], pw, acc);
            pw *= x;
        }
        out[i] = fma(H, acc, y_old[i]);
    }
}

// The same interpolant for the emission loop of run(): te lies in (t_old, t_new], so only the exact return at the end
// of the step remains (a select, no branch), and with H x = te - t_old the polynomial is a plain Horner scheme:
// y_old + (te - t_old) (Q0 + x (Q1 + x (...))) -- 2 + N M FP64 instructions per point instead of 4 + N (M + 1) + M.
template <class Tab, int N>
NBK_DEV void rk_dense_eval_in_step(double te, double t_old, double t_new, const double (&y_old)[N], const double (&y_new)[N],
                                   const double (&Q)[N][Tab::M], double (&out)[N], double inv_H) {
    const double dt = te - t_old, x = dt * inv_H;
    const bool at_end = te == t_new;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = Q[i][Tab::M - 1];
#pragma unroll
        for (int k = Tab::M - 2; k >= 0; --k) acc = fma(acc, x, Q[i][k]);
        const double v = fma(dt, acc, y_old[i]);
        out[i] = at_end ? y_new[i] : v;
    }
}

template <class Tab, int N>
NBK_DEV void rk_dense(double te, double t_old, double t_new, const double (&y_old)[N], const double (&y_new)[N],
                      const double (&K)[Tab::S + 1][N], double (&out)[N]) {
    double Q[N][Tab::M];
    rk_dense_coeffs<Tab, N>(K, Q);
    rk_dense_eval<Tab, N>(te, t_old, t_new, y_old, y_new, Q, out);
}

// DOP853 dense output (DOP853Interpolator, explicit.py:406-481): three extra stages (rows 13..15) built
// on the accepted step [t_old, t], then F[0..6]; evaluated by the alternating x / (1-x) Horner scheme.
template <class Rhs>
NBK_DEV void dop853_prepare(double t_old, double t, const double (&y_old)[Rhs::N], const double (&y)[Rhs::N],
                            const double (&K)[13][Rhs::N], const double (&p)[Rhs::NP > 0 ? Rhs::NP : 1],
                            double (&F)[7][Rhs::N]) {
    using Tab = tab_dop853;
    constexpr int N = Rhs::N;
    const double h = t - t_old;
    double Ke[3][N];
#pragma unroll
    for (int s = 13; s < 16; ++s) {
        double ys[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < s; ++j)
                if (Tab::a(s, j) != 0.0) acc = fma(Tab::cb().a[s][j], j < 13 ? K[j < 13 ? j : 0][i] : Ke[j >= 13 ? j - 13 : 0][i], acc);
            ys[i] = fma(acc, h, y_old[i]);
        }
        Rhs::eval(fma(Tab::cb().c[s], h, t_old), ys, p, Ke[s - 13]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double dy = y[i] - y_old[i];
        F[0][i] = dy;
        F[1][i] = h * K[0][i] - dy;
        F[2][i] = 2 * dy - h * (K[12][i] + K[0][i]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (Tab::d(r, j) != 0.0) acc = fma(Tab::cb().d[r][j], j < 13 ? K[j < 13 ? j : 0][i] : Ke[j >= 13 ? j - 13 : 0][i], acc);
            F[3 + r][i] = h * acc;
        }
    }
}

template <int N>
NBK_DEV void dop853_eval(double te, double t_old, double t, const double (&y_old)[N], const double (&y)[N],
                         const double (&F)[7][N], double (&out)[N]) {
    if (te == t) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = y[i];
        return;
    }
    if (te == t_old) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = y_old[i];
        return;
    }
    const double x = (te - t_old) / (t - t_old);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            acc += F[6 - k][i];
            acc *= (k % 2 == 0) ? x : (1 - x);
        }
        out[i] = acc + y_old[i];
    }
}

// scipy select_initial_step (direction +1, unbounded interval), SURVEY.md row a4.
template <class Rhs>
NBK_DEV double initial_step(double t0, const double (&y0)[Rhs::N], const double (&f0)[Rhs::N],
                            const double (&p)[Rhs::NP > 0 ? Rhs::NP : 1], int q, double rtol, double atol) {
    constexpr int N = Rhs::N;
    double sc[N], y1[N], f1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) sc[i] = atol + fabs(y0[i]) * rtol;
    const double d0 = rms_scaled<N>(y0, sc), d1 = rms_scaled<N>(f0, sc);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
#pragma unroll
    for (int i = 0; i < N; ++i) y1[i] = y0[i] + h0 * f0[i];
    Rhs::eval(t0 + h0, y1, p, f1);
#pragma unroll
    for (int i = 0; i < N; ++i) f1[i] -= f0[i];
    const double d2 = rms_scaled<N>(f1, sc) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / (q + 1));
    return fmin(100 * h0, h1);
}

// Final status of a trajectory for this call.  Failures are also counted in work_counter[1], so that the host
// learns "did anything fail" from one 8-byte copy instead of a reduction kernel (which could not start while a
// persistent kernel of another stream fills the GPU).
NBK_DEV void store_status(const nbk_kargs &a, long long idx, int status) {
    a.status[idx] = status;
    if (status >= NBK_TRAJ_NONFINITE) atomicAdd(a.work_counter + 1, 1ULL);
}

// -------------------------------------------------------------------------------------
// parameter fetch: SoA [np][B] or shared [np]
// -------------------------------------------------------------------------------------
template <int NP>
NBK_DEV void load_params(const nbk_kargs &a, long long idx, double (&p)[NP > 0 ? NP : 1]) {
    if (NP == 0) {
        p[0] = 0.0;
        return;
    }
#pragma unroll
    for (int c = 0; c < NP; ++c) p[c] = a.params_shared ? __ldg(a.params + c) : __ldg(a.params + (long long)c * a.B + idx);
}


// =====================================================================================
// Events (run_events, nbkode/core.py:334-447 + nbkode/event_handler.py).  Compiled only into
// programs built with -DNBK_NEVENTS=k: the user (or the Python tracer) supplies
//   NBK_EVENT nbk_event(int k, double t, const double* y, const double* p)
// After every accepted step each event function is evaluated at the new point; a sign change
// against its previous value (filtered by the event's direction) is located by bisection on the
// step's dense output -- tolerance |g| <= 1.48e-8, at most 50 halvings (nbcompat/zeros.py:508-533)
// -- and recorded with the interpolated state; a terminal event stops the trajectory and its
// (t, y) replaces the output row the trajectory was integrating towards (core.py:641-644).
// =====================================================================================
#ifdef NBK_NEVENTS
template <int N, int NPX>
NBK_DEV void events_start(const nbk_kargs &a, long long idx, double t, const double (&y)[N], const double (&p)[NPX],
                          double (&ev_last)[NBK_NEVENTS]) {
    // build_handler(events, self.t, self.y): last_value = func(t, y), empty lists (event_handler.py:127-135)
#pragma unroll
    for (int k = 0; k < NBK_NEVENTS; ++k) {
        ev_last[k] = nbk_event(k, t, y, p);
        a.ev_count[idx * NBK_NEVENTS + k] = 0;
    }
}

// Dense: functor (double te, double (&out)[N]) over the step [t_old, t_new] that was just accepted.
// Returns true if a terminal event fired; t_term then holds that event's time.
template <int N, int NPX, class Dense>
NBK_DEV bool events_after_step(const nbk_kargs &a, long long idx, double t_old, double t_new, const double (&y_new)[N],
                               const double (&p)[NPX], double (&ev_last)[NBK_NEVENTS], Dense &dense, double &t_term,
                               int &status) {
    bool terminate = false;
#pragma unroll
    for (int k = 0; k < NBK_NEVENTS; ++k) {
        const double value = nbk_event(k, t_new, y_new, p);
        const double last = ev_last[k];
        const bool up = last <= 0.0 && value >= 0.0, down = last >= 0.0 && value <= 0.0;
        const int dir = a.ev_dir[k];
        const bool trigger = (up && dir > 0) || (down && dir < 0) || ((up || down) && dir == 0);
        ev_last[k] = value;
        if (!trigger) continue;
        double root = t_new, out[N];
        if (value != 0.0) {
            double left = t_old, right = t_new;
            dense(left, out);
            double yl = nbk_event(k, left, out, p);
            if (yl == 0.0) {
                root = left;
            } else if (!(yl * value < 0.0)) {
                status = NBK_TRAJ_EVBRACKET;
                return true;
            } else {
                double mid = left;
                for (int it = 0; it < 50; ++it) {
                    mid = (left + right) / 2;
                    dense(mid, out);
                    const double ym = nbk_event(k, mid, out, p);
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
        const int cnt = a.ev_count[slot];
        // "This is required to avoid duplicates" (event_handler.py:158-160)
        const bool dup = cnt > 0 && cnt <= a.ev_cap && a.ev_t[slot * a.ev_cap + cnt - 1] == root;
        if (!dup) {
            if (cnt < a.ev_cap) {
                a.ev_t[slot * a.ev_cap + cnt] = root;
                dense(root, out);
#pragma unroll
                for (int c = 0; c < N; ++c) a.ev_y[(slot * a.ev_cap + cnt) * N + c] = out[c];
            }
            a.ev_count[slot] = cnt + 1;
        }
        if ((a.ev_terminal >> k) & 1) {
            // EventHandler.evaluate: min_t is never raised, so the LAST terminal event of the list wins
            terminate = true;
            t_term = root;
        }
    }
    return terminate;
}

// dense output of one accepted adaptive step (RkInterpolator / DOP853Interpolator), prepared lazily
template <class Tab, class Rhs>
struct step_dense {
    static constexpr int N = Rhs::N, NPX = Rhs::NP > 0 ? Rhs::NP : 1;
    double t_old, t_new;
    const double (&y_old)[N];
    const double (&y_new)[N];
    const double (&K)[Tab::S + 1][N];
    const double (&p)[NPX];
    double F[Tab::DOP ? 7 : 1][N];
    bool prepared;
    NBK_DEV step_dense(double t0, double t1, const double (&y0)[N], const double (&y1)[N], const double (&K_)[Tab::S + 1][N],
                       const double (&p_)[NPX])
        : t_old(t0), t_new(t1), y_old(y0), y_new(y1), K(K_), p(p_), prepared(false) {}
    NBK_DEV void operator()(double te, double (&out)[N]) {
        if constexpr (Tab::DOP) {
            if (!prepared) {
                dop853_prepare<Rhs>(t_old, t_new, y_old, y_new, K, p, F);
                prepared = true;
            }
            dop853_eval<N>(te, t_old, t_new, y_old, y_new, F, out);
        } else {
            rk_dense<Tab, N>(te, t_old, t_new, y_old, y_new, K, out);
        }
    }
};
#endif  // NBK_NEVENTS

// =====================================================================================
// Adaptive advance kernel body (RungeKutta23 / RungeKutta45), persistent with lane refill.
// MODE RUN : for ti in tgrid: while t < ti: step; emit dense(ti)      (core.py:598-619)
// MODE STEP: up to n_steps accepted steps, each emitted as (t, y)      (core.py:483-537)
// Every loop iteration is one ATTEMPT for every running lane; accept/reject is resolved with
// selects, so lanes of a warp only diverge when a trajectory is loaded, emits or retires.
// =====================================================================================
template <class Tab, class Rhs, int MODE, bool EV = false>
NBK_DEV void advance_adaptive(const nbk_kargs &a) {
    constexpr int N = Rhs::N, NP = Rhs::NP, NPX = NP > 0 ? NP : 1, S = Tab::S;
    const unsigned FULL = NBK_FULL_MASK;
    const int lane = threadIdx.x & 31;
    const long long B = a.B;
    const double rtol = a.rtol, atol = a.atol;
    const double safety = a.safety, min_factor = a.min_factor, max_factor = a.max_factor;
    const double bound = a.bound;
    // t_bound / upto_t is usually infinite: a kernel-uniform flag skips the guard of every accepted step (DADD + DSETP + LDC)
    const bool bounded = (__double2hiint(a.bound) & 0x7fffffff) < 0x7ff00000;  // finite (an integer test: stays off the FP64 pipe)
    const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;

    // Loop-carried per-lane state: t, h, y, K (K[0] = f at the start of the step, K[S] = f at its
    // end) and the parameters.  The OLD end of the last accepted step (history slot 0) is not
    // carried: it is written to HBM at the moment a trajectory finishes, when it is still live.
    long long idx = -1;
    bool running = false, finishing = false, queue_empty = false;
    double t = 0, h = 0, te_next = 0;
    double y[N], K[S + 1][N], p[NPX];
    int nacc = 0, natt = 0;  // accepted steps / attempts of this call
    int kout = 0;            // RUN: next grid index
    bool fresh = true;       // the next attempt opens a new step (K[S] holds f(t, y))
    int status = NBK_TRAJ_OK;
    int inner = NBK_INNER;   // attempts until the next look at the queue (adapted per warp, NBK_INNER_ADAPT)

    // ---- tail compaction (see NBK_COMPACT above).  A parked trajectory is its loop-carried state: t, h, te_next, y, K,
    // p (NFD doubles), its index and four counters.  The pool holds at most CAP of them, field-major so that the lanes
    // of a warp touch consecutive words.  Everything below the lock is warp-uniform control flow.
    constexpr int NWARP_MAX = NBK_BLOCK / 32;
    constexpr int NFD = 3 + N + (S + 1) * N + NPX;
    constexpr int CAP0 = NBK_POOL_BYTES / ((NFD + 3) * 8);
    constexpr int CAP = CAP0 > 32 * (NWARP_MAX - 1) ? 32 * (NWARP_MAX - 1) : CAP0;
    constexpr bool COMPACT = NBK_COMPACT && !EV && NWARP_MAX > 1 && CAP >= 16;
    constexpr int CAPS = COMPACT ? CAP : 1;
#if NBK_COMPACT
    __shared__ double pool_d[COMPACT ? NFD : 1][CAPS];
    __shared__ long long pool_idx[CAPS];
    __shared__ int pool_i[4][CAPS];
    __shared__ int pool_lock, pool_count, pool_live, warp_idle[NWARP_MAX];
    if constexpr (COMPACT) {
        if (threadIdx.x == 0) pool_lock = 0, pool_count = 0, pool_live = (1 << (blockDim.x >> 5)) - 1;  // bit w: warp w is live
        if (threadIdx.x < NWARP_MAX) warp_idle[threadIdx.x] = 0;
        __syncthreads();
    }
#endif
#ifdef NBK_NEVENTS
    double ev_last[NBK_NEVENTS];  // EV: value of every event function at the last accepted point
#endif

#ifdef NBK_DEBUG_COMPACT
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        a.work_counter[6] = ns;
    }
    unsigned long long dbg_reps = 0;  // attempts (loop iterations) this warp executed after it saw the queue empty
#endif
    for (;;) {
        // ---------------------------------------------------------------- (1) retire
        if (finishing) {
            if (nacc > 0) {
                const bool clean = status != NBK_TRAJ_NONFINITE && status != NBK_TRAJ_MAXATTEMPTS;
                a.ts[B + idx] = t;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    a.ys[(long long)(N + c) * B + idx] = y[c];
                    a.fs[(long long)c * B + idx] = K[0][c];
                    a.fs[(long long)(N + c) * B + idx] = clean ? K[S][c] : K[0][c];
                }
#pragma unroll
                for (int s = 0; s <= S; ++s)
#pragma unroll
                    for (int c = 0; c < N; ++c) a.K[((long long)s * N + c) * B + idx] = K[s][c];
            }
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
            finishing = false;
            idx = -1;
        }
        // ---------------------------------------------------------------- (2) refill
        const unsigned idle = __ballot_sync(FULL, !running);
#if NBK_INNER_ADAPT
        // The number of attempts between two looks at the queue follows the workload (per warp): long trajectories want
        // few looks (vote / branch overhead, refills batched over more lanes), short ones more (a finished lane idles
        // until the next look) -- but never fewer than 8 attempts apart: measured on 29-attempt trajectories (Lorenz to
        // t = 0.5, 1 M) 2..4 apart 0.724 ms, 8 apart 0.664, 16 apart 0.687; on 484-attempt ones 8 / 16: 6.845 / 6.796.
        if (!queue_empty) {
            const int nid = __popc(idle);
            inner = nid >= 4 ? NBK_INNER / 2 : (nid <= 2 ? NBK_INNER : inner);  // (DOP853 / Lorenz, ~75 attempts: 8 apart 6.21 ms, 16 apart 6.27)
        } else {
            inner = NBK_INNER;
        }
#endif
        if (idle) {
            if (queue_empty) {
#if NBK_COMPACT
                if constexpr (COMPACT) {
                    // The queue is empty: from here on trajectories only finish.  Under the CTA's pool lock this warp
                    // (1) fills its idle lanes with parked trajectories, then (2) if it is still not full and the other
                    // live warps of the CTA have room for all of its trajectories, parks them and retires.  Warp 0
                    // never parks and only leaves when it is the last live warp with nothing to do (the pool is empty
                    // then: it would have taken otherwise), so no trajectory is orphaned.
                    const int warp = (int)(threadIdx.x >> 5), nwarp = (int)(blockDim.x >> 5);
                    if (lane == 0)
                        while (atomicCAS(&pool_lock, 0, 1) != 0) {}
                    __syncwarp();
                    __threadfence_block();
                    int cnt = *(volatile int *)&pool_count;
                    int live = *(volatile int *)&pool_live;
                    const int nidle = __popc(idle);
                    const int take = nidle < cnt ? nidle : cnt;
                    const int irank = __popc(idle & ((1u << lane) - 1u));
                    if (!running && irank < take) {
                        const int slot = cnt - 1 - irank;
                        int f = 0;
                        t = pool_d[f++][slot];
                        h = pool_d[f++][slot];
                        te_next = pool_d[f++][slot];
#pragma unroll
                        for (int c = 0; c < N; ++c) y[c] = pool_d[f++][slot];
#pragma unroll
                        for (int q = 0; q <= S; ++q)
#pragma unroll
                            for (int c = 0; c < N; ++c) K[q][c] = pool_d[f++][slot];
#pragma unroll
                        for (int c = 0; c < NPX; ++c) p[c] = pool_d[f++][slot];
                        idx = pool_idx[slot];
                        nacc = pool_i[0][slot];
                        natt = pool_i[1][slot];
                        kout = pool_i[2][slot];
                        fresh = pool_i[3][slot] != 0;
                        status = NBK_TRAJ_OK;
                        running = true;
                    }
                    cnt -= take;
                    const int act = 32 - nidle + take;
                    // Warps retire in a fixed order -- 3, then 2, then 1; warp 0 is the collector that never parks and
                    // leaves last.  Measured on B200 (tools/micro/fp64_smsp.cu): the warps with the SAME index of the four
                    // resident CTAs of an SM sit on four different sub-partitions, warps with different indices may
                    // share one.  With "whoever is last" as the survivor the four survivors of an SM collided on the FP64
                    // pipes and ran at half speed (no gain at all from compacting); with warp 0 as the only survivor
                    // 125 000 trajectories took 1.17 instead of 1.28 ms, with the fixed retirement order 1.12 ms.
                    bool leave = act == 0 && (warp != 0 || live == 1);
                    const bool may_park = warp != 0 && (live >> (warp + 1)) == 0;
                    if (act > 0 && act < 32 && may_park && cnt + act <= CAP) {
                        int room = -cnt;
                        for (int w = 0; w < nwarp; ++w)
                            if (w != warp) room += *(volatile int *)&warp_idle[w];
                        if (room >= act) {
                            const unsigned busy = __ballot_sync(FULL, running);
                            if (running) {
                                const int slot = cnt + __popc(busy & ((1u << lane) - 1u));
                                int f = 0;
                                pool_d[f++][slot] = t;
                                pool_d[f++][slot] = h;
                                pool_d[f++][slot] = te_next;
#pragma unroll
                                for (int c = 0; c < N; ++c) pool_d[f++][slot] = y[c];
#pragma unroll
                                for (int q = 0; q <= S; ++q)
#pragma unroll
                                    for (int c = 0; c < N; ++c) pool_d[f++][slot] = K[q][c];
#pragma unroll
                                for (int c = 0; c < NPX; ++c) pool_d[f++][slot] = p[c];
                                pool_idx[slot] = idx;
                                pool_i[0][slot] = nacc;
                                pool_i[1][slot] = natt;
                                pool_i[2][slot] = kout;
                                pool_i[3][slot] = fresh ? 1 : 0;
                                running = false;
                                idx = -1;
                            }
                            cnt += act;
                            leave = true;
                        }
                    }
                    __syncwarp();
#ifdef NBK_DEBUG_COMPACT
                    if (lane == 0) {  // diagnostics only (tools/compact_debug.py): the Python shell allocates 8 words
                        atomicAdd(a.work_counter + 2, 1ULL);                                   // lock acquisitions
                        atomicAdd(a.work_counter + 3, (unsigned long long)take);               // trajectories taken
                        if (leave && act > 0) atomicAdd(a.work_counter + 4, (unsigned long long)act);  // parked
                        if (leave) atomicAdd(a.work_counter + 5, 1ULL);                        // warps retired here
                    }
#endif
                    if (lane == 0) {
                        warp_idle[warp] = leave ? 0 : 32 - act;
                        pool_count = cnt;
                        pool_live = leave ? (live & ~(1 << warp)) : live;
                        __threadfence_block();
                        atomicExch(&pool_lock, 0);
                    }
                    if (leave) break;
                    if (act == 0) __nanosleep(400);  // the collector waits for the other warps to park or finish
                } else
#endif
                if (idle == FULL) break;
            } else if (__popc(idle) >= NBK_REFILL_MIN) {
                const int nidle = __popc(idle);
                unsigned long long base = 0;
                const int leader = __ffs(idle) - 1;
                if (lane == leader) base = atomicAdd(a.work_counter, (unsigned long long)nidle);
                base = __shfl_sync(FULL, base, leader);
                if ((long long)(base + nidle) >= B) {
#ifdef NBK_DEBUG_COMPACT
                    if (!queue_empty && lane == 0) {
                        unsigned long long ns;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
                        a.work_counter[8 + 4096 + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % 4096] = ns;
                    }
#endif
                    queue_empty = true;
                }
                const long long mine = (long long)base + __popc(idle & ((1u << lane) - 1u));
                if (!running && mine < B) {
                    idx = mine;
                    t = a.ts[B + idx];
                    h = a.h[idx];
#pragma unroll
                    for (int c = 0; c < N; ++c) y[c] = a.ys[(long long)(N + c) * B + idx];
                    double t_old = t, y_old[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) y_old[c] = y[c];
                    if (a.pristine) {
                        // fresh from the init kernel: both history slots hold (t0, y0), K is zero
#pragma unroll
                        for (int s = 0; s <= S; ++s)
#pragma unroll
                            for (int c = 0; c < N; ++c) K[s][c] = 0.0;
                    } else {
                        t_old = a.ts[idx];
#pragma unroll
                        for (int c = 0; c < N; ++c) y_old[c] = a.ys[(long long)c * B + idx];
#pragma unroll
                        for (int s = 0; s <= S; ++s)
#pragma unroll
                            for (int c = 0; c < N; ++c) K[s][c] = a.K[((long long)s * N + c) * B + idx];
                    }
                    load_params<NP>(a, idx, p);
                    nacc = natt = 0;
                    kout = 0;
                    fresh = true;
                    // a trajectory that failed earlier (non-finite / budget) stays failed: the host may
                    // check the status lazily, after queueing further calls
                    status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
                    running = true;
                    if (status != NBK_TRAJ_OK) {
                        running = false;
                    } else if (MODE == NBK_MODE_RUN) {
                        // grid points inside the stored last step are answered from history
                        // without stepping (core.py:401-408)
                        if constexpr (Tab::DOP) {
                            if (kout < a.m && __ldg(a.tgrid + kout) <= t) {
                                double F[7][N];
#pragma unroll
                                for (int k = 0; k < 7; ++k)
#pragma unroll
                                    for (int c = 0; c < N; ++c) F[k][c] = 0.0;
                                if (t_old < t) dop853_prepare<Rhs>(t_old, t, y_old, y, K, p, F);
                                while (kout < a.m && __ldg(a.tgrid + kout) <= t) {
                                    double out[N];
                                    dop853_eval<N>(__ldg(a.tgrid + kout), t_old, t, y_old, y, F, out);
#pragma unroll
                                    for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                                    ++kout;
                                }
                            }
                        } else {
                            while (kout < a.m && __ldg(a.tgrid + kout) <= t) {
                                double out[N];
                                rk_dense<Tab, N>(__ldg(a.tgrid + kout), t_old, t, y_old, y, K, out);
#pragma unroll
                                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                                ++kout;
                            }
                        }
                        running = kout < a.m;
                        te_next = running ? __ldg(a.tgrid + kout) : d_inf();
#ifdef NBK_NEVENTS
                        if constexpr (EV) events_start<N, NPX>(a, idx, t, y, p, ev_last);
#endif
                    } else {
                        running = a.n_steps > 0;
                    }
                    // f(t, y) of the stored state becomes K[0] of the first attempt
#pragma unroll
                    for (int c = 0; c < N; ++c) K[S][c] = a.fs[(long long)(N + c) * B + idx];
                    // t_bound / upto_t guard of the first step (core.py:176-184); later steps are
                    // guarded right after the step before them is accepted
                    if (running && (t + h > bound)) {
                        status = NBK_TRAJ_BOUND;
                        running = false;
                    }
                    finishing = !running;
                    if (status >= NBK_TRAJ_NONFINITE && MODE == NBK_MODE_RUN) {
                        for (int k = 0; k < a.m; ++k)
                            for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + k * a.o_sk + c * a.o_sc] = d_nan();
                    }
                }
            }
        }
        // ---------------------------------------------------------------- (3) attempts
        // NBK_INNER attempts per queue check: a finished lane waits at most NBK_INNER-1 attempts
        // (of ~500 per trajectory) and the vote/branch overhead is paid once per NBK_INNER.
#pragma unroll 1
        for (int rep = 0; rep < inner; ++rep) {
#ifdef NBK_DEBUG_COMPACT
            if (queue_empty) ++dbg_reps;
#endif
            if (running) {
                if (fresh) {  // FSAL: last stage of the accepted step opens the next one
#pragma unroll
                    for (int c = 0; c < N; ++c) K[0][c] = K[S][c];
                }
                double ynew[N], err2;
                rk_attempt<Tab, Rhs>(t, h, y, K, p, ynew, err2, rtol, atol);
                // err2 >= 0 or NaN: err2 < 1  <=>  high word below that of 1.0 (unsigned: a NaN with the
                // sign bit set must not pass)
                const bool accept = (unsigned)__double2hiint(err2) < 0x3ff00000u;
                const double t_new = t + h;
                h *= step_factor<Tab>(err2, safety, min_factor, max_factor, a.x_lo, a.x_hi);
                ++natt;
                bool done = false;
                if (accept) {
                    if (MODE == NBK_MODE_RUN && EV) {
#ifdef NBK_NEVENTS
                        // events first, then the grid points inside the step (core.py:638-645)
                        step_dense<Tab, Rhs> dense(t, t_new, y, ynew, K, p);
                        double t_term = 0.0;
                        if (events_after_step<N, NPX>(a, idx, t, t_new, ynew, p, ev_last, dense, t_term, status)) {
                            if (status != NBK_TRAJ_EVBRACKET) {
                                double out[N];
                                dense(t_term, out);
#pragma unroll
                                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                                a.ev_tterm[idx] = t_term;
                                ++kout;
                                status = NBK_TRAJ_EVENT;
                            }
                            done = true;
                        } else if (t_new >= te_next) {
                            do {
                                double out[N];
                                dense(te_next, out);
#pragma unroll
                                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                                ++kout;
                                te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                            } while (t_new >= te_next);
                            done = kout >= a.m;
                        }
#endif
                    } else if (MODE == NBK_MODE_RUN) {
#ifdef NBK_EXP_NO_EMIT
                        // EXPERIMENT ONLY (tuning variant, never in the shipped library): grid points are counted but
                        // neither evaluated nor stored -- the time of this build is the floor of ANY restructuring of
                        // the emission path (profiles/r02_c3_rk23_emission_bound.json)
                        if (t_new >= te_next) {
                            do {
                                ++kout;
                                te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                            } while (t_new >= te_next);
                            done = kout >= a.m;
                        }
#else
                        if (t_new >= te_next) {
                            // interpolant of this step, built once and evaluated at every grid point inside it
                            double F[Tab::DOP ? 7 : N][Tab::DOP ? N : Tab::M];
                            if constexpr (Tab::DOP) dop853_prepare<Rhs>(t, t_new, y, ynew, K, p, F);
                            else rk_dense_coeffs<Tab, N>(K, F);
                            // 1 / H by the reciprocal seed + two Newton steps (truncation 2^-92: the rounded result is
                            // within an ulp of the IEEE quotient, 5 instructions instead of ~20 with a slow-path call)
                            const double H = t_new - t, r0 = fast_rcp(H);
                            const double inv_H = fma(r0, fma(-H, r0, 1.0), r0);
                            double *dst = a.y_out + (idx * a.o_sb + kout * a.o_sk);
                            // the grid time after the next one is fetched one iteration ahead: the loop condition
                            // never waits for a load (it did: 9 % of the warp samples of C3 sat on that compare)
                            const double *tg = a.tgrid + kout + 1;
                            double te_after = kout + 1 < a.m ? __ldg(tg) : d_inf();
                            do {
                                double out[N];
                                if constexpr (Tab::DOP) dop853_eval<N>(te_next, t, t_new, y, ynew, F, out);
                                else rk_dense_eval_in_step<Tab, N>(te_next, t, t_new, y, ynew, F, out, inv_H);
#pragma unroll
                                for (int c = 0; c < N; ++c) dst[c * a.o_sc] = out[c];
                                dst += a.o_sk;
                                ++kout;
                                ++tg;
                                te_next = te_after;
                                te_after = kout + 1 < a.m ? __ldg(tg) : d_inf();
                            } while (t_new >= te_next);
                            done = kout >= a.m;
                        }
#endif
                    } else {
                        if (a.emit) {
                            a.t_out[idx * a.to_sb + nacc] = t_new;
#pragma unroll
                            for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + nacc * a.o_sk + c * a.o_sc] = ynew[c];
                        }
                        done = nacc + 1 >= a.n_steps;
                    }
                }
                // guards the reference does not have (it would spin forever): non-finite error
                // norm, collapsed step size, attempt budget.
                if (d_nonfinite(err2) || d_collapsed(h)) {
                    status = NBK_TRAJ_NONFINITE;
                    done = true;
                } else if (!done && natt >= max_att) {
                    status = NBK_TRAJ_MAXATTEMPTS;
                    done = true;
                }
                // t_bound / upto_t guard of the NEXT step, with the proposed h (core.py:176-184)
                if (bounded && accept && !done && (t_new + h > bound)) {
                    status = NBK_TRAJ_BOUND;
                    done = true;
                }
                if (done) {
                    // history slot 0 = start of the last accepted step.  After an accepted attempt that is
                    // the pre-commit (t, y); after an abnormal stop on a rejected attempt the window
                    // degenerates to the current point.
                    a.ts[idx] = t;
#pragma unroll
                    for (int c = 0; c < N; ++c) a.ys[(long long)c * B + idx] = y[c];
                    running = false;
                    finishing = true;
                }
                // commit (predicated moves, no divergence)
                if (accept) {
                    t = t_new;
#pragma unroll
                    for (int c = 0; c < N; ++c) y[c] = ynew[c];
                    ++nacc;
                }
                fresh = accept;
            }
        }
    }
#ifdef NBK_DEBUG_COMPACT
    if (lane == 0) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        a.work_counter[8 + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % 4096] = ns;
        a.work_counter[8 + 8192 + (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % 4096] = dbg_reps;
    }
#endif
}

// =====================================================================================
// Adams-Bashforth-k advance (ForwardEuler == k=1).  All trajectories take (nearly) the same
// number of steps, so the mapping is static: thread i owns trajectory i and consecutive
// lanes read/write consecutive addresses of every SoA array.
// =====================================================================================
template <int L, int N>
NBK_DEV void hermite_window(double te, const double (&ts)[L], const double (&ys)[L][N], const double (&fs)[L][N],
                            double (&out)[N]) {
    const double t0 = ts[0];
    if (te == t0) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = ys[0][i];
        return;
    }
    const double dt = ts[L - 1] - t0;
    const double T = (te - t0) / dt;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double dy = ys[L - 1][i] - ys[0][i];
        out[i] = ys[0][i] + T * dy + T * (T - 1) * ((1 - 2 * T) * dy + dt * ((T - 1) * fs[0][i] + T * fs[L - 1][i]));
    }
}

// The same with the two end points of the window given explicitly (the ring-buffered advance loops below keep the
// window rotated in registers) and the quotient by fast_div (T within ~1 ulp of the reference's division).
// PAST_T0: the caller knows te > t0 (the emission loops: every grid point up to the start of the step was written by an
// earlier step), so the reference's exact return at the window start is dead code there.
template <int N, bool PAST_T0 = false>
NBK_DEV void hermite_ends(double te, double t0, const double (&y0)[N], const double (&f0)[N], double t1, const double (&y1)[N],
                          const double (&f1)[N], double (&out)[N]) {
    if (!PAST_T0 && te == t0) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = y0[i];
        return;
    }
    const double dt = t1 - t0;
    const double T = fast_div(te - t0, dt);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double dy = y1[i] - y0[i];
        out[i] = y0[i] + T * dy + T * (T - 1) * ((1 - 2 * T) * dy + dt * ((T - 1) * f0[i] + T * f1[i]));
    }
}

// history window -> HBM, logical slot l living in physical slot (l + ROT) % L
template <int L, int N, int ROT>
NBK_DEV void store_window(const nbk_kargs &a, long long idx, const double (&ts)[L], const double (&ys)[L][N],
                          const double (&fs)[L][N]) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
        a.ts[(long long)l * a.B + idx] = ts[(l + ROT) % L];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            a.ys[((long long)l * N + c) * a.B + idx] = ys[(l + ROT) % L][c];
            a.fs[((long long)l * N + c) * a.B + idx] = fs[(l + ROT) % L][c];
        }
    }
}

#ifdef NBK_NEVENTS
// dense output of the fixed-step solvers (window Hermite) as the functor events_after_step takes
template <int L, int N>
struct window_dense {
    const double (&ts)[L];
    const double (&ys)[L][N];
    const double (&fs)[L][N];
    NBK_DEV window_dense(const double (&t_)[L], const double (&y_)[L][N], const double (&f_)[L][N]) : ts(t_), ys(y_), fs(f_) {}
    NBK_DEV void operator()(double te, double (&out)[N]) { hermite_window<L, N>(te, ts, ys, fs, out); }
};
#endif

template <int ORDER, class Rhs, int MODE, bool EV = false>
NBK_DEV void advance_ab(const nbk_kargs &a) {
    constexpr int N = Rhs::N, NP = Rhs::NP, NPX = NP > 0 ? NP : 1, L = tab_ab<ORDER>::LEN;
    const long long B = a.B;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B) return;
    double ts[L], ys[L][N], fs[L][N], p[NPX];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        ts[l] = a.ts[(long long)l * B + idx];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            ys[l][c] = a.ys[((long long)l * N + c) * B + idx];
            fs[l][c] = a.fs[((long long)l * N + c) * B + idx];
        }
    }
    const double h = a.h[idx];
    load_params<NP>(a, idx, p);
    double hB[ORDER];
#pragma unroll
    for (int j = 0; j < ORDER; ++j) hB[j] = h * tab_ab<ORDER>::b(j);

    int kout = 0, status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
    long long nsteps = 0;
    double te_next = d_inf();
    bool running;
#ifdef NBK_NEVENTS
    double ev_last[NBK_NEVENTS];
#endif
    if (status != NBK_TRAJ_OK) {
        // failed earlier (start-up): stays failed, outputs are NaN
        running = false;
        if (MODE == NBK_MODE_RUN)
            for (int k = 0; k < a.m; ++k)
                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + k * a.o_sk + c * a.o_sc] = d_nan();
    } else if (MODE == NBK_MODE_RUN) {
        while (kout < a.m && __ldg(a.tgrid + kout) <= ts[L - 1]) {
            double out[N];
            hermite_window<L, N>(__ldg(a.tgrid + kout), ts, ys, fs, out);
#pragma unroll
            for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
            ++kout;
        }
        running = kout < a.m;
        if (running) te_next = __ldg(a.tgrid + kout);
#ifdef NBK_NEVENTS
        if constexpr (EV) events_start<N, NPX>(a, idx, ts[L - 1], ys[L - 1], p, ev_last);
#endif
    } else {
        running = a.n_steps > 0;
    }
    if constexpr (!EV) {
        // Ring-buffered loop: L consecutive steps are unrolled with the window rotating through the same registers
        // (logical slot l of rotation r is physical slot (l + r) % L), so a push overwrites the oldest slot instead
        // of moving L - 1 slots (AdamsBashforth4 on n = 2: 30 register moves per step in the shifting version).
        // A trajectory that stops inside a block SKIPS the block's remaining steps instead of leaving the loop: control
        // flow rejoins right behind each step, so no values have to be marshalled into fixed registers for a far
        // exit (the `break` version paid 10 register moves per step for that); `rot` remembers the rotation the
        // window was left in.  Step counters are 32-bit (the budget is clamped like in the adaptive kernels).
        int nst = 0, rot = 0;
        const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;
        const int n_steps = a.n_steps > 2147483647LL ? 2147483647 : (int)a.n_steps;
        double *dst = MODE == NBK_MODE_RUN ? a.y_out + (idx * a.o_sb + kout * a.o_sk) : a.y_out + idx * a.o_sb;
        const double *tg = a.tgrid + kout + 1;  // RUN: the grid time after te_next, fetched one point ahead
        double te_after = (MODE == NBK_MODE_RUN && running && kout + 1 < a.m) ? __ldg(tg) : d_inf();
        while (running) {
#pragma unroll
            for (int r = 0; r < L; ++r) {
                const int newest = (L - 1 + r) % L, oldest = r % L;
                const double t_new = ts[newest] + h;
                if (running && t_new > a.bound) {  // guard of this step (core.py:176-184)
                    status = NBK_TRAJ_BOUND;
                    running = false;
                }
                if (running) {
                    double ynew[N], fnew[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        double acc = hB[0] * fs[(L - ORDER + r) % L][c];
#pragma unroll
                        for (int j = 1; j < ORDER; ++j) acc = fma(hB[j], fs[(L - ORDER + j + r) % L][c], acc);
                        ynew[c] = acc + ys[newest][c];
                    }
                    Rhs::eval(t_new, ynew, p, fnew);
                    ts[oldest] = t_new;  // push: the oldest slot becomes the newest one of rotation r + 1
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        ys[oldest][c] = ynew[c];
                        fs[oldest][c] = fnew[c];
                    }
                    ++nst;
                    rot = (r + 1) % L;
                    if (MODE == NBK_MODE_RUN) {
                        while (t_new >= te_next) {
                            double out[N];
                            hermite_ends<N, true>(te_next, ts[(r + 1) % L], ys[(r + 1) % L], fs[(r + 1) % L], t_new, ynew, fnew, out);
#pragma unroll
                            for (int c = 0; c < N; ++c) dst[c * a.o_sc] = out[c];
                            dst += a.o_sk;
                            ++kout;
                            ++tg;
                            te_next = te_after;
                            te_after = kout + 1 < a.m ? __ldg(tg) : d_inf();
                        }
                        running = kout < a.m;
                    } else {
                        if (a.emit) {
                            a.t_out[idx * a.to_sb + (nst - 1)] = t_new;
#pragma unroll
                            for (int c = 0; c < N; ++c) dst[c * a.o_sc] = ynew[c];
                            dst += a.o_sk;
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
                case 0: store_window<L, N, 0>(a, idx, ts, ys, fs); break;
                case 1: store_window<L, N, 1 % L>(a, idx, ts, ys, fs); break;
                case 2: store_window<L, N, 2 % L>(a, idx, ts, ys, fs); break;
                case 3: store_window<L, N, 3 % L>(a, idx, ts, ys, fs); break;
                default: store_window<L, N, 4 % L>(a, idx, ts, ys, fs); break;
            }
            a.n_acc[idx] += nst;
        }
        store_status(a, idx, status);
        a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : nst;
        return;
    }
    while (running) {
        if (ts[L - 1] + h > a.bound) {
            status = NBK_TRAJ_BOUND;
            break;
        }
        const double t_new = ts[L - 1] + h;
        double ynew[N], fnew[N];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double acc = hB[0] * fs[L - ORDER][c];
#pragma unroll
            for (int j = 1; j < ORDER; ++j) acc = fma(hB[j], fs[L - ORDER + j][c], acc);
            ynew[c] = acc + ys[L - 1][c];
        }
        Rhs::eval(t_new, ynew, p, fnew);
        // push: shift the window by one (register renaming after unrolling)
#pragma unroll
        for (int l = 0; l < L - 1; ++l) {
            ts[l] = ts[l + 1];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                ys[l][c] = ys[l + 1][c];
                fs[l][c] = fs[l + 1][c];
            }
        }
        ts[L - 1] = t_new;
#pragma unroll
        for (int c = 0; c < N; ++c) {
            ys[L - 1][c] = ynew[c];
            fs[L - 1][c] = fnew[c];
        }
        ++nsteps;
#ifdef NBK_NEVENTS
        if constexpr (EV && MODE == NBK_MODE_RUN) {
            // the bisection brackets the previous point ts[L-2] and the new one; for L > 2 the window Hermite
            // does not pass through ts[L-2], which is where the reference's bisect raises ValueError
            window_dense<L, N> dense(ts, ys, fs);
            double t_term = 0.0;
            if (events_after_step<N, NPX>(a, idx, ts[L - 2], t_new, ynew, p, ev_last, dense, t_term, status)) {
                if (status != NBK_TRAJ_EVBRACKET) {
                    double out[N];
                    dense(t_term, out);
#pragma unroll
                    for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                    a.ev_tterm[idx] = t_term;
                    ++kout;
                    status = NBK_TRAJ_EVENT;
                }
                break;
            }
        }
#endif
        if (MODE == NBK_MODE_RUN) {
            if (t_new >= te_next) {
                double *dst = a.y_out + (idx * a.o_sb + kout * a.o_sk);
                do {
                    double out[N];
                    hermite_ends<N>(te_next, ts[0], ys[0], fs[0], ts[L - 1], ys[L - 1], fs[L - 1], out);
#pragma unroll
                    for (int c = 0; c < N; ++c) dst[c * a.o_sc] = out[c];
                    dst += a.o_sk;
                    ++kout;
                    te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                } while (t_new >= te_next);
            }
            running = kout < a.m;
        } else {
            if (a.emit) {
                a.t_out[idx * a.to_sb + (nsteps - 1)] = t_new;
#pragma unroll
                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + (nsteps - 1) * a.o_sk + c * a.o_sc] = ynew[c];
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
        for (int l = 0; l < L; ++l) {
            a.ts[(long long)l * B + idx] = ts[l];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                a.ys[((long long)l * N + c) * B + idx] = ys[l][c];
                a.fs[((long long)l * N + c) * B + idx] = fs[l][c];
            }
        }
        a.n_acc[idx] += nsteps;
    }
    store_status(a, idx, status);
    a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : (int)nsteps;
}

// =====================================================================================
// Implicit linear multistep (AdamsMoulton1..5 / BackwardEuler, BDF1..6): the advance_ab driver with the
// implicit step of nbkode/multistep/core.py:115-170.  The root of  g(y) = y - h Bn f(t_new, y) - Kc  is found
// exactly as the reference does it (nbkode/nbcompat/zeros.py): for n > 1 Newton's method with a central
// finite-difference Jacobian (eps = 1e-10) and an LU solve with partial pivoting, started from the current
// y, stopped when ||g||_2 <= 1.48e-8, at most 50 iterations (newton_hd, :460-505); for n == 1 the secant
// method of scipy's `newton` (:350-392) started from y and y (1 + 1e-4) +- 1e-4, stopped when the iterate
// moves by <= 1.48e-8.  A step that does not converge fails the trajectory (the reference raises RuntimeError).
// =====================================================================================
template <int N>
NBK_DEV bool lu_solve(double (&J)[N][N], double (&r)[N]) {  // J x = r, result in r (np.linalg.solve == LAPACK dgesv)
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double best = fabs(J[k][k]);
#pragma unroll
        for (int q = k + 1; q < N; ++q) {
            const double v = fabs(J[q][k]);
            if (v > best) {
                best = v;
                piv = q;
            }
        }
        ok = ok && best > 0.0;
#pragma unroll
        for (int q = k + 1; q < N; ++q) {  // row interchange k <-> piv with selects (all indices stay static)
            const bool sw = piv == q;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                const double up = J[k][c], lo = J[q][c];
                J[k][c] = sw ? lo : up;
                J[q][c] = sw ? up : lo;
            }
            const double up = r[k], lo = r[q];
            r[k] = sw ? lo : up;
            r[q] = sw ? up : lo;
        }
        const double inv = 1.0 / J[k][k];
#pragma unroll
        for (int q = k + 1; q < N; ++q) {
            const double m = J[q][k] * inv;
#pragma unroll
            for (int c = k + 1; c < N; ++c) J[q][c] = fma(-m, J[k][c], J[q][c]);
            r[q] = fma(-m, r[k], r[q]);
        }
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        double acc = r[k];
#pragma unroll
        for (int c = k + 1; c < N; ++c) acc = fma(-J[k][c], r[c], acc);
        r[k] = acc / J[k][k];
    }
    return ok;
}

template <class Rhs>
NBK_DEV void implicit_root(double t, double hBn, const double (&x)[Rhs::N], const double (&Kc)[Rhs::N],
                           const double (&p)[Rhs::NP > 0 ? Rhs::NP : 1], double (&g)[Rhs::N]) {
    double f[Rhs::N];
    Rhs::eval(t, x, p, f);
#pragma unroll
    for (int i = 0; i < Rhs::N; ++i) g[i] = x[i] - hBn * f[i] - Kc[i];  // y - h_Bn * rhs(t, y) - K  (core.py:121-122)
}

// returns false if the iteration does not converge (or meets a singular Jacobian / non-finite values)
template <class Rhs>
NBK_DEV bool implicit_solve(double t, double hBn, const double (&Kc)[Rhs::N], const double (&p)[Rhs::NP > 0 ? Rhs::NP : 1],
                            double (&y)[Rhs::N]) {
    constexpr int N = Rhs::N;
    const double tol = 1.48e-8;
    if constexpr (N == 1) {
        // secant method (zeros.py:350-392 with fprime = None, x1 = None)
        double p0[1] = {1.0 * y[0]}, p1[1], q0[1], q1[1];
        const double eps = 1e-4;
        p1[0] = y[0] * (1 + eps);
        p1[0] += p1[0] >= 0 ? eps : -eps;
        implicit_root<Rhs>(t, hBn, p0, Kc, p, q0);
        implicit_root<Rhs>(t, hBn, p1, Kc, p, q1);
        if (fabs(q1[0]) < fabs(q0[0])) {
            double tmp = p0[0]; p0[0] = p1[0]; p1[0] = tmp;
            tmp = q0[0]; q0[0] = q1[0]; q1[0] = tmp;
        }
        for (int it = 0; it < 50; ++it) {
            double pn;
            if (q1[0] == q0[0]) {
                if (p1[0] != p0[0]) return false;  // "tolerance reached": j_newton raises RuntimeError
                y[0] = (p1[0] + p0[0]) / 2.0;
                return true;
            }
            if (fabs(q1[0]) > fabs(q0[0])) pn = (-q0[0] / q1[0] * p1[0] + p0[0]) / (1 - q0[0] / q1[0]);
            else pn = (-q1[0] / q0[0] * p0[0] + p1[0]) / (1 - q1[0] / q0[0]);
            if (fabs(pn - p1[0]) <= tol) {
                y[0] = pn;
                return true;
            }
            p0[0] = p1[0];
            q0[0] = q1[0];
            p1[0] = pn;
            implicit_root<Rhs>(t, hBn, p1, Kc, p, q1);
        }
        return false;
    } else {
        double val[N];
        implicit_root<Rhs>(t, hBn, y, Kc, p, val);
        int it = 0;
        for (;;) {
            double d2 = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) d2 = fma(val[i], val[i], d2);
            if (sqrt(d2) <= tol) return true;  // isclose(distance, 0, rtol=0, atol=1.48e-8); NaN never passes
            const double eps = 1e-10;
            double J[N][N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double x1[N], x2[N], f1[N], f2[N];
#pragma unroll
                for (int c = 0; c < N; ++c) x1[c] = x2[c] = y[c];
                x1[i] += eps;
                x2[i] -= eps;
                implicit_root<Rhs>(t, hBn, x1, Kc, p, f1);
                implicit_root<Rhs>(t, hBn, x2, Kc, p, f2);
#pragma unroll
                for (int c = 0; c < N; ++c) J[c][i] = (f1[c] - f2[c]) / (2 * eps);
            }
            double delta[N];
#pragma unroll
            for (int i = 0; i < N; ++i) delta[i] = -val[i];
            if (!lu_solve<N>(J, delta)) return false;
#pragma unroll
            for (int i = 0; i < N; ++i) y[i] = y[i] + delta[i];
            implicit_root<Rhs>(t, hBn, y, Kc, p, val);
            if (++it > 50) return false;  // iteration_counter > maxiter (zeros.py:487-488), checked after the update
        }
    }
}

template <int ID, class Rhs, int MODE, bool EV = false>
NBK_DEV void advance_im(const nbk_kargs &a) {
    using Tab = tab_im<ID>;
    constexpr int N = Rhs::N, NP = Rhs::NP, NPX = NP > 0 ? NP : 1, L = Tab::LEN, KK = Tab::K;
    const long long B = a.B;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B) return;
    double ts[L], ys[L][N], fs[L][N], p[NPX];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        ts[l] = a.ts[(long long)l * B + idx];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            ys[l][c] = a.ys[((long long)l * N + c) * B + idx];
            fs[l][c] = a.fs[((long long)l * N + c) * B + idx];
        }
    }
    const double h = a.h[idx];
    load_params<NP>(a, idx, p);
    const double hBn = h * Tab::bn();
    int kout = 0, status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
    long long nsteps = 0;
    double te_next = d_inf();
    bool running;
#ifdef NBK_NEVENTS
    double ev_last[NBK_NEVENTS];
#endif
    if (status != NBK_TRAJ_OK) {
        running = false;
        if (MODE == NBK_MODE_RUN)
            for (int k = 0; k < a.m; ++k)
                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + k * a.o_sk + c * a.o_sc] = d_nan();
    } else if (MODE == NBK_MODE_RUN) {
        while (kout < a.m && __ldg(a.tgrid + kout) <= ts[L - 1]) {
            double out[N];
            hermite_window<L, N>(__ldg(a.tgrid + kout), ts, ys, fs, out);
#pragma unroll
            for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
            ++kout;
        }
        running = kout < a.m;
        if (running) te_next = __ldg(a.tgrid + kout);
#ifdef NBK_NEVENTS
        if constexpr (EV) events_start<N, NPX>(a, idx, ts[L - 1], ys[L - 1], p, ev_last);
#endif
    } else {
        running = a.n_steps > 0;
    }
    while (running) {
        if (ts[L - 1] + h > a.bound) {
            status = NBK_TRAJ_BOUND;
            break;
        }
        const double t_new = ts[L - 1] + h;
        double Kc[N], ynew[N], fnew[N];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            // K = h*B @ fs[-K:] - A @ ys[-K:]   (core.py:133, :150)
            double hb = 0.0, ay = 0.0;
#pragma unroll
            for (int j = 0; j < KK; ++j) {
                if (Tab::b(j) != 0.0) hb = fma(h * Tab::b(j), fs[L - KK + j][c], hb);
                if (Tab::a(j) != 0.0) ay = fma(Tab::a(j), ys[L - KK + j][c], ay);
            }
            Kc[c] = hb - ay;
            ynew[c] = ys[L - 1][c];  // Newton / secant start from the current state
        }
        if (!implicit_solve<Rhs>(t_new, hBn, Kc, p, ynew)) {
            status = NBK_TRAJ_NEWTON;
            break;
        }
        Rhs::eval(t_new, ynew, p, fnew);
#pragma unroll
        for (int l = 0; l < L - 1; ++l) {
            ts[l] = ts[l + 1];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                ys[l][c] = ys[l + 1][c];
                fs[l][c] = fs[l + 1][c];
            }
        }
        ts[L - 1] = t_new;
#pragma unroll
        for (int c = 0; c < N; ++c) {
            ys[L - 1][c] = ynew[c];
            fs[L - 1][c] = fnew[c];
        }
        ++nsteps;
#ifdef NBK_NEVENTS
        if constexpr (EV && MODE == NBK_MODE_RUN) {
            window_dense<L, N> dense(ts, ys, fs);
            double t_term = 0.0;
            if (events_after_step<N, NPX>(a, idx, ts[L - 2], t_new, ynew, p, ev_last, dense, t_term, status)) {
                if (status != NBK_TRAJ_EVBRACKET) {
                    double out[N];
                    dense(t_term, out);
#pragma unroll
                    for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                    a.ev_tterm[idx] = t_term;
                    ++kout;
                    status = NBK_TRAJ_EVENT;
                }
                break;
            }
        }
#endif
        if (MODE == NBK_MODE_RUN) {
            if (t_new >= te_next) {
                double *dst = a.y_out + (idx * a.o_sb + kout * a.o_sk);
                do {
                    double out[N];
                    hermite_ends<N>(te_next, ts[0], ys[0], fs[0], ts[L - 1], ys[L - 1], fs[L - 1], out);
#pragma unroll
                    for (int c = 0; c < N; ++c) dst[c * a.o_sc] = out[c];
                    dst += a.o_sk;
                    ++kout;
                    te_next = kout < a.m ? __ldg(a.tgrid + kout) : d_inf();
                } while (t_new >= te_next);
            }
            running = kout < a.m;
        } else {
            if (a.emit) {
                a.t_out[idx * a.to_sb + (nsteps - 1)] = t_new;
#pragma unroll
                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + (nsteps - 1) * a.o_sk + c * a.o_sc] = ynew[c];
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
        for (int l = 0; l < L; ++l) {
            a.ts[(long long)l * B + idx] = ts[l];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                a.ys[((long long)l * N + c) * B + idx] = ys[l][c];
                a.fs[((long long)l * N + c) * B + idx] = fs[l][c];
            }
        }
        a.n_acc[idx] += nsteps;
    }
    store_status(a, idx, status);
    a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : (int)nsteps;
}

// =====================================================================================
// Construction kernels (one thread per trajectory; cold path: IEEE division / sqrt / pow)
// =====================================================================================
template <class Rhs>
NBK_DEV void load_initial(const nbk_kargs &a, long long idx, double (&y)[Rhs::N],
                          double (&p)[Rhs::NP > 0 ? Rhs::NP : 1]) {
    constexpr int N = Rhs::N, NP = Rhs::NP;
#pragma unroll
    for (int c = 0; c < N; ++c) y[c] = a.y0[idx * a.y0_sb + c * a.y0_sc];
    if (NP == 0) p[0] = 0.0;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        if (a.params_shared) {
            p[c] = a.params[c];  // the host already placed the shared vector
        } else {
            p[c] = a.p0[idx * a.p0_sb + c * a.p0_sc];
            a.params[(long long)c * a.B + idx] = p[c];  // SoA copy used by the advance kernels
        }
    }
}

template <class Tab, class Rhs>
NBK_DEV void init_adaptive(const nbk_kargs &a) {
    constexpr int N = Rhs::N, NPX = Rhs::NP > 0 ? Rhs::NP : 1, S = Tab::S;
    const long long B = a.B;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B) return;
    double y[N], f[N], p[NPX];
    load_initial<Rhs>(a, idx, y, p);
    Rhs::eval(a.t0, y, p, f);
    double h = a.h0 ? a.h0[idx] : a.h_init;
    if (!(h == h)) h = initial_step<Rhs>(a.t0, y, f, p, Tab::Q, a.rtol, a.atol);
    a.ts[idx] = a.t0;
    a.ts[B + idx] = a.t0;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        a.ys[(long long)c * B + idx] = y[c];
        a.ys[(long long)(N + c) * B + idx] = y[c];
        a.fs[(long long)c * B + idx] = f[c];
        a.fs[(long long)(N + c) * B + idx] = f[c];
    }
#pragma unroll
    for (int s = 0; s <= S; ++s)
#pragma unroll
        for (int c = 0; c < N; ++c) a.K[((long long)s * N + c) * B + idx] = 0.0;
    a.h[idx] = h;
    a.n_acc[idx] = 0;
    a.n_rej[idx] = 0;
    a.status[idx] = NBK_TRAJ_OK;
}

// Start-up points of an adaptive first stepper with DEFAULT tolerances (rtol 1e-3, atol 1e-6):
// dense output at the absolute times hfix, 2 hfix, ..., npts*hfix (multistep/core.py:60-65).
template <class Tab, class Rhs>
NBK_DEV bool rk_startup_points(double t0, const double (&y0)[Rhs::N], const double (&f0)[Rhs::N],
                               const double (&p)[Rhs::NP > 0 ? Rhs::NP : 1], double hfix, int npts, double *pt,
                               double (*py)[Rhs::N], double (*pf)[Rhs::N]) {
    constexpr int N = Rhs::N, S = Tab::S;
    const double rtol = 1e-3, atol = 1e-6;
    double t = t0, t_old = t0, y[N], y_old[N], K[S + 1][N];
#pragma unroll
    for (int c = 0; c < N; ++c) {
        y[c] = y_old[c] = y0[c];
#pragma unroll
        for (int s = 0; s <= S; ++s) K[s][c] = 0.0;
        K[S][c] = f0[c];
    }
    double h = initial_step<Rhs>(t0, y0, f0, p, Tab::Q, rtol, atol);
    bool fresh = true;
    int budget = 1000000;
    for (int j = 1; j <= npts; ++j) {
        const double ti = j * hfix;
        while (t < ti) {
            if (fresh) {
#pragma unroll
                for (int c = 0; c < N; ++c) K[0][c] = K[S][c];
            }
            double ynew[N], err2;
            rk_attempt<Tab, Rhs>(t, h, y, K, p, ynew, err2, rtol, atol);
            const bool accept = pos_less(err2, 1.0);
            const double t_new = t + h;
            h *= step_factor<Tab>(err2, 0.9, 0.2, 10.0, 1e-13, 1e8);
            if (accept) {
                t_old = t;
                t = t_new;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    y_old[c] = y[c];
                    y[c] = ynew[c];
                }
            }
            fresh = accept;
            if (d_nonfinite(err2) || d_collapsed(h) || --budget < 0) return false;
        }
        double out[N];
        rk_dense<Tab, N>(ti, t_old, t, y_old, y, K, out);
        pt[j - 1] = ti;
#pragma unroll
        for (int c = 0; c < N; ++c) py[j - 1][c] = out[c];
        Rhs::eval(ti, out, p, pf[j - 1]);
    }
    return true;
}

template <int ORDER, class Rhs>
NBK_DEV void init_ab(const nbk_kargs &a) {
    constexpr int N = Rhs::N, NPX = Rhs::NP > 0 ? Rhs::NP : 1, L = tab_ab<ORDER>::LEN;
    const long long B = a.B;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B) return;
    double y[N], f[N], p[NPX];
    load_initial<Rhs>(a, idx, y, p);
    Rhs::eval(a.t0, y, p, f);
    const double h = a.h0 ? a.h0[idx] : a.h_init;
    // history: L copies of (t0, y0, f0) (buffer.py:14-19), then ORDER-1 start-up points pushed
    double pt[5], py[5][N], pf[5][N];
    int npush = 0, status = NBK_TRAJ_OK;
    const int fsr = a.first_stepper;
    if (ORDER > 1 && fsr != 0) {
        bool ok = true;
        if (fsr == 45) {
            ok = rk_startup_points<tab_rk45, Rhs>(a.t0, y, f, p, h, ORDER - 1, pt, py, pf);
        } else if (fsr == 23) {
            ok = rk_startup_points<tab_rk23, Rhs>(a.t0, y, f, p, h, ORDER - 1, pt, py, pf);
        } else {
            // fixed-step first stepper AdamsBashforth-k' built with ITS default start-up
            // (RungeKutta45), then ORDER-1 steps of it (multistep/core.py:54-59)
            const int k2 = fsr, L2 = k2 > 2 ? k2 : 2;
            double st[5], sy[5][N], sf[5][N];
            for (int l = 0; l < L2; ++l) {
                st[l] = a.t0;
                for (int c = 0; c < N; ++c) {
                    sy[l][c] = y[c];
                    sf[l][c] = f[c];
                }
            }
            if (k2 > 1) {
                double qt[4], qy[4][N], qf[4][N];
                ok = rk_startup_points<tab_rk45, Rhs>(a.t0, y, f, p, h, k2 - 1, qt, qy, qf);
                for (int j = 0; j < k2 - 1; ++j) {
                    for (int l = 0; l < L2 - 1; ++l) {
                        st[l] = st[l + 1];
                        for (int c = 0; c < N; ++c) {
                            sy[l][c] = sy[l + 1][c];
                            sf[l][c] = sf[l + 1][c];
                        }
                    }
                    st[L2 - 1] = qt[j];
                    for (int c = 0; c < N; ++c) {
                        sy[L2 - 1][c] = qy[j][c];
                        sf[L2 - 1][c] = qf[j][c];
                    }
                }
            }
            for (int j = 0; j < ORDER - 1; ++j) {
                const double t_new = st[L2 - 1] + h;
                double yn[N], fn[N];
                for (int c = 0; c < N; ++c) {
                    double acc = (h * ab_weight(k2, 0)) * sf[L2 - k2][c];
                    for (int i = 1; i < k2; ++i) acc = fma(h * ab_weight(k2, i), sf[L2 - k2 + i][c], acc);
                    yn[c] = acc + sy[L2 - 1][c];
                }
                Rhs::eval(t_new, yn, p, fn);
                for (int l = 0; l < L2 - 1; ++l) {
                    st[l] = st[l + 1];
                    for (int c = 0; c < N; ++c) {
                        sy[l][c] = sy[l + 1][c];
                        sf[l][c] = sf[l + 1][c];
                    }
                }
                st[L2 - 1] = t_new;
                pt[j] = t_new;
                for (int c = 0; c < N; ++c) {
                    sy[L2 - 1][c] = yn[c];
                    sf[L2 - 1][c] = fn[c];
                    py[j][c] = yn[c];
                    pf[j][c] = fn[c];
                }
            }
        }
        npush = ORDER - 1;
        if (!ok) status = NBK_TRAJ_NONFINITE;
    }
    // window after npush pushes: (L - npush) copies of the initial point, then the pushed ones
    for (int l = 0; l < L; ++l) {
        const int j = l - (L - npush);
        a.ts[(long long)l * B + idx] = j < 0 ? a.t0 : pt[j];
        for (int c = 0; c < N; ++c) {
            a.ys[((long long)l * N + c) * B + idx] = j < 0 ? y[c] : py[j][c];
            a.fs[((long long)l * N + c) * B + idx] = j < 0 ? f[c] : pf[j][c];
        }
    }
    a.h[idx] = h;
    a.n_acc[idx] = 0;
    a.n_rej[idx] = 0;
    store_status(a, idx, status);
}

// =====================================================================================
// Fixed-step explicit Runge-Kutta (Runge2, Runge3, Heun3, RungeKutta4, RungeKutta3_8):
// stage loop nbkode/runge_kutta/explicit.py:28-46, step runge_kutta/core.py:33-42, dense output
// = Hermite over the last step (core.py:577-596 with LEN_HISTORY = 2).  Static mapping like the
// Adams-Bashforth kernels: all trajectories take the same number of steps.
// =====================================================================================
template <int ID, class Rhs>
NBK_DEV void init_erk(const nbk_kargs &a) {
    constexpr int N = Rhs::N, NPX = Rhs::NP > 0 ? Rhs::NP : 1, S = tab_erk<ID>::S;
    const long long B = a.B;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B) return;
    double y[N], f[N], p[NPX];
    load_initial<Rhs>(a, idx, y, p);
    Rhs::eval(a.t0, y, p, f);
    a.ts[idx] = a.t0;
    a.ts[B + idx] = a.t0;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        a.ys[(long long)c * B + idx] = y[c];
        a.ys[(long long)(N + c) * B + idx] = y[c];
        a.fs[(long long)c * B + idx] = f[c];
        a.fs[(long long)(N + c) * B + idx] = f[c];
#pragma unroll
        for (int s = 0; s < S; ++s) a.K[((long long)s * N + c) * B + idx] = 0.0;
    }
    a.h[idx] = a.h0 ? a.h0[idx] : a.h_init;
    a.n_acc[idx] = 0;
    a.n_rej[idx] = 0;
    a.status[idx] = NBK_TRAJ_OK;
}

template <int ID, class Rhs, int MODE, bool EV = false>
NBK_DEV void advance_erk(const nbk_kargs &a) {
    using Tab = tab_erk<ID>;
    constexpr int N = Rhs::N, NP = Rhs::NP, NPX = NP > 0 ? NP : 1, S = Tab::S;
    const long long B = a.B;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B) return;
    double ts[2], ys[2][N], fs[2][N], K[S][N], p[NPX];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        ts[l] = a.ts[(long long)l * B + idx];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            ys[l][c] = a.ys[((long long)l * N + c) * B + idx];
            fs[l][c] = a.fs[((long long)l * N + c) * B + idx];
        }
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int c = 0; c < N; ++c) K[s][c] = 0.0;
    const double h = a.h[idx];
    load_params<NP>(a, idx, p);
    int kout = 0, status = a.status[idx] >= NBK_TRAJ_NONFINITE ? a.status[idx] : NBK_TRAJ_OK;
    long long nsteps = 0;
    double te_next = d_inf();
    bool running;
#ifdef NBK_NEVENTS
    double ev_last[NBK_NEVENTS];
#endif
    if (status != NBK_TRAJ_OK) {
        running = false;
        if (MODE == NBK_MODE_RUN)
            for (int k = 0; k < a.m; ++k)
                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + k * a.o_sk + c * a.o_sc] = d_nan();
    } else if (MODE == NBK_MODE_RUN) {
        while (kout < a.m && __ldg(a.tgrid + kout) <= ts[1]) {
            double out[N];
            hermite_window<2, N>(__ldg(a.tgrid + kout), ts, ys, fs, out);
#pragma unroll
            for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
            ++kout;
        }
        running = kout < a.m;
        if (running) te_next = __ldg(a.tgrid + kout);
#ifdef NBK_NEVENTS
        if constexpr (EV) events_start<N, NPX>(a, idx, ts[1], ys[1], p, ev_last);
#endif
    } else {
        running = a.n_steps > 0;
    }
    if constexpr (!EV) {
        // Two steps unrolled with the two history slots swapping roles (a push overwrites the older point: no window
        // shift), f of the newest point used directly as the first stage (no copy into K[0]), and -- as in
        // advance_ab -- a trajectory that stops inside the block skips the rest of it instead of branching out.
        // 32-bit step counters.  Saves ~25 register moves of a 100-instruction RungeKutta4 step on n = 2.
        int nst = 0, rot = 0;
        const int max_att = a.max_attempts > 2147483647LL ? 2147483647 : (int)a.max_attempts;
        const int n_steps = a.n_steps > 2147483647LL ? 2147483647 : (int)a.n_steps;
        double *dst = MODE == NBK_MODE_RUN ? a.y_out + (idx * a.o_sb + kout * a.o_sk) : a.y_out + idx * a.o_sb;
        const double *tg = a.tgrid + kout + 1;  // RUN: the grid time after te_next, fetched one point ahead
        double te_after = (MODE == NBK_MODE_RUN && running && kout + 1 < a.m) ? __ldg(tg) : d_inf();
        while (running) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int nw = (1 + r) % 2, od = r % 2;  // physical slots of the newest / the older point
                const double t = ts[nw], t_new = t + h;
                if (running && t_new > a.bound) {  // guard of this step (core.py:176-184)
                    status = NBK_TRAJ_BOUND;
                    running = false;
                }
                if (running) {
#pragma unroll
                    for (int s = 1; s < S; ++s) {
                        double ystage[N];
#pragma unroll
                        for (int c = 0; c < N; ++c) {
                            double acc = 0.0;
                            bool first = true;
#pragma unroll
                            for (int j = 0; j < s; ++j) {
                                if (Tab::a(s, j) != 0.0) {
                                    const double kj = j == 0 ? fs[nw][c] : K[j > 0 ? j : 1][c];
                                    acc = first ? Tab::a(s, j) * kj : fma(Tab::a(s, j), kj, acc);
                                    first = false;
                                }
                            }
                            ystage[c] = fma(h, acc, ys[nw][c]);
                        }
                        Rhs::eval(fma(h, Tab::c(s), t), ystage, p, K[s]);
                    }
                    double ynew[N], fnew[N];
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        // y + (h*B) . K with the vector h*B formed first, as the reference does
                        double acc = 0.0;
                        bool first = true;
#pragma unroll
                        for (int j = 0; j < S; ++j) {
                            if (Tab::b(j) != 0.0) {
                                const double kj = j == 0 ? fs[nw][c] : K[j > 0 ? j : 1][c];
                                acc = first ? (h * Tab::b(j)) * kj : fma(h * Tab::b(j), kj, acc);
                                first = false;
                            }
                        }
                        ynew[c] = ys[nw][c] + acc;
                    }
                    Rhs::eval(t_new, ynew, p, fnew);
                    if (MODE == NBK_MODE_RUN) {
                        // the window of the reference after its push is [t, t_new]: emit before the older slot is reused
                        while (t_new >= te_next) {
                            double out[N];
                            hermite_ends<N, true>(te_next, t, ys[nw], fs[nw], t_new, ynew, fnew, out);
#pragma unroll
                            for (int c = 0; c < N; ++c) dst[c * a.o_sc] = out[c];
                            dst += a.o_sk;
                            ++kout;
                            ++tg;
                            te_next = te_after;
                            te_after = kout + 1 < a.m ? __ldg(tg) : d_inf();
                        }
                    }
                    ts[od] = t_new;  // push: the older slot becomes the newest one
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        ys[od][c] = ynew[c];
                        fs[od][c] = fnew[c];
                    }
                    ++nst;
                    rot = (r + 1) % 2;
                    if (MODE == NBK_MODE_RUN) {
                        running = kout < a.m;
                    } else {
                        if (a.emit) {
                            a.t_out[idx * a.to_sb + (nst - 1)] = t_new;
#pragma unroll
                            for (int c = 0; c < N; ++c) dst[c * a.o_sc] = ynew[c];
                            dst += a.o_sk;
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
            if (rot == 0) store_window<2, N, 0>(a, idx, ts, ys, fs);
            else store_window<2, N, 1>(a, idx, ts, ys, fs);
#pragma unroll
            for (int c = 0; c < N; ++c) a.K[(long long)c * B + idx] = rot == 0 ? fs[0][c] : fs[1][c];  // f at the start of the last step
#pragma unroll
            for (int s = 1; s < S; ++s)
#pragma unroll
                for (int c = 0; c < N; ++c) a.K[((long long)s * N + c) * B + idx] = K[s][c];
            a.n_acc[idx] += nst;
        }
        store_status(a, idx, status);
        a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : nst;
        return;
    }
    while (running) {
        if (ts[1] + h > a.bound) {
            status = NBK_TRAJ_BOUND;
            break;
        }
        const double t = ts[1];
#pragma unroll
        for (int c = 0; c < N; ++c) K[0][c] = fs[1][c];
#pragma unroll
        for (int s = 1; s < S; ++s) {
            double ystage[N];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double acc = 0.0;
                bool first = true;
#pragma unroll
                for (int j = 0; j < s; ++j) {
                    if (Tab::a(s, j) != 0.0) {
                        acc = first ? Tab::a(s, j) * K[j][c] : fma(Tab::a(s, j), K[j][c], acc);
                        first = false;
                    }
                }
                ystage[c] = fma(h, acc, ys[1][c]);
            }
            Rhs::eval(fma(h, Tab::c(s), t), ystage, p, K[s]);
        }
        const double t_new = t + h;
        double ynew[N], fnew[N];
#pragma unroll
        for (int c = 0; c < N; ++c) {
            // y + (h*B) . K with the vector h*B formed first, as the reference does
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int j = 0; j < S; ++j) {
                if (Tab::b(j) != 0.0) {
                    acc = first ? (h * Tab::b(j)) * K[j][c] : fma(h * Tab::b(j), K[j][c], acc);
                    first = false;
                }
            }
            ynew[c] = ys[1][c] + acc;
        }
        Rhs::eval(t_new, ynew, p, fnew);
        ts[0] = ts[1];
        ts[1] = t_new;
#pragma unroll
        for (int c = 0; c < N; ++c) {
            ys[0][c] = ys[1][c];
            fs[0][c] = fs[1][c];
            ys[1][c] = ynew[c];
            fs[1][c] = fnew[c];
        }
        ++nsteps;
#ifdef NBK_NEVENTS
        if constexpr (EV && MODE == NBK_MODE_RUN) {
            window_dense<2, N> dense(ts, ys, fs);
            double t_term = 0.0;
            if (events_after_step<N, NPX>(a, idx, ts[0], t_new, ynew, p, ev_last, dense, t_term, status)) {
                if (status != NBK_TRAJ_EVBRACKET) {
                    double out[N];
                    dense(t_term, out);
#pragma unroll
                    for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + kout * a.o_sk + c * a.o_sc] = out[c];
                    a.ev_tterm[idx] = t_term;
                    ++kout;
                    status = NBK_TRAJ_EVENT;
                }
                break;
            }
        }
#endif
        if (MODE == NBK_MODE_RUN) {
            if (t_new >= te_next) {
                double *dst = a.y_out + (idx * a.o_sb + kout * a.o_sk);
                // te_next > ts[0] here (earlier grid points were written by earlier steps); the grid time after the
                // next one is fetched one point ahead so that the loop condition never waits for a load
                const double *tg = a.tgrid + kout + 1;
                double te_after = kout + 1 < a.m ? __ldg(tg) : d_inf();
                do {
                    double out[N];
                    hermite_ends<N, true>(te_next, ts[0], ys[0], fs[0], ts[1], ys[1], fs[1], out);
#pragma unroll
                    for (int c = 0; c < N; ++c) dst[c * a.o_sc] = out[c];
                    dst += a.o_sk;
                    ++kout;
                    ++tg;
                    te_next = te_after;
                    te_after = kout + 1 < a.m ? __ldg(tg) : d_inf();
                } while (t_new >= te_next);
            }
            running = kout < a.m;
        } else {
            if (a.emit) {
                a.t_out[idx * a.to_sb + (nsteps - 1)] = t_new;
#pragma unroll
                for (int c = 0; c < N; ++c) a.y_out[idx * a.o_sb + (nsteps - 1) * a.o_sk + c * a.o_sc] = ynew[c];
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
        for (int l = 0; l < 2; ++l) {
            a.ts[(long long)l * B + idx] = ts[l];
#pragma unroll
            for (int c = 0; c < N; ++c) {
                a.ys[((long long)l * N + c) * B + idx] = ys[l][c];
                a.fs[((long long)l * N + c) * B + idx] = fs[l][c];
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int c = 0; c < N; ++c) a.K[((long long)s * N + c) * B + idx] = K[s][c];
        a.n_acc[idx] += nsteps;
    }
    store_status(a, idx, status);
    a.n_out[idx] = MODE == NBK_MODE_RUN ? kout : (int)nsteps;
}

}  // namespace nbk
#endif
