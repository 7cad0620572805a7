"""GPU parity tests of the CTA-per-trajectory regime (K3, csrc/nbk_cta.cuh): large systems
(BASELINE.json configs[3] linear n=128 with a shared matrix, configs[4] 1-D reaction-diffusion
method of lines n=2048) against the C oracle and the reference's golden vectors."""

import math

import numpy as np
import pytest

from conftest import load_golden, parity_bar

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from oracle import nbk_oracle_c as oc  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402


def within(y, ref, rtol, atol, slack=1.0):
    return np.abs(np.asarray(y) - np.asarray(ref)) <= slack * (rtol * np.abs(ref) + atol)


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b))))


@pytest.mark.parametrize("cls,n,B", [("RungeKutta45", 2048, 48), ("RungeKutta23", 2048, 8), ("RungeKutta45", 100, 40),
                                     ("RungeKutta45", 700, 12), ("RungeKutta45", 3000, 6), ("RungeKutta45", 48, 16),
                                     ("RungeKutta45", 1024, 8), ("RungeKutta23", 1000, 8)])
def test_rd1d_vs_oracle(cls, n, B):
    """C5 generator (kappa, r, phase sweep).  n = 2048 / 1024 run the "full" programs (every thread owns four valid
    components, no bounds checks compiled in); n = 48 / 100 / 700 / 1000 / 3000 the any-n programs with 1, 2, 4, 4 and
    12 components per thread and a ragged last thread (some shapes compiled by NVRTC)."""
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    ts = [0.25, 0.5]
    ref = oc.run_ensemble(cls, "rd1d", y0, p, ts, rtol=1e-6, atol=1e-9, nthreads=8)
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    assert s._regime == 1
    assert rel(s.h, ref["h0"]) < 1e-12
    t, y = s.run(ts)
    assert y.shape == (B, 2, n)
    assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
    # The explicit pairs run this stiff grid at the edge of their stability region (h d ~ 0.8, first step far outside
    # it), where rounding noise in the highest grid mode is amplified ~40x per stage and reaches the error estimate: two
    # faithful implementations differ by O(tolerance).  The checker here is the BIT-EXACT NumPy restatement of the
    # reference; bars = measured + margin (profiles/r02_parity.json, B200):
    #   RungeKutta45 (the pair of BASELINE configs[4]): GPU vs restatement <= 0.63x on every shape but one (0.08x at
    #     n = 2048, 0.04x on configs[4] itself to t = 1) -> the plain north-star bar, 1x; n = 700 (eight trajectories):
    #     1.14x, with the C port 1.14x away from the GPU as well -> 1.5x there;
    #   RungeKutta23 (not a BASELINE pair on this grid): GPU vs restatement 1.54x / 1.91x at n = 2048 / 1000, while the
    #     C port differs from the SAME restatement by 0.76x / 1.91x -- the noise floor of the problem itself -> 2.5x.
    # Against the C port (all B trajectories): measured <= 1.14x (RungeKutta45, n = 700), 1.42x (RungeKutta23) -> 1.5x / 2.5x.
    nnp = min(B, 12 if n >= 2048 else 8)
    ynp = []
    for i in range(nnp):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
        ynp.append(onp.run(o, ts)[1])
        assert abs(int(s.n_accepted[i]) - o.n_accepted) <= 1
    rk45 = cls == "RungeKutta45"
    parity_bar(f"c5/{cls}/n{n}/gpu_vs_numpy_restatement", y[:nnp], np.array(ynp), 1e-6, 1e-9,
               (1.5 if n == 700 else 1.0) if rk45 else 2.5)
    parity_bar(f"c5/{cls}/n{n}/c_port_vs_numpy_restatement(noise floor, not a bar)", ref["ys"][:nnp], np.array(ynp), 1e-6, 1e-9, 1e9)
    parity_bar(f"c5/{cls}/n{n}/gpu_vs_c_port", y, ref["ys"], 1e-6, 1e-9, 1.5 if rk45 else 2.5)
    # (two CPU restatements with different FMA contraction differ by up to 0.99x on the RungeKutta23 / n = 1000 case)
    assert within(y, ref["ys"], 1e-6, 1e-9).mean() > 0.99
    # the end state sits at each implementation's own (noise-dependent) last step: loose check only
    # (within one step: whether the last step lands just before or just after ts[-1] is noise, too)
    assert rel(s.t, ref["t_end"]) < 2e-2 and rel(s.y, ref["y_end"]) < 2e-2
    assert (s.status == 0).all() and s.K.shape == (B, s.STAGE_ROWS, n) and s.cache.ys.shape == (B, 2, n)


def test_linear_shared_matrix_vs_oracle():
    """C4: y' = A y, n = 128, one dense A shared by the whole ensemble."""
    B, n = 300, 128
    y0, A = onp.linear_ensemble(B, n=n, seed=2)
    ref = oc.run_ensemble("RungeKutta45", "linear", y0, A, [1.0, 10.0], rtol=1e-6, atol=1e-9, params_shared=True, nthreads=8)
    s = nbk.RungeKutta45("linear", 0.0, y0, A, rtol=1e-6, atol=1e-9)
    assert s._regime == 1 and not s.is_jit
    t, y = s.run([1.0, 10.0])
    assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
    assert within(y, ref["ys"], 1e-6, 1e-9).all() and rel(y, ref["ys"]) < 1e-9


def test_golden_large_n_cases_through_cta_kernels():
    for c in load_golden("adaptive_single.json"):
        if c["rhs"] not in ("rd1d", "linear"):
            continue
        kw = dict(c["kw"])
        s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), **kw)
        assert s._regime == 1
        assert rel(float(s.h), c["h0"]) < 1e-12 and rel(s.f, c["f0"]) < 1e-13
        t, y = s.run(c["ts"])
        assert s.n_accepted == c["n_accepted"]
        assert rel(y, c["ys"]) < 1e-9 and rel(s.K, c["K_end"]) < 1e-6
        assert isinstance(s.t, float) and s.y.shape == (len(c["y0"]),)


USER_RD1D = """
NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n) {
    const double yl = y[i == 0 ? n - 1 : i - 1], yr = y[i + 1 == n ? 0 : i + 1];
    return p[0] * (yr - 2 * y[i] + yl) + p[1] * y[i] * (1 - y[i]);
}
"""


def test_user_cuda_rhs_and_api_in_cta_regime():
    B, n = 10, 600
    y0, p = onp.rd1d_ensemble(B, n=n, seed=5)
    a = nbk.RungeKutta45("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    b = nbk.RungeKutta45(nbk.CudaRHS(USER_RD1D), 0.0, y0, p, rtol=1e-6, atol=1e-9)
    assert b.is_jit and b._regime == 1
    ta, ya = a.run([0.2, 0.4])
    tb, yb = b.run([0.2, 0.4])
    # same counts; values agree to tolerance (a different FMA contraction of the stencil is enough to
    # change the amplified rounding noise, see test_rd1d_vs_oracle)
    assert np.abs(a.n_accepted - b.n_accepted).max() <= 1
    parity_bar("c5/user_cuda_rhs_vs_builtin/n600", yb, ya, 1e-6, 1e-9, 2.0)  # measured 1.54x: two FMA contractions of one stencil
    # step / skip / resumed run / interpolate / t_bound in the CTA regime
    c = nbk.RungeKutta45("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t1, y1 = c.step(n=5)
    assert t1.shape == (B, 5) and y1.shape == (B, 5, n)
    c.skip(n=3)
    o = onp.AdaptiveRK("RungeKutta45", onp.rhs_rd1d, 0.0, y0[0], p[0], rtol=1e-6, atol=1e-9)
    tr, yr = onp.steps(o, n=8)
    # the second step (h grows 10x, far outside the stability region) amplifies rounding noise into the error
    # estimate, so the step SIZES of two correct implementations differ at the 1e-5 level from the third step on
    # (emulating either FMA contraction of the stencil in NumPy gives rel t 2-3e-5, rel y 0.8-1.1e-5)
    assert rel(t1[0], tr[:5]) < 1e-3 and rel(y1[0], yr[:5]) < 5e-5 and rel(c.t[0], tr[-1]) < 1e-3
    tc, yc = c.run([3.0])  # resumed: continues from the 8 steps above
    tf, yf = nbk.RungeKutta45("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9).run([3.0])
    assert np.array_equal(yc, yf)  # same kernels, same arithmetic: the split run is bit-identical
    lo, hi = c.cache.ts
    mid = c.interpolate(float(max(lo.max(), hi.min() - 1e-4)))
    assert mid.shape == (B, n) and np.isfinite(mid).all()
    d = nbk.RungeKutta45("rd1d", 0.0, y0[0], p[0], rtol=1e-6, atol=1e-9, t_bound=0.1)
    with pytest.raises(ValueError):
        d.run([0.2])
    td, yd = d.step(upto_t=0.1)
    assert len(td) > 0 and td[-1] <= 0.1 and d.status == 1


def test_linear_dmma_kernel_k4():
    """K4 (csrc/nbk_mma.cuh): right-hand side of a 16-trajectory tile as an FP64 DMMA GEMM.  Checked
    against the C oracle and against the generic CTA kernel (NBK_NO_MMA=1) on the same ensemble,
    including a partial last tile, n < 128 (zero-padded rows), dense output and step()."""
    import os

    for n, B in ((128, 300), (96, 37), (8, 5)):
        y0, A = onp.linear_ensemble(B, n=n, seed=2)
        ts = [0.5, 1.0, 4.0, 10.0]
        ref = oc.run_ensemble("RungeKutta45", "linear", y0, A, ts, rtol=1e-6, atol=1e-9, params_shared=True, nthreads=8)
        s = nbk.RungeKutta45("linear", 0.0, y0, A, rtol=1e-6, atol=1e-9)
        assert s._regime == 1 and s.last_launch()["block"] == 0  # nothing launched yet by run/step
        t, y = s.run(ts)
        assert s.last_launch()["block"] == 256  # the DMMA program
        assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
        assert within(y, ref["ys"], 1e-6, 1e-9).all() and rel(y, ref["ys"]) < 1e-9
        assert rel(s.t, ref["t_end"]) < 1e-7 and rel(s.y, ref["y_end"]) < 1e-7 and rel(s.h, ref["h_end"]) < 1e-6
        os.environ["NBK_NO_MMA"] = "1"
        try:
            g = nbk.RungeKutta45("linear", 0.0, y0, A, rtol=1e-6, atol=1e-9)
        finally:
            del os.environ["NBK_NO_MMA"]
        tg, yg = g.run(ts)
        assert g.last_launch()["block"] != 256
        assert np.array_equal(g.n_accepted, s.n_accepted) and rel(y, yg) < 1e-11 and rel(s.K, g.K) < 1e-8
        # resumed run and step() through K4
        r = nbk.RungeKutta45("linear", 0.0, y0, A, rtol=1e-6, atol=1e-9)
        _, y1 = r.run(ts[:2])
        _, y2 = r.run(ts[2:])
        assert np.array_equal(np.concatenate([y1, y2], 1), y)
        mid = 0.5 * (r.cache.ts[0].max() + r.cache.ts[1].min())
        if r.cache.ts[0].max() <= mid <= r.cache.ts[1].min():
            assert rel(r.interpolate(float(mid)), g.interpolate(float(mid))) < 1e-10
        q = nbk.RungeKutta45("linear", 0.0, y0, A, rtol=1e-6, atol=1e-9)
        tq, yq = q.step(n=4)
        o = onp.AdaptiveRK("RungeKutta45", onp.rhs_linear, 0.0, y0[0], A, rtol=1e-6, atol=1e-9)
        tr, yr = onp.steps(o, n=4)
        assert rel(tq[0], tr) < 1e-8 and rel(yq[0], yr) < 1e-8  # step times: the controller feeds ulp changes of E.K back into h
    s23 = nbk.RungeKutta23("linear", 0.0, y0, A, rtol=1e-5, atol=1e-8)
    ref = oc.run_ensemble("RungeKutta23", "linear", y0, A, [2.0], rtol=1e-5, atol=1e-8, params_shared=True)
    assert within(s23.run([2.0])[1], ref["ys"], 1e-5, 1e-8).all() and np.abs(s23.n_accepted - ref["n_accepted"]).max() <= 1


def test_python_vector_rhs_traced_into_cta_regime():
    """A NumPy-style right-hand side of a large system (what a reference user writes) runs through K3 unmodified."""
    n, B = 2048, 6
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)

    def rd(t, y, p):
        d, r = p
        return d * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + r * y * (1 - y)

    s = nbk.RungeKutta45(rd, 0.0, y0, p, rtol=1e-6, atol=1e-9)
    assert s.is_jit and s._regime == 1
    t, y = s.run([0.05, 0.1])
    b = nbk.RungeKutta45("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    tb, yb = b.run([0.05, 0.1])
    assert np.abs(s.n_accepted - b.n_accepted).max() <= 1
    tol = 1e-6 * np.abs(yb) + 1e-9
    parity_bar("c5/traced_python_rhs_vs_builtin", y, yb, 1e-6, 1e-9, 1.5)  # stability-limited grid (test_rd1d_vs_oracle)

    # non-periodic stencil assembled by slice assignment, captured grid, time-dependent source: against the NumPy oracle
    m = 100
    x = np.linspace(0, 1, m)

    def heat(t, y, p):
        dy = np.zeros_like(y)
        dy[1:-1] = p[0] * (y[2:] - 2 * y[1:-1] + y[:-2]) + np.sin(x[1:-1]) * np.exp(-t)
        dy[0] = -y[0]
        dy[-1] = 0.0
        return dy

    rng = np.random.default_rng(8)
    y0h = rng.uniform(0, 1, (3, m))
    ph = np.array([[30.0], [60.0], [90.0]])
    s = nbk.RungeKutta23(heat, 0.0, y0h, ph, rtol=1e-5, atol=1e-8)
    t, y = s.run([0.02, 0.05])
    for i in range(3):
        o = onp.AdaptiveRK("RungeKutta23", heat, 0.0, y0h[i], ph[i], rtol=1e-5, atol=1e-8)
        _, yo = onp.run(o, [0.02, 0.05])
        assert abs(s.n_accepted[i] - o.n_accepted) <= 1
        parity_bar(f"cta/heat_rk23_vs_numpy_restatement/{i}", y[i], yo, 1e-5, 1e-8, 2.5)


FIXED_CTA = [("ForwardEuler", {}), ("AdamsBashforth2", {}), ("AdamsBashforth4", {}), ("AdamsBashforth5", {"first_stepper_cls": "RungeKutta23"}),
             ("AdamsBashforth3", {"first_stepper_cls": None}), ("RungeKutta4", {}), ("Heun3", {}), ("Runge2", {}), ("Runge3", {}),
             ("RungeKutta3_8", {})]


@pytest.mark.parametrize("cls,kw", FIXED_CTA, ids=lambda x: x if isinstance(x, str) else "-".join(f"{v}" for v in x.values()) or "default")
def test_fixed_step_solvers_in_cta_regime(cls, kw):
    """Adams-Bashforth / ForwardEuler / fixed-step Runge-Kutta for large systems (one CTA per trajectory) against the
    C oracle: the fixed-step bar of 1e-12 relative.  Where an adaptive start-up with DEFAULT tolerances (rtol 1e-3)
    feeds the history the bar is 1e-6, far inside that start-up's own tolerance (rtol |y| + atol ~ 1e-3): on this
    stability-limited grid rounding differences reach the start-up's error estimate and one accept/reject decision
    may flip; measured 1e-13 on two of three trajectories and up to 8e-8 on the third."""
    fs = kw.get("first_stepper_cls", "auto")
    adaptive_startup = cls.startswith("AdamsBashforth") and int(cls[-1]) > 1 and fs is not None
    tol = 1e-6 if adaptive_startup else 1e-12
    for n, B, h, T in ((2048, 3, 2e-4, 0.01), (100, 5, 1e-3, 0.05)):
        y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
        if n != 2048:
            p = p * (n / 2048.0) ** 2  # keep h d < stability limit of the explicit methods on the coarser grid
        ts = np.linspace(0, T, 6)
        ref = oc.run_ensemble(cls, "rd1d", y0, p, ts, h=h, first_stepper=fs, nthreads=4)
        s = getattr(nbk, cls)("rd1d", 0.0, y0, p, h=h, **kw)
        assert s._regime == 1 and s.is_jit
        t, y = s.run(ts)
        err = np.max(np.abs(y - ref["ys"]), axis=(1, 2)) / np.maximum(1.0, np.max(np.abs(ref["ys"]), axis=(1, 2)))
        assert err.max() <= tol, (cls, n, err.max())
        assert np.array_equal(s.t, ref["t_end"]) and (s.status == 0).all()
        assert rel(s.y, ref["y_end"]) <= tol and rel(s.f, ref["f_end"]) <= 1e3 * tol


def test_fixed_step_cta_api_linear_and_traced_python():
    rng = np.random.default_rng(5)
    n, B = 70, 6
    A = rng.standard_normal((n, n)) / math.sqrt(n) - np.eye(n)
    y0 = rng.standard_normal((B, n))
    ref = oc.run_ensemble("RungeKutta4", "linear", y0, A, [0.5, 1.0], h=0.01, params_shared=True)
    s = nbk.RungeKutta4("linear", 0.0, y0, A, h=0.01)
    t, y = s.run([0.5, 1.0])
    assert rel(y, ref["ys"]) <= 1e-12
    # step / skip / upto_t with the reference's value flow
    s = nbk.AdamsBashforth2("rd1d", 0.0, *onp.rd1d_ensemble(2, n=300, seed=1), h=1e-4, first_stepper_cls=None)
    tt, yy = s.step(n=5)
    assert tt.shape == (2, 5) and yy.shape == (2, 5, 300) and np.allclose(tt[0], 1e-4 * np.arange(1, 6), rtol=1e-12)
    s.skip(upto_t=1.05e-3)
    assert np.allclose(s.t, 1e-3, rtol=1e-9)
    assert s.cache.ys.shape == (2, 2, 300)
    # a NumPy-style Python right-hand side through the vector tracer, fixed step
    def rd(t, y, p):
        return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)

    y0, p = onp.rd1d_ensemble(2, n=512, seed=2)
    p = p / 16.0
    a = nbk.RungeKutta4(rd, 0.0, y0, p, h=1e-3)
    b = nbk.RungeKutta4("rd1d", 0.0, y0, p, h=1e-3)
    assert rel(a.run([0.02])[1], b.run([0.02])[1]) <= 1e-11
    # the implicit families are per-thread only: a clear error
    with pytest.raises(ValueError, match="per-thread"):
        nbk.BDF2("rd1d", 0.0, y0, p, h=1e-3)


@pytest.mark.parametrize("n,B", [(1500, 12), (2048, 10), (3000, 6)])
def test_dop853_beyond_1024_components(n, B):
    """DOP853 with one CTA per trajectory beyond n = 1024: four components per thread up to 2048 (the stage vectors then
    spill past the register file), the 256-thread shape above -- against the C port: within the tolerance, accepted steps
    within +-1; run_events on the same systems."""
    y0, p = onp.rd1d_ensemble(B, n=n, seed=2)
    p[:, 0] *= 0.05 * (100.0 / n) ** 2  # a non-stiff grid: rounding differences stay at rounding level
    ts = [1.0, 3.0, 6.0]
    kw = dict(rtol=1e-8, atol=1e-11)   # 7 to 14 accepted steps per trajectory
    ref = oc.run_ensemble("DOP853", "rd1d", y0, p, ts, nthreads=8, **kw)
    s = nbk.DOP853("rd1d", 0.0, y0, p, **kw)
    assert s._regime == 1 and s._group == 0
    t, y = s.run(ts)
    assert np.all(np.abs(np.asarray(s.n_accepted) - ref["n_accepted"]) <= 1)
    parity_bar(f"dop853_large/n{n}/gpu_vs_c_port", y, ref["ys"], 1e-8, 1e-11, 1.0)

    def level(t, yv):
        return yv[n // 3] - 0.7

    s2 = nbk.DOP853("rd1d", 0.0, y0, p, **kw)
    t2, y2, te, ye = s2.run_events(ts, [level])
    assert np.array_equal(y2, y)  # a non-terminal event does not change the trajectory
    cnt = s2.event_counts
    for b in np.nonzero(cnt[:, 0] > 0)[0][:4]:
        assert abs(ye[0][b, 0, n // 3] - 0.7) < 1e-7


def test_auto_regime_for_mid_size_traced_systems():
    """A traced Python right-hand side with 16 < n <= 64 runs in the CTA regime by default (the per-thread kernels
    leave the register file around n = 12); regime="thread" forces the per-thread form; both agree with the oracle."""
    rng = np.random.default_rng(21)
    n, B = 32, 200
    A = rng.standard_normal((n, n)) / math.sqrt(n) - np.eye(n)
    y0 = rng.standard_normal((B, n))

    def lin(t, y):
        return A @ y

    auto = nbk.RungeKutta45(lin, 0.0, y0, rtol=1e-6, atol=1e-9)
    thread = nbk.RungeKutta45(lin, 0.0, y0, rtol=1e-6, atol=1e-9, regime="thread")
    assert auto._regime == 1 and thread._regime == 0
    ts = [1.0, 5.0]
    ya, yt = auto.run(ts)[1], thread.run(ts)[1]
    ref = oc.run_ensemble("RungeKutta45", "linear", y0, A, ts, rtol=1e-6, atol=1e-9, params_shared=True, nthreads=4)
    for y, s in ((ya, auto), (yt, thread)):
        assert within(y, ref["ys"], 1e-6, 1e-9).all()
        assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
    # a small system written component by component, forced into the CTA regime, and the parameter block as a matrix
    def lorenz(t, y, p):
        x, yy, z = y
        return np.array([p[0] * (yy - x), x * (p[1] - z) - yy, x * yy - p[2] * z])

    y0l, pl = onp.lorenz_ensemble(64, seed=5)
    a = nbk.RungeKutta45(lorenz, 0.0, y0l, pl, rtol=1e-6, atol=1e-9, regime="cta")
    b = nbk.RungeKutta45("lorenz", 0.0, y0l, pl, rtol=1e-6, atol=1e-9)
    assert a._regime == 1 and b._regime == 0
    assert within(a.run([2.0])[1], b.run([2.0])[1], 1e-6, 1e-9, slack=50).all()  # chaos: 50x like smoke()
    assert np.abs(a.n_accepted - b.n_accepted).max() <= 1

    def linp(t, y, p):
        return p @ y

    c = nbk.RungeKutta23(linp, 0.0, y0, A, rtol=1e-5, atol=1e-8)  # shared (n, n) parameter block
    assert c._regime == 1
    refc = oc.run_ensemble("RungeKutta23", "linear", y0, A, [2.0], rtol=1e-5, atol=1e-8, params_shared=True, nthreads=4)
    assert within(c.run([2.0])[1], refc["ys"], 1e-5, 1e-8).all()
    # DOP853 has a CTA regime too; regime="thread" keeps the per-thread kernels (e.g. for run_events)
    d = nbk.DOP853(lin, 0.0, y0[:4], rtol=1e-6, atol=1e-9)
    e = nbk.DOP853(lin, 0.0, y0[:4], rtol=1e-6, atol=1e-9, regime="thread")
    assert d._regime == 1 and e._regime == 0
    assert within(d.run([1.0, 2.0])[1], e.run([1.0, 2.0])[1], 1e-6, 1e-9).all() and np.abs(d.n_accepted - e.n_accepted).max() <= 1


def test_dop853_in_cta_regime():
    """DOP853 (explicit.py:282-351, 406-481) with one CTA per trajectory: 13 stage vectors per thread, E5/E3 error norm as
    two block reductions, dense output with the three extra stages evaluated by the whole CTA.  Against the C oracle on
    the reaction-diffusion generator: n = 100 (one component per thread, ragged), 512 (the "full" program), 700 (two
    components per thread); resumed runs, step(), interpolate()."""
    for n, B in ((100, 6), (512, 4), (700, 4)):
        y0, p = onp.rd1d_ensemble(B, n=n, seed=11)
        # keep the grid non-stiff (d = kappa n^2 of order 0.1): the 8th-order pair crosses t = 0.3 in 2-9 steps, and on a
        # stiff grid two CPU restatements that differ only in FMA contraction are already 6x (n = 512) to 112x (n = 700)
        # the tolerance apart -- nothing to pin a kernel against
        p[:, 0] *= 0.1 * (100.0 / n) ** 2
        ts = [0.5, 2.0, 6.0]
        ref = oc.run_ensemble("DOP853", "rd1d", y0, p, ts, rtol=1e-6, atol=1e-9, nthreads=4)
        s = nbk.DOP853("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
        assert s._regime == 1 and rel(s.h, ref["h0"]) < 1e-12
        t, y = s.run(ts)
        assert y.shape == (B, 3, n) and (s.status == 0).all() and s.K.shape == (B, 13, n)
        assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
        assert within(y, ref["ys"], 1e-6, 1e-9).all()
        r = nbk.DOP853("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
        _, y1 = r.run(ts[:1])
        _, y2 = r.run(ts[1:])
        assert np.array_equal(np.concatenate([y1, y2], 1), y)  # split run: same kernels, same arithmetic
        lo, hi = r.cache.ts
        mid = float(max(lo.max(), hi.min() - 1e-5))
        if lo.max() <= mid <= hi.min():
            assert np.isfinite(r.interpolate(mid)).all()
    q = nbk.DOP853("rd1d", 0.0, y0[0], p[0], rtol=1e-6, atol=1e-9)
    tq, yq = q.step(n=4)
    o = onp.AdaptiveRK("DOP853", onp.rhs_rd1d, 0.0, y0[0], p[0], rtol=1e-6, atol=1e-9)
    tr, yr = onp.steps(o, n=4)
    assert rel(tq, tr) < 1e-3 and rel(yq, yr) < 5e-5  # step sizes carry amplified rounding noise (see above)
    # dense output inside a step against the NumPy restatement's interpolator on a smooth (non-stiff) problem
    y0s, ps = onp.rd1d_ensemble(1, n=64, seed=12)
    ps[:, 0] *= 1e-3
    u = nbk.DOP853("rd1d", 0.0, y0s[0], ps[0], rtol=1e-8, atol=1e-11, regime="cta")
    v = onp.AdaptiveRK("DOP853", onp.rhs_rd1d, 0.0, y0s[0], ps[0], rtol=1e-8, atol=1e-11)
    grid = np.linspace(0.05, 2.0, 40)
    _, yu = u.run(grid)
    _, yv = onp.run(v, grid)
    assert u._regime == 1 and rel(yu, yv) < 1e-10


def test_cta_edge_shapes():
    """Limits of the regime: the largest state (n = 8192: 32 components per thread), one above it (refused with the
    C-ABI's message), a single unbatched trajectory, a grid point at t0, and one trajectory more than resident CTAs."""
    # n = 8192, B = 3 against the C oracle (stability-limited grid: measured + margin, see test_rd1d_vs_oracle)
    n = 8192
    y0, p = onp.rd1d_ensemble(3, n=n, seed=7)
    p[:, 0] *= 1.0 / 64  # keep the step count small: kappa / dx^2 grows with n^2
    ref = oc.run_ensemble("RungeKutta45", "rd1d", y0, p, [0.05], rtol=1e-6, atol=1e-9, nthreads=3)
    s = nbk.RungeKutta45("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    _, y = s.run([0.05])
    assert s.last_launch()["block"] == 256 and np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
    parity_bar("c5/RungeKutta45/n8192/gpu_vs_c_port", y, ref["ys"], 1e-6, 1e-9, 1.5)
    with pytest.raises((ValueError, nbk.NbkError), match="8192"):
        nbk.RungeKutta45("rd1d", 0.0, np.zeros((2, 8193)), np.ones((2, 2)))
    # one unbatched trajectory: reference shapes (m, n), scalar t; the first grid point is t0 itself
    y1, p1 = onp.rd1d_ensemble(1, n=256, seed=8)
    u = nbk.RungeKutta45("rd1d", 0.0, y1[0], p1[0], rtol=1e-6, atol=1e-9)
    t, yy = u.run([0.0, 0.1])
    assert yy.shape == (2, 256) and np.array_equal(yy[0], y1[0]) and np.ndim(u.t) == 0
    o = oc.run_ensemble("RungeKutta45", "rd1d", y1, p1, [0.1], rtol=1e-6, atol=1e-9, nthreads=1)
    parity_bar("c5/RungeKutta45/n256_unbatched/gpu_vs_c_port", yy[1], o["ys"][0, 0], 1e-6, 1e-9, 1.5)
    # more trajectories than resident CTAs by one: the queue hands the last one to whichever CTA finishes first
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    B = sm * 8 + 1
    y2, p2 = onp.rd1d_ensemble(B, n=128, seed=9)
    a = nbk.RungeKutta45("rd1d", 0.0, y2, p2, rtol=1e-6, atol=1e-9)
    _, ya = a.run([0.2])
    b = nbk.RungeKutta45("rd1d", 0.0, y2[::-1].copy(), p2[::-1].copy(), rtol=1e-6, atol=1e-9)
    _, yb = b.run([0.2])
    assert np.array_equal(ya, yb[::-1])  # position in the ensemble does not change a trajectory's result


@pytest.mark.parametrize("cls,rhs,n,B", [("RungeKutta45", "rd1d", 200, 40), ("DOP853", "rd1d", 150, 24), ("RungeKutta23", "rd1d", 2048, 6),
                                         ("RungeKutta45", "linear", 128, 50)])
def test_run_events_with_one_cta_per_trajectory(cls, rhs, n, B):
    """run_events for large systems (round 2): traced event callables through the CTA kernels' run_events variant -- for
    the shared-matrix linear system the solver's run / step go through the DMMA kernel and run_events through the generic
    CTA program on the same state -- against the NumPy restatement of the reference's run_events on a prefix: same events
    in the same order, times to the bisection's resolution, a terminal event truncates the output."""
    if rhs == "linear":
        y0, p = onp.linear_ensemble(B, n=n, seed=2)
        f, shared, ts = onp.rhs_linear, True, [1.0, 4.0, 9.0]
    else:
        y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
        p[:, 0] *= 0.4 * (100.0 / n) ** 2
        f, shared, ts = onp.rhs_rd1d, False, [0.5, 1.5, 3.0]

    def crossing(t, y):
        return y[0] - 0.5 * y0[0, 0]

    def gap(t, y):
        return y[n // 2] - y[n // 2 + 2]

    gap.direction = -1

    def level(t, y):
        return y[1] + y[2] - 0.9 * (y0[0, 1] + y0[0, 2])

    level.terminal = True
    evs = [crossing, gap, level]
    s = getattr(nbk, cls)(rhs, 0.0, y0, p, rtol=1e-6, atol=1e-9)
    assert s._regime == 1 and s._group == 0
    t, y, te, ye = s.run_events(ts, evs)
    cnt, st = s.event_counts, s.status
    assert cnt.shape == (B, 3) and set(np.unique(st)) <= {0, -1}
    total = 0
    for b in range(min(B, 5)):
        o = onp.AdaptiveRK(cls, f, 0.0, y0[b], p if shared else p[b], rtol=1e-6, atol=1e-9)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        k = len(t_ref)
        assert int(s._nout[b].item()) == k and (st[b] == -1) == (len(te_ref[2]) > 0)
        assert np.abs(y[b, :k - 1] - y_ref[:k - 1]).max() <= 2e-6 * max(1.0, np.abs(y_ref).max()) if k > 1 else True
        for j in range(3):
            assert cnt[b, j] == len(te_ref[j])
            total += len(te_ref[j])
            if len(te_ref[j]):
                assert np.abs(te[j][b, :len(te_ref[j])] - np.array(te_ref[j])).max() < 1e-5
                assert np.abs(ye[j][b, :len(te_ref[j])] - np.array(ye_ref[j])).max() < 1e-4 * max(1.0, np.abs(y_ref).max())
    assert total > 0


@pytest.mark.parametrize("cls,n,B,kw", [("RungeKutta4", 300, 40, dict(h=0.005)), ("AdamsBashforth2", 24, 200, dict(h=0.005, first_stepper_cls=None)),
                                        ("ForwardEuler", 2048, 12, dict(h=0.002)), ("Heun3", 13, 120, dict(h=0.005))])
def test_run_events_of_the_fixed_step_classes_componentwise(cls, n, B, kw):
    """run_events of the fixed-step classes for component-wise right-hand sides -- as sub-warp teams (n = 13, 24) and with
    one CTA per trajectory (n = 300, 2048) -- against the NumPy restatement of the reference's run_events on a prefix:
    same events in the same order, event times to the bisection's resolution, grid rows to 1e-12, a terminal event
    truncates the output."""
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    p[:, 0] *= 0.05 * (100.0 / n) ** 2
    ts = [0.25, 0.5, 1.0]

    def crossing(t, y):
        return y[0] - 0.6

    def gap(t, y):
        return y[n // 2] - y[n // 2 + 2]

    gap.direction = -1

    def level(t, y):
        return y[1] + y[2] - 1.5

    level.terminal = True
    level.direction = 1
    evs = [crossing, gap, level]
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
    assert s._regime == 1
    t, y, te, ye = s.run_events(ts, evs)
    cnt, st = s.event_counts, s.status
    assert cnt.shape == (B, 3) and set(np.unique(st)) <= {0, -1}
    total = 0
    for b in range(min(B, 6)):
        if cls in ("RungeKutta4", "Heun3"):
            o = onp.FixedRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], h=kw["h"])
        else:
            o = onp.AdamsBashforth(1 if cls == "ForwardEuler" else int(cls[-1]), onp.rhs_rd1d, 0.0, y0[b], p[b], h=kw["h"],
                                   first_stepper=None)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        k = len(t_ref)
        assert int(s._nout[b].item()) == k and (st[b] == -1) == (len(te_ref[2]) > 0)
        keep = k - 1 if st[b] == -1 else k
        if keep:
            assert rel(y[b, :keep], y_ref[:keep]) <= 1e-12
        assert np.isnan(y[b, k:]).all()
        for j in range(3):
            assert cnt[b, j] == len(te_ref[j]), (cls, b, j)
            total += len(te_ref[j])
            if len(te_ref[j]):
                assert np.abs(te[j][b, :len(te_ref[j])] - np.array(te_ref[j])).max() < 1e-7
                assert np.abs(ye[j][b, :len(te_ref[j])] - np.array(ye_ref[j])).max() < 1e-7
    assert total > 0
