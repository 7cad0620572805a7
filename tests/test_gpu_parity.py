"""GPU parity tests: the CUDA path (Python shell -> ctypes C-ABI -> sm_100a kernels) against
the reference's golden vectors and the C oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * fixed-step methods: 1e-12 relative error (max-norm over the trajectory's outputs);
  * adaptive methods: final/dense states within rtol*|y| + atol and accepted-step counts
    within +-1.
"""

import math

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from oracle import nbk_oracle_c as oc  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402

FIXED_TOL = 1e-12


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b))))


def within(y, ref, rtol, atol, slack=1.0):
    return np.abs(np.asarray(y) - np.asarray(ref)) <= slack * (rtol * np.abs(ref) + atol)


def test_library_is_loaded_and_native():
    from nbkode_b200 import _lib

    assert _lib.lib().nbk_version() == 1
    with open("/proc/self/maps") as fh:
        assert "libnbkode_b200.so" in fh.read()
    tf, ms = nbk.fp64_peak(0, 4000)
    assert tf > 5.0, tf  # B200 FP64 vector peak is ~30-37 TFLOP/s


# ------------------------------------------------------------------ C1: README example
def _readme_func(t, y, k):
    return -k * y


def test_c1_readme():
    for c in load_golden("c1_readme.json"):
        s = getattr(nbk, c["cls"])(_readme_func, 0.0, 1.0, params=np.array([0.1]), **c["kw"])
        assert s.is_jit  # traced Python -> CUDA C -> NVRTC
        t, y = s.run([0, 5, 10])
        assert y.shape == (3, 1)
        if s.FIXED_STEP:
            assert rel(y, c["ys"]) <= FIXED_TOL and s.t == c["t_end"]
        else:
            assert rel(float(s.h), c["h_end"]) < 1e-7
            assert within(y, c["ys"], 1e-3, 1e-6).all() and rel(y, c["ys"]) < 1e-10
            assert abs(s.t - c["t_end"]) < 1e-7


# ------------------------------------------------------------------ single trajectories
def _rhs_for(c):
    """built-in where one exists, otherwise the oracle's Python function traced to CUDA C"""
    return c["rhs"] if c["rhs"] in ("lorenz", "vdp", "decay") else onp.RHS[c["rhs"]]


@pytest.mark.parametrize("c", load_golden("adaptive_single.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_adaptive_single(c):
    kw = dict(c["kw"])
    rtol, atol = kw.get("rtol", 1e-3), kw.get("atol", 1e-6)
    s = getattr(nbk, c["cls"])(_rhs_for(c), 0.0, np.array(c["y0"]), np.array(c["p"]), **kw)
    assert rel(float(s.h), c["h0"]) < 1e-13
    assert rel(s.f, c["f0"]) < 1e-14
    t, y = s.run(c["ts"])
    assert y.shape == (len(c["ts"]), len(c["y0"]))
    assert abs(s.n_accepted - c["n_accepted"]) <= 1
    assert within(y, c["ys"], rtol, atol).all()
    assert within(s.y, c["y_end"], rtol, atol, slack=2).all()
    # tighter diagnostic: same step sequence => dense output agrees to ~1e-9 even for Lorenz
    if s.n_accepted == c["n_accepted"]:
        assert rel(y, c["ys"]) < 1e-8
        assert rel(s.K, c["K_end"]) < 1e-6 and rel(s.cache.ts, [c["t_old"], c["t_end"]]) < 1e-8


@pytest.mark.parametrize("c", load_golden("fixed_single.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{c['kw']}")
def test_fixed_single(c):
    kw = {k: (None if v == "None" else v) for k, v in c["kw"].items()}
    s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"], **kw)
    tol = FIXED_TOL  # Lorenz rows included: measured 1.8e-14 at most (profiles/r02_parity.json)
    assert rel(s.cache.ts, c["init_ts"]) < 1e-15
    assert rel(s.cache.ys, c["init_ys"]) <= tol and rel(s.cache.fs, c["init_fs"]) <= tol
    t, y = s.run(c["ts"])
    assert rel(y, c["ys"]) <= tol
    assert s.t == c["t_end"]
    assert rel(s.y, c["y_end"]) <= tol and rel(s.f, c["f_end"]) <= tol


@pytest.mark.parametrize("c", load_golden("erk_fixed.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_fixed_rk_single(c):
    """Fixed-step explicit Runge-Kutta classes: 1e-12 bar."""
    s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"])
    tol = FIXED_TOL
    t, y = s.run(c["ts"])
    assert rel(y, c["ys"]) <= tol and s.t == c["t_end"]
    assert rel(s.y, c["y_end"]) <= tol and rel(s.f, c["f_end"]) <= tol
    if "step10_t" in c:
        s = getattr(nbk, c["cls"])("vdp", 0.0, [2.0, 0.0], [1.0], h=0.05)
        tt, yy = s.step(n=10)
        assert rel(tt, c["step10_t"]) < 1e-15 and rel(yy, c["step10_y"]) <= FIXED_TOL and rel(s.K, c["step10_K"]) <= FIXED_TOL


def test_fixed_rk_ensemble_vs_oracle():
    B = 2048
    y0, p = onp.vdp_ensemble(B, seed=1)
    ts = np.linspace(0, 10, 101)
    for cls in ("RungeKutta4", "Heun3", "Runge2", "Runge3", "RungeKutta3_8"):
        ref = oc.run_ensemble(cls, "vdp", y0, p, ts, h=0.01, nthreads=8)
        s = getattr(nbk, cls)("vdp", 0.0, y0, p, h=0.01)
        t, y = s.run(ts)
        per_traj = np.max(np.abs(y - ref["ys"]), axis=(1, 2)) / np.maximum(1.0, np.max(np.abs(ref["ys"]), axis=(1, 2)))
        assert per_traj.max() <= FIXED_TOL, (cls, per_traj.max())
        assert np.array_equal(s.t, ref["t_end"])


# ------------------------------------------------------------------ C2: Lorenz ensemble
def test_lorenz_golden_rows():
    g = load_golden("lorenz_ensemble_rk45.json")
    rows = g["rows"]
    y0 = np.array([r["y0"] for r in rows])
    p = np.array([r["p"] for r in rows])
    s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=g["rtol"], atol=g["atol"])
    assert not s.is_jit
    assert rel(s.h, [r["h0"] for r in rows]) < 1e-13
    t, y = s.run([g["T"]])
    assert y.shape == (len(rows), 1, 3)
    dn = np.abs(s.n_accepted - np.array([r["n_accepted"] for r in rows]))
    assert dn.max() <= 1
    ref = np.array([r["ys"] for r in rows])
    assert within(y, ref, g["rtol"], g["atol"], slack=1).all()
    assert (s.status == 0).all()


def test_lorenz_ensemble_vs_oracle_20k():
    B = 20000
    y0, p = onp.lorenz_ensemble(B, seed=123)
    ref = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [10.0], rtol=1e-6, atol=1e-9, nthreads=8)
    s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y = s.run([10.0])
    ok_y = within(y, ref["ys"], 1e-6, 1e-9).all(axis=(1, 2))
    dn = np.abs(s.n_accepted - ref["n_accepted"])
    # a flipped accept/reject decision (ulp-level change of the error norm next to 1.0) moves a
    # chaotic trajectory by ~tol*1e4: report the pass fraction, expect (almost) all
    assert ok_y.mean() >= 0.999, ok_y.mean()
    assert (dn <= 1).mean() >= 0.999, (dn <= 1).mean()
    same = dn == 0
    assert same.mean() > 0.99
    assert rel(y[same], ref["ys"][same]) < 1e-6
    assert rel(s.t[same], ref["t_end"][same]) < 1e-7
    assert int(s.n_rejected.sum()) > 0 and abs(int(s.n_rejected.sum()) - int(ref["n_rejected"].sum())) < 0.001 * B * 60


def test_lorenz_full_size_properties():
    """BASELINE.json configs[1] at full size (1M trajectories): size-independent properties."""
    g = load_golden("lorenz_ensemble_rk45.json")
    B = g["B_total"]
    y0, p = onp.lorenz_ensemble(B, seed=g["seed"])
    s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y = s.run([10.0])
    assert (s.status == 0).all()
    t_end = s.t
    assert (t_end >= 10.0).all() and (t_end < 10.2).all()
    nacc = s.n_accepted
    assert 4.0e8 < nacc.sum() < 4.6e8 and 50 < nacc.min() and nacc.max() < 1200
    # the rows the live reference integrated sit at their places in the full ensemble
    idx = np.array([r["row"] for r in g["rows"]])
    ref = np.array([r["ys"] for r in g["rows"]])
    assert within(y[idx], ref, 1e-6, 1e-9).all()
    assert np.abs(nacc[idx] - np.array([r["n_accepted"] for r in g["rows"]])).max() <= 1
    # permutation invariance: a trajectory's arithmetic does not depend on its lane / queue slot
    perm = np.random.default_rng(0).permutation(B)[:200_000]
    s2 = nbk.RungeKutta45("lorenz", 0.0, y0[perm], p[perm], rtol=1e-6, atol=1e-9)
    t2, y2 = s2.run([10.0])
    assert np.array_equal(y2, y[perm]) and np.array_equal(s2.n_accepted, nacc[perm])


# ------------------------------------------------------------------ C3: Van der Pol, dense output
def test_vdp_ensemble_dense_output():
    g = load_golden("vdp_ensemble.json")
    ts = np.linspace(0, 20, g["n_ts"])
    y0 = np.array([r["y0"] for r in g["rk23"]])
    p = np.array([r["p"] for r in g["rk23"]])
    s = nbk.RungeKutta23("vdp", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y = s.run(ts)
    ref = np.array([r["ys"] for r in g["rk23"]])
    assert np.abs(s.n_accepted - np.array([r["n_accepted"] for r in g["rk23"]])).max() <= 1
    assert within(y[:, g["keep"]], ref, 1e-6, 1e-9).all()
    s = nbk.AdamsBashforth4("vdp", 0.0, y0, p, h=0.01)
    t, y = s.run(ts)
    assert rel(y[:, g["keep"]], [r["ys"] for r in g["ab4"]]) <= FIXED_TOL
    assert np.array_equal(s.t, [r["t_end"] for r in g["ab4"]])


def test_vdp_ensemble_vs_oracle_4k():
    B = 4096
    y0, p = onp.vdp_ensemble(B, seed=1)
    ts = np.linspace(0, 20, 200)
    for cls, kw in (("RungeKutta23", dict(rtol=1e-6, atol=1e-9)), ("AdamsBashforth4", dict(h=0.01)),
                    ("AdamsBashforth2", dict(h=0.01)), ("ForwardEuler", dict(h=0.01))):
        ref = oc.run_ensemble(cls, "vdp", y0, p, ts, nthreads=8, **kw)
        s = getattr(nbk, cls)("vdp", 0.0, y0, p, **kw)
        t, y = s.run(ts)
        if s.FIXED_STEP:
            per_traj = np.max(np.abs(y - ref["ys"]), axis=(1, 2)) / np.maximum(1.0, np.max(np.abs(ref["ys"]), axis=(1, 2)))
            assert per_traj.max() <= FIXED_TOL, (cls, per_traj.max())
            assert np.array_equal(s.t, ref["t_end"])
        else:
            assert within(y, ref["ys"], 1e-6, 1e-9).all(axis=(1, 2)).mean() >= 0.999
            assert (np.abs(s.n_accepted - ref["n_accepted"]) <= 1).mean() >= 0.999


# ------------------------------------------------------------------ engine properties
def test_batch_equals_loop_of_singles_and_nvrtc_equals_builtin():
    y0, p = onp.lorenz_ensemble(40, seed=3)
    ts = [1.0, 2.5]
    s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y = s.run(ts)
    for i in (0, 7, 39):
        s1 = nbk.RungeKutta45("lorenz", 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
        t1, y1 = s1.run(ts)
        assert np.array_equal(y1, y[i]) and s1.n_accepted == s.n_accepted[i] and s1.t == s.t[i]
    sj = nbk.RungeKutta45(onp.rhs_lorenz, 0.0, y0, p, rtol=1e-6, atol=1e-9)  # traced -> NVRTC
    assert sj.is_jit
    tj, yj = sj.run(ts)
    assert np.array_equal(sj.n_accepted, s.n_accepted) and rel(yj, y) < 1e-11


def test_resumed_and_unsorted_runs():
    y0, p = onp.lorenz_ensemble(64, seed=9)
    ts = np.linspace(0.3, 3.0, 10)
    a = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    ta, ya = a.run(ts)
    b = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    _, y1 = b.run(ts[:4])
    _, y2 = b.run(ts[4:])
    assert np.array_equal(ya, np.concatenate([y1, y2], axis=1))
    c = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    tc, yc = c.run(ts[::-1])
    assert np.array_equal(tc, ts[::-1]) and np.array_equal(yc, ya[:, ::-1])


def test_torch_tensors_stay_on_device():
    y0, p = onp.vdp_ensemble(1000, seed=4)
    y0_t, p_t = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
    s = nbk.RungeKutta23("vdp", 0.0, y0_t, p_t, rtol=1e-6, atol=1e-9)
    t, y = s.run([1.0, 2.0])
    assert isinstance(y, torch.Tensor) and y.is_cuda and y.shape == (1000, 2, 2)
    s2 = nbk.RungeKutta23("vdp", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    assert np.array_equal(s2.run([1.0, 2.0])[1], y.cpu().numpy())


def test_nonfinite_guard_terminates():
    s = nbk.RungeKutta45("decay", 0.0, np.array([[1.0], [1.0]]), np.array([[-1e308], [0.1]]), first_step=1.0)
    with pytest.raises(RuntimeError):
        s.run([10.0])
    assert list(s.status) == [2, 0]


def test_shared_matrix_params_traced_linear():
    rng = np.random.default_rng(5)
    n, B = 6, 50
    A = rng.standard_normal((n, n)) / math.sqrt(n) - np.eye(n)
    y0 = rng.standard_normal((B, n))
    s = nbk.RungeKutta45(onp.rhs_linear, 0.0, y0, A, rtol=1e-6, atol=1e-9)  # params shared: A is (n, n) != (B, ...)
    t, y = s.run([1.0, 5.0])
    ref = oc.run_ensemble("RungeKutta45", "linear", y0, A, [1.0, 5.0], rtol=1e-6, atol=1e-9, params_shared=True)
    assert within(y, ref["ys"], 1e-6, 1e-9).all() and np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1


# ------------------------------------------------------------------ DOP853
_DOP = load_golden("dop853.json")


@pytest.mark.parametrize("c", _DOP["single"], ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_dop853_single(c):
    kw = dict(c["kw"])
    rtol, atol = kw.get("rtol", 1e-3), kw.get("atol", 1e-6)
    s = nbk.DOP853(_rhs_for(c), 0.0, np.array(c["y0"]), np.array(c["p"]), **kw)
    assert rel(float(s.h), c["h0"]) < 1e-13
    t, y = s.run(c["ts"])
    assert y.shape == (len(c["ts"]), len(c["y0"]))
    assert abs(s.n_accepted - c["n_accepted"]) <= 1
    assert within(y, c["ys"], rtol, atol).all()
    assert within(s.y, c["y_end"], rtol, atol, slack=2).all()
    if s.n_accepted == c["n_accepted"]:
        assert rel(y, c["ys"]) < 1e-7
        assert rel(s.K, c["K_end"]) < 1e-5 and rel(s.cache.ts, [c["t_old"], c["t_end"]]) < 1e-7


def test_dop853_golden_rows_and_oracle_ensemble():
    g = _DOP["ensemble"]
    rows = g["rows"]
    y0 = np.array([r["y0"] for r in rows])
    p = np.array([r["p"] for r in rows])
    s = nbk.DOP853("lorenz", 0.0, y0, p, rtol=g["rtol"], atol=g["atol"])
    assert not s.is_jit
    assert rel(s.h, [r["h0"] for r in rows]) < 1e-13
    t, y = s.run([g["T"]])
    assert np.abs(s.n_accepted - np.array([r["n_accepted"] for r in rows])).max() <= 1
    assert within(y, np.array([r["ys"] for r in rows]), g["rtol"], g["atol"]).all()
    # 20k trajectories against the C oracle, dense output on a grid
    B = 20000
    y0, p = onp.lorenz_ensemble(B, seed=321)
    ts = np.linspace(0.5, 10, 20)
    ref = oc.run_ensemble("DOP853", "lorenz", y0, p, ts, rtol=1e-6, atol=1e-9, nthreads=8)
    s = nbk.DOP853("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y = s.run(ts)
    dn = np.abs(s.n_accepted - ref["n_accepted"])
    # Lorenz to t = 10 amplifies rounding differences ~1e4x and the 8(5,3) tableau (coefficients up to ~43) starts from
    # larger ones than RK45: the bit-exact NumPy restatement of the reference and the C oracle themselves disagree beyond
    # rtol|y|+atol on 4 of 3000 of these trajectories (0.13 %, identical step counts).  Measured on B200
    # (profiles/r02_parity.json): 33 of 20 000 trajectories (0.165 %) leave the bar somewhere on the grid, none before
    # t = 4, step counts identical on all 20 000.  Bars = measured + margin.
    ok = within(y, ref["ys"], 1e-6, 1e-9).all(axis=(1, 2))
    assert ok.mean() >= 0.9975, ok.mean()
    assert within(y[:, :8], ref["ys"][:, :8], 1e-6, 1e-9).all(axis=(1, 2)).all()  # t <= 4: little amplification
    assert (dn <= 1).all() and (dn == 0).mean() >= 0.9995
    assert (s.status == 0).all()


def test_dop853_api_flow():
    c = _DOP["api"]
    mk = lambda: nbk.DOP853("decay", 0.0, np.array([1.0]), np.array([0.01]))  # noqa: E731
    s = mk()
    ts, ys = s.step(n=20)
    assert rel(ts, c["ts20"]) < 1e-9 and rel(ys, c["ys20"]) < 1e-10
    s = mk()
    t, y = s.run(np.linspace(0, 4, 9))
    assert rel(y, c["resume_a"]) < 1e-11
    t, y = s.run(c["resume_b_t"])  # first point lies inside the stored last step
    assert rel(y, c["resume_b"]) < 1e-11
    s = mk()
    t, y = s.run(np.array(c["run_rev_t"]))
    assert rel(y, c["run_rev_y"]) < 1e-11
    # a traced Python right-hand side goes through NVRTC with the same kernels
    sj = nbk.DOP853(lambda t, y, k: -k * y, 0.0, np.array([1.0]), np.array([0.01]))
    assert sj.is_jit
    tj, yj = sj.run(np.linspace(0, 4, 9))
    assert rel(yj, c["resume_a"]) < 1e-11


# ------------------------------------------------------------------ implicit multistep
@pytest.mark.parametrize("c", load_golden("implicit_fixed.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}-{c['kw']}")
def test_implicit_multistep_single(c):
    """BackwardEuler / AdamsMoulton1..5 / BDF1..6 against the live reference.  Bar 5e-7 relative: the reference's own
    results move by up to 1.4e-7 under a one-ulp change of y0 (Newton stopped at 1.48e-8 with a finite-difference
    Jacobian), see tests/test_kernel_logic_host.py::test_implicit_multistep_vs_reference."""
    kw = {k: (None if v == "None" else v) for k, v in c["kw"].items()}
    s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"], **kw)
    assert s.is_jit and s.IMPLICIT
    assert rel(s.cache.ts, c["init_ts"]) < 1e-15
    t, y = s.run(c["ts"])
    assert rel(y, c["ys"]) < 5e-7
    assert s.t == c["t_end"]
    assert rel(s.y, c["y_end"]) < 5e-7 and rel(s.f, c["f_end"]) < 5e-6
    assert rel(s.cache.ts, c["end_ts"]) < 1e-15


def test_implicit_multistep_ensemble_api_and_failure():
    B = 64
    y0, p = onp.vdp_ensemble(B, seed=1)
    ts = np.linspace(0, 2, 9)
    for cls in ("BDF2", "AdamsMoulton4", "BackwardEuler"):
        s = getattr(nbk, cls)("vdp", 0.0, y0, p, h=0.01)
        t, y = s.run(ts)
        assert y.shape == (B, 9, 2) and (s.status == 0).all()
        for i in (0, 31, 63):
            o = onp.ImplicitMultistep(cls, onp.rhs_vdp, 0.0, y0[i], p[i], h=0.01)
            _, yo = onp.run(o, ts)
            assert rel(y[i], yo) < 5e-7 and s.t[i] == o.t
    # step / skip / traced Python right-hand side (same value flow as the explicit fixed-step solvers)
    s = nbk.BDF3(lambda t, y, k: -k * y, 0.0, np.array([1.0, 2.0]), np.array([0.01, 0.05]), h=0.25, first_stepper_cls=None)
    tt, yy = s.step(n=4)
    o = onp.ImplicitMultistep("BDF3", onp.rhs_decay, 0.0, [1.0, 2.0], np.array([0.01, 0.05]), h=0.25, first_stepper=None)
    to, yo = onp.steps(o, n=4)
    assert rel(tt, to) < 1e-15 and rel(yy, yo) < 5e-7
    s.skip(n=2)
    assert abs(s.t - 1.5) < 1e-14
    # a step whose Newton iteration cannot converge fails the trajectory (the reference raises RuntimeError)
    bad = nbk.BackwardEuler("decay", 0.0, np.array([[1.0, 1.0], [1.0, 1.0]]), np.array([[np.nan, 0.1], [0.1, 0.1]]), h=1.0)
    with pytest.raises(RuntimeError):
        bad.run([3.0])
    assert list(bad.status) == [5, 0]
    # n > 8 is rejected with a clear message
    with pytest.raises(ValueError, match="n <= 8"):
        nbk.BDF2("decay", 0.0, np.ones(9), np.array([0.1]), h=0.1)


@pytest.mark.parametrize("cls", ["RungeKutta45", "RungeKutta23"])
def test_grid_point_exactly_at_a_step_end(cls):
    """A grid time equal to the end of a step returns y_new itself (RkInterpolator, explicit.py:371-403): the kernels patch
    that point after the per-point loop."""
    y0, p = onp.lorenz_ensemble(64, seed=4)
    kw = dict(rtol=1e-6, atol=1e-9)
    a = getattr(nbk, cls)("lorenz", 0.0, y0, p, **kw)
    a.run([0.7])
    t_end, y_end = a.t, a.y
    for b in (0, 17, 63):
        g = getattr(nbk, cls)("lorenz", 0.0, y0[b], p[b], **kw)
        _, out = g.run([0.3, 0.5 * (0.3 + t_end[b]), t_end[b]])
        assert np.array_equal(out[2], y_end[b])
