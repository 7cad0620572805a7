"""Port of the reference's API-contract tests (nbkode/testsuite/test_solver.py::
test_f1_public_api, test_solver_fs.py::test_first_stepper, test_against_scipy.py) to the
GPU engine, for the explicit solvers it ships.  Unbatched inputs must give the reference's
shapes, values (checked against golden vectors / the NumPy oracle) and exceptions."""

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_equal

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbkode  # noqa: E402

solvers = nbkode.get_solvers()

y0_1 = np.atleast_1d(1.0)
y0_2 = np.atleast_1d([1.0, 2.0])


def f1(t, x, k):
    return -k * x


@pytest.mark.parametrize("solver", solvers, ids=lambda s: s.__name__)
def test_f1_public_api(solver):
    kwargs = dict(h=0.1)
    if issubclass(solver, nbkode.multistep.Multistep):
        kwargs["first_stepper_cls"] = None

    sol = solver(f1, 0.0, y0_1, params=(0.01,), **kwargs)
    ts, ys = sol.step(n=20)

    # STEP
    sol = solver(f1, 0.0, y0_1, params=(0.01,), **kwargs)
    t, y = sol.step()
    assert_allclose(t, np.atleast_1d(sol.t))
    assert_allclose(t, ts[0])
    assert_allclose(y, np.atleast_1d(sol.y))
    assert isinstance(sol.t, float)
    assert sol.y.shape == y0_1.shape
    assert sol.f.shape == y0_1.shape

    t, y = sol.step(n=2)
    assert_allclose(t, ts[1:3])
    assert_allclose(y, ys[1:3])

    t, y = sol.step(upto_t=ts[5])
    assert len(t) == 3
    assert_allclose(t[-2:], ts[4:6])
    assert_allclose(y[-2:], ys[4:6])

    t, y = sol.step(n=5, upto_t=ts[7])
    assert len(t) == 2
    assert_allclose(t[-2:], ts[6:8])
    assert_allclose(y[-2:], ys[6:8])

    t, y = sol.step(n=1, upto_t=ts[-1])
    assert len(t) == 1
    assert sol.t == ts[8]

    # SKIP
    sol.skip()
    assert sol.t == ts[9]
    sol.skip(n=2)
    assert sol.t == ts[11]
    sol.skip(upto_t=0.5 * (ts[13] + ts[14]))
    assert sol.t == ts[13]
    sol.skip(n=5, upto_t=0.5 * (ts[15] + ts[16]))
    assert sol.t == ts[15]
    sol.skip(n=1, upto_t=ts[-1])
    assert sol.t == ts[16]

    # t_bound
    sol = solver(f1, 0.0, y0_1, params=(0.01,), t_bound=ts[0], **kwargs)
    with pytest.raises(ValueError):
        sol.step(upto_t=ts[2])
    with pytest.raises(ValueError):
        sol.step(n=3, upto_t=ts[2])
    with pytest.raises(ValueError):
        sol.skip(upto_t=ts[2])
    with pytest.raises(ValueError):
        sol.skip(n=3, upto_t=ts[2])
    with pytest.raises(RuntimeError):
        sol.step(n=2)
    with pytest.raises(RuntimeError):
        sol.skip(n=2)

    sol = solver(f1, 0.0, y0_1, params=(0.01,), t_bound=10_000_000, **kwargs)
    trev_ = np.linspace(0, 10, 20)[::-1]
    trev, yrev = sol.run(trev_)
    assert_allclose(trev_, trev)
    assert yrev.shape == (len(trev_),) + y0_1.shape

    sol = solver(f1, 0.0, y0_1, params=(0.01,), t_bound=10_000_000, **kwargs)
    tvec = np.linspace(0, 10, 20)
    t, y = sol.run(tvec)
    assert_equal(t, trev[::-1])
    assert_equal(y, yrev[::-1])

    with pytest.raises(ValueError):
        sol.run(0)
    with pytest.raises(ValueError):
        sol.run(10_000_000 + 1)


@pytest.mark.parametrize("c", load_golden("api_semantics.json"), ids=lambda c: c["cls"])
def test_api_values_match_reference(c):
    """Same call sequence as test_f1_public_api, values compared with the live reference."""
    cls = getattr(nbkode, c["cls"])
    kw = {"h": 0.1}
    if cls.FIXED_STEP:
        kw["first_stepper_cls"] = None
    tol = 1e-12 if cls.FIXED_STEP else 1e-9
    mk = lambda **extra: cls(f1, 0.0, np.array([1.0]), params=np.array([0.01]), **kw, **extra)  # noqa: E731
    s = mk()
    ts, ys = s.step(n=20)
    assert_allclose(ts, c["ts20"], rtol=tol, atol=0)
    assert_allclose(ys, c["ys20"], rtol=tol, atol=0)
    s = mk()
    names = [q[0] for q in c["seq"]]
    t, y = s.step()
    assert_allclose(t, c["seq"][names.index("step")][1], rtol=tol)
    t, y = s.step(n=2)
    assert_allclose(y, c["seq"][names.index("step_n2")][2], rtol=tol)
    t, y = s.step(upto_t=ts[5])
    assert_allclose(t, c["seq"][names.index("step_upto5")][1], rtol=tol)
    t, y = s.step(n=5, upto_t=ts[7])
    assert_allclose(t, c["seq"][names.index("step_n5_upto7")][1], rtol=tol)
    t, y = s.step(n=1, upto_t=ts[-1])
    assert_allclose(t, c["seq"][names.index("step_n1_uptolast")][1], rtol=tol)
    s.skip()
    assert_allclose(s.t, c["seq"][names.index("skip")][1], rtol=tol)
    s.skip(n=2)
    assert_allclose(s.t, c["seq"][names.index("skip_n2")][1], rtol=tol)
    s = mk(t_bound=10_000_000)
    t, y = s.run(np.linspace(0, 10, 20)[::-1])
    assert_allclose(t, c["run_rev_t"])
    assert_allclose(y, c["run_rev_y"], rtol=max(tol, 1e-10))
    s = mk()
    ta, ya = s.run(np.linspace(0, 4, 9))
    assert_allclose(ya, c["resume_a"], rtol=max(tol, 1e-10))
    tb, yb = s.run(np.array(c["resume_b_t"]))
    assert_allclose(yb, c["resume_b"], rtol=max(tol, 1e-10))


@pytest.mark.parametrize("solver", [nbkode.AdamsBashforth3, nbkode.AdamsBashforth5], ids=lambda s: s.__name__)
def test_first_stepper(solver):
    sol = solver(f1, 0.0, y0_1, params=(0.01,), h=0.01, first_stepper_cls=None)
    assert sol.t == 0.0
    sol.step()
    assert sol.t == sol.h
    for fs in ("AdamsBashforth1", nbkode.AdamsBashforth1, "RungeKutta45", nbkode.RungeKutta45):
        sol = solver(f1, 0.0, y0_1, params=(0.01,), h=0.01, first_stepper_cls=fs)
        assert sol.t == (sol.ORDER - 1) * sol.h


def test_fixed_step_loose_accuracy():
    # nbkode/testsuite/test_solver_fs.py:42-58 (the reference's own, very loose, bar)
    solver = nbkode.AdamsBashforth5(f1, 0.0, y0_1, params=(0.01,), h=0.01)
    solver.skip(upto_t=10)
    assert_allclose(solver.y, np.exp(-0.01 * 10), atol=0.1)
    solver = nbkode.AdamsBashforth5(f1, 0.0, y0_2, params=(0.01, 0.05), h=0.01)
    solver.skip(upto_t=10)
    assert_allclose(solver.y[0], np.exp(-0.01 * 10), rtol=0.15)
    assert_allclose(solver.y[1], 2.0 * np.exp(-0.05 * 10), rtol=0.15)


def exponential2(t, x):
    return np.asarray([-0.01, -0.05]) * x


@pytest.mark.parametrize("name", ["RK23", "RK45", "DOP853"])
def test_against_scipy_lockstep(name):
    """nbkode/testsuite/test_against_scipy.py:64-83: lock-step step() against SciPy."""
    from scipy import integrate

    nb_cls = {"RK23": nbkode.RungeKutta23, "RK45": nbkode.RungeKutta45, "DOP853": nbkode.DOP853}[name]
    nb = nb_cls(exponential2, 0, y0_2)
    sp = getattr(integrate, name)(exponential2, 0, y0_2, t_bound=30)
    assert_allclose(nb.f, sp.f)
    assert_allclose(nb.h, sp.h_abs)
    while True:
        nb.step()
        sp.step()
        if sp.status != "running":
            break
        assert_allclose(nb.t, sp.t)
        assert_allclose(nb.y, sp.y)
        assert_allclose(nb.f, sp.f)
        assert_allclose(nb.h, sp.h_abs)
        assert_allclose(nb.K, sp.K)


@pytest.mark.parametrize("name", ["RK45", "DOP853"])
def test_interpolate_against_scipy(name):
    """nbkode/testsuite/test_against_scipy.py:84-97 (test_interpolate; the reference excludes RK23, whose dense
    output does not match SciPy's -- we mirror the reference's)."""
    from scipy import integrate

    nb_cls = {"RK23": nbkode.RungeKutta23, "RK45": nbkode.RungeKutta45, "DOP853": nbkode.DOP853}[name]
    t_eval = np.linspace(0, 300, 160)
    nb_t, nb_y = nb_cls(exponential2, 0, y0_2, t_bound=500).run(t_eval)
    sp = integrate.solve_ivp(exponential2, [0, 500], y0_2, t_eval=t_eval, method=name)
    assert_allclose(nb_t, sp.t)
    assert_allclose(nb_y, sp.y.T)


def test_interpolate_method_and_errors():
    s = nbkode.RungeKutta45("vdp", 0.0, [2.0, 0.0], [1.0])
    s.step(n=3)
    lo, hi = s.cache.ts
    mid = s.interpolate(0.5 * (lo + hi))
    assert mid.shape == (2,)
    assert_allclose(s.interpolate(hi), s.y)
    with pytest.raises(ValueError):
        s.interpolate(lo - 1e-3)
    with pytest.raises(ValueError):
        s.interpolate(hi + 1e-3)
    with pytest.raises(nbkode.CompileError):
        nbkode.RungeKutta45(nbkode.CudaRHS("NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy) { dy[0] = nope; }"), 0.0, [1.0])
