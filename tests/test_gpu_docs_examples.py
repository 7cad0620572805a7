"""The reference's documentation examples, run as written through nbkode_b200 on the GPU and checked against the
NumPy restatement of the reference (docs/tutorial.rst, docs/migration.rst, docs/events.rst, README.rst).

A user who follows the tutorial types these lines; each one has to work unchanged: scalar y0 and scalar params,
lists for y0 / params, tuple-returning right-hand sides, run(scalar), run resumed on a later grid, step(n=),
step(upto_t=), skip(upto_t=), the state properties, get_solvers / get_groups queries.
"""

import math
from math import sin

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbkode  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402

RTOL, ATOL = 1e-3, 1e-6  # the defaults of the reference (nbkode/core.py:696-705)


def within(y, ref, slack=1.0):
    return np.all(np.abs(np.asarray(y) - np.asarray(ref)) <= slack * (RTOL * np.abs(ref) + ATOL))


def test_tutorial_run_and_resume():
    """docs/tutorial.rst:28-59."""
    def rhs(t, y):
        return -0.1 * y

    y0 = 1.
    t0 = 0
    solver = nbkode.RungeKutta45(rhs, t0, y0)
    ts = np.linspace(0, 10, 100)
    ts, ys = solver.run(ts)
    o = onp.AdaptiveRK("RungeKutta45", rhs, t0, y0)
    _, yo = onp.run(o, np.linspace(0, 10, 100))
    assert ys.shape == (100, 1) and within(ys, yo) and np.max(np.abs(ys - yo)) < 1e-12
    assert np.allclose(ys[:, 0], np.exp(-0.1 * ts), rtol=1e-3)
    ts2 = np.linspace(20, 40, 100)
    ts2, ys2 = solver.run(ts2)
    _, yo2 = onp.run(o, np.linspace(20, 40, 100))
    assert ys2.shape == (100, 1) and np.max(np.abs(ys2 - yo2)) < 1e-12


def test_tutorial_step_and_skip():
    """docs/tutorial.rst:73-98."""
    def rhs(t, y):
        return -0.1 * y

    solver = nbkode.RungeKutta45(rhs, 0, 1.)
    ts, ys = solver.step(n=10)
    o = onp.AdaptiveRK("RungeKutta45", rhs, 0, 1.)
    to, yo = onp.steps(o, n=10)
    assert ts.shape == (10,) and ys.shape == (10, 1)
    assert np.max(np.abs(ts - to)) < 1e-11 and np.max(np.abs(ys - yo)) < 1e-12
    solver = nbkode.RungeKutta45(rhs, 0, 1.)
    solver.skip(upto_t=10)
    # "until the next timepoint will go beyond upto_t" (core.py:269-313, _skip: `while step(t_end, ...)`): the solver
    # stops at the last step whose successor would pass t = 10 -- with the default tolerances on this problem the step
    # size grows tenfold per step, so that is well before 10
    o = onp.AdaptiveRK("RungeKutta45", rhs, 0, 1.)
    onp.steps(o, upto_t=10)
    assert abs(solver.t - o.t) < 1e-10 and solver.t <= 10 and solver.t + solver.h > 10
    assert np.max(np.abs(solver.y - o.y)) < 1e-12 and np.max(np.abs(solver.f - o.f)) < 1e-12
    # "You can also use upto_t here": the last returned step is the last one that does not pass upto_t
    solver = nbkode.RungeKutta45(rhs, 0, 1.)
    ts, ys = solver.step(upto_t=5.0)
    assert ts[-1] <= 5.0 and len(ts) >= 2


def test_tutorial_parameters_scalar_and_vector():
    """docs/tutorial.rst:119-172: params as a plain float, y0 / params as lists, the three spellings of a 2-vector
    right-hand side."""
    def rhs(t, y, p):
        return p * y

    solver = nbkode.RungeKutta45(rhs, 0, 1., params=-0.1)
    ts, ys = solver.run(np.linspace(0, 10, 100))
    o = onp.AdaptiveRK("RungeKutta45", rhs, 0, 1., np.asarray(-0.1))
    _, yo = onp.run(o, np.linspace(0, 10, 100))
    assert np.max(np.abs(ys - yo)) < 1e-12

    y0 = [1., 2.]
    p = [-0.1, -0.5]
    solver = nbkode.RungeKutta45(rhs, 0, y0, params=p)
    ts, ys = solver.run(np.linspace(0, 10, 100))
    o = onp.AdaptiveRK("RungeKutta45", rhs, 0, y0, np.asarray(p))
    _, yo = onp.run(o, np.linspace(0, 10, 100))
    assert ys.shape == (100, 2) and np.max(np.abs(ys - yo)) < 1e-12

    def rhs_asarray(t, y):
        return np.asarray([-0.1 * y[0], -0.5 * y[1]])

    def rhs_tuple(t, y):
        return -0.1 * y[0], -0.5 * y[1]

    for f in (rhs_asarray, rhs_tuple):
        _, yf = nbkode.RungeKutta45(f, 0, y0).run(np.linspace(0, 10, 100))
        assert np.max(np.abs(yf - yo)) < 1e-12


def test_tutorial_solver_queries():
    """docs/tutorial.rst:183-221."""
    assert len(nbkode.get_solvers()) == 26
    fixed = nbkode.get_solvers(fixed_step=True)
    assert nbkode.RungeKutta45 not in fixed and nbkode.ForwardEuler in fixed
    explicit = nbkode.get_solvers(implicit=False)
    assert nbkode.BackwardEuler not in explicit and nbkode.RungeKutta45 in explicit
    assert set(nbkode.get_solvers("euler")) == {nbkode.ForwardEuler, nbkode.BackwardEuler}
    both = nbkode.get_solvers("Adams-Bashforth", "Euler")
    assert nbkode.AdamsBashforth3 in both and nbkode.ForwardEuler in both and nbkode.RungeKutta45 not in both
    assert "Runge-Kutta" in nbkode.get_groups() and "Euler" in nbkode.get_groups()


def test_migration_from_solve_ivp():
    """docs/migration.rst:49-118 (the scipy side is run too)."""
    from scipy.integrate import solve_ivp

    def exponential_decay(t, y):
        return -0.5 * y

    sol = nbkode.RungeKutta45(exponential_decay, 0, [2, 4, 8])
    t, y = sol.run(10)
    out = solve_ivp(exponential_decay, [0, 10], [2, 4, 8])
    assert np.ndim(t) == 1 and len(t) == 1 and t[0] == 10 and y.shape == (1, 3)
    # same method, same default tolerances, same initial step: the two agree far inside the tolerance
    assert np.allclose(y[0], out.y[:, -1], rtol=5e-4)
    t2, y2 = sol.run(20)
    assert t2[0] == 20 and np.allclose(y2[0], np.array([2, 4, 8]) * np.exp(-10.0), rtol=5e-2, atol=1e-5)
    sol = nbkode.RungeKutta45(exponential_decay, 0, [2, 4, 8])
    t, y = sol.run([0, 1, 2, 4, 10])
    out = solve_ivp(exponential_decay, [0, 10], [2, 4, 8], t_eval=[0, 1, 2, 4, 10])
    assert y.shape == (5, 3) and np.allclose(y, out.y.T, rtol=2e-3, atol=1e-5)
    o = onp.AdaptiveRK("RungeKutta45", exponential_decay, 0, [2., 4., 8.])
    _, yo = onp.run(o, [0, 1, 2, 4, 10])
    assert np.max(np.abs(y - yo)) < 1e-12

    def upward_cannon(t, y):
        return [y[1], -0.5]

    def hit_ground(t, y):
        return y[0]

    hit_ground.terminal = True
    hit_ground.direction = -1
    sol = nbkode.RungeKutta45(upward_cannon, 0, [0, 10])
    t, y, t_events, y_events = sol.run_events(100, events=hit_ground)
    out = solve_ivp(upward_cannon, [0, 100], [0, 10], events=hit_ground)
    assert abs(t_events[0][0] - out.t_events[0][0]) < 1e-6 and abs(t_events[0][0] - 40.0) < 1e-6
    assert np.allclose(y_events[0][0], out.y_events[0][0], atol=1e-6)


def test_readme_example():
    """README.rst: the quick start, with ForwardEuler and the adaptive default side by side."""
    def func(t, y, k):
        return -k * y

    t0, y0 = 0., 1.
    for cls, ref in ((nbkode.ForwardEuler, lambda: onp.AdamsBashforth(1, func, t0, y0, np.array([1.]))),
                     (nbkode.RungeKutta45, lambda: onp.AdaptiveRK("RungeKutta45", func, t0, y0, np.array([1.])))):
        solver = cls(func, t0, y0, params=(1,))
        ts, ys = solver.run([0, 5, 10])
        _, yo = onp.run(ref(), [0, 5, 10])
        assert ys.shape == (3, 1) and np.max(np.abs(ys - yo)) < 1e-12


def _pendulum(t, y, p):
    dy = np.empty(2)
    dy[0] = y[1]
    dy[1] = -p[0] * sin(y[0]) - p[1] * y[1] + 0.2 * math.cos(t)
    return dy


def test_numba_spelling_and_njit_dispatchers():
    """A right-hand side written the way numba users write them (`from math import sin`, `np.empty(2)` filled element by
    element) and the same function already wrapped by `numba.njit` -- the reference accepts both (core.py:125-132) --
    give the same ensemble results as the NumPy restatement calling the very same callables."""
    numba = pytest.importorskip("numba")
    jitted = numba.njit(_pendulum)
    rng = np.random.default_rng(12)
    B = 300
    y0 = np.stack([rng.uniform(-2, 2, B), rng.uniform(-1, 1, B)], 1)
    p = np.stack([rng.uniform(0.5, 2, B), rng.uniform(0.05, 0.5, B)], 1)
    ts = np.linspace(0.5, 10, 20)
    kw = dict(rtol=1e-6, atol=1e-9)
    a = nbkode.RungeKutta45(_pendulum, 0.0, y0, p, **kw)
    b = nbkode.RungeKutta45(jitted, 0.0, y0, p, **kw)
    _, ya = a.run(ts)
    _, yb = b.run(ts)
    assert np.array_equal(ya, yb)
    for i in (0, 1, 150, 299):
        o = onp.AdaptiveRK("RungeKutta45", _pendulum, 0.0, y0[i], p[i], **kw)
        _, yo = onp.run(o, ts)
        assert abs(int(np.asarray(a.n_accepted)[i]) - o.n_accepted) <= 1
        assert np.all(np.abs(ya[i] - yo) <= 1e-6 * np.abs(yo) + 1e-9)
    s = nbkode.RungeKutta4(jitted, 0.0, y0[0], p[0], h=0.01)
    _, y4 = s.run(ts)
    o = onp.FixedRK("RungeKutta4", _pendulum, 0.0, y0[0], p[0], h=0.01)
    _, yo = onp.run(o, ts)
    assert np.max(np.abs(y4 - yo)) < 1e-12
