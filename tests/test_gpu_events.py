"""GPU parity tests of run_events against the live reference's golden vectors
(tests/golden/events.json) and the NumPy oracle's restatement of nbkode/event_handler.py."""

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b)))) if a.size else 0.0


@pytest.mark.parametrize("c", load_golden("events.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{'+'.join(c['events'])}")
def test_events_single_vs_reference(c):
    s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), **c["kw"])
    t, y, te, ye = s.run_events(c["ts"], [onp.EVENTS[e] for e in c["events"]])
    assert len(t) == len(c["t"]) and y.shape == (len(c["t"]), len(c["y0"]))
    tol_y = 2e-6 if c["rhs"] == "lorenz" else 1e-7
    assert rel(t, c["t"]) < 1e-7
    assert rel(y, c["y"]) < tol_y
    assert [len(x) for x in te] == [len(x) for x in c["t_events"]]
    for a, b, ya, yb, al in zip(te, c["t_events"], ye, c["y_events"], c["y_events_aliased"]):
        assert rel(a, b) < 1e-6
        for u, v, aliased in zip(ya, yb, al):
            if not aliased:
                assert rel(u, v) < 1e-5
    assert abs(s.t - c["t_end"]) < 1e-6 * max(1.0, abs(c["t_end"]))


def test_events_ensemble_vs_oracle():
    B = 48
    y0, p = onp.lorenz_ensemble(B, seed=17)
    ts = np.linspace(0.5, 6.0, 12)
    events = [onp.ev_y0, onp.ev_z25_up, onp.ev_x_m8_terminal]
    s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-8, atol=1e-10)
    t, y, te, ye = s.run_events(ts, events)
    assert y.shape == (B, 12, 3) and len(te) == 3 and s.event_counts.shape == (B, 3)
    n_stopped = 0
    for i in range(B):
        o = onp.AdaptiveRK("RungeKutta45", onp.rhs_lorenz, 0.0, y0[i], p[i], rtol=1e-8, atol=1e-10)
        to, yo, teo, yeo = onp.run_events(o, ts, events)
        k = len(to)
        ti = t[i] if t.ndim == 2 else t
        assert np.isnan(y[i, k:]).all() and np.isfinite(y[i, :k]).all()
        assert rel(ti[:k], to) < 1e-7 and rel(y[i, :k], yo) < 1e-5
        n_stopped += k < len(ts) or to[-1] != ts[k - 1]
        for j in range(3):
            assert s.event_counts[i, j] == len(teo[j]), (i, j)
            if teo[j]:
                assert rel(te[j][i, :len(teo[j])], teo[j]) < 1e-6
                assert rel(ye[j][i, :len(teo[j])], np.array(yeo[j])) < 1e-4
            assert np.isnan(te[j][i, len(teo[j]):]).all()
    assert n_stopped > 0 and t.ndim == 2


def test_events_cuda_source_params_and_errors():
    y0, p = onp.vdp_ensemble(256, seed=3)
    src = """
    NBK_EVENT nbk_event(int k, double t, const double* y, const double* p) {
        return k == 0 ? y[0] : y[1] - 0.25 * p[0];   // the second event uses the trajectory's own parameter
    }
    """
    ev = nbk.CudaEvents(src, 2, terminal=[False, False], direction=[0, 1])
    s = nbk.RungeKutta23("vdp", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y, te, ye = s.run_events([5.0, 10.0], ev)
    assert t.shape == (2,) and y.shape == (256, 2, 2)
    # every recorded event satisfies its equation on the returned state (bisection tolerance 1.48e-8 on g)
    cnt = s.event_counts
    for i in (0, 100, 255):
        for j in range(cnt[i, 0]):
            assert abs(ye[0][i, j, 0]) < 5e-8
        for j in range(cnt[i, 1]):
            assert abs(ye[1][i, j, 1] - 0.25 * p[i, 0]) < 5e-8
    assert cnt.sum() > 256
    # the same events as a Python callable with the (t, y, p) signature
    def g0(t, y):
        return y[0]

    def g1(t, y, p):
        return y[1] - 0.25 * p[0]

    g1.direction = 1
    s2 = nbk.RungeKutta23("vdp", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t2, y2, te2, ye2 = s2.run_events([5.0, 10.0], [g0, g1])
    assert np.array_equal(s2.event_counts, cnt) and np.array_equal(y2, y)
    assert np.allclose(te2[0], te[0], equal_nan=True, rtol=0, atol=1e-12)
    # capacity overflow is reported, not silently truncated
    s3 = nbk.RungeKutta23("vdp", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    with pytest.raises(RuntimeError, match="max_events"):
        s3.run_events([40.0], ev, max_events=2)
    # no events -> plain run
    s4 = nbk.RungeKutta23("vdp", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    out = s4.run_events([5.0, 10.0], None)
    assert out[2] == [] and out[3] == [] and rel(out[1], y) < 1e-13  # (the plain run divides 1 / H once per step)
    # resumed: a second run_events call continues from the state the first one left
    t5, y5, te5, ye5 = s.run_events([12.0], ev)
    assert y5.shape == (256, 1, 2) and np.isfinite(y5).all()


def test_events_docs_cannon_example():
    """docs/events.rst:41-60: cannon fired upward, terminal event on impact, direction -1."""
    def upward_cannon(t, y):
        return np.array([y[1], -0.5])

    def hit_ground(t, y):
        return y[0]

    hit_ground.terminal = True
    hit_ground.direction = -1
    s = nbk.RungeKutta45(upward_cannon, 0.0, np.array([0.0, 10.0]))
    t, y, te, ye = s.run_events(100, events=hit_ground)
    assert len(te) == 1 and len(te[0]) == 1 and abs(te[0][0] - 40.0) < 1e-6
    assert len(t) == 1 and abs(t[0] - 40.0) < 1e-6 and abs(y[0, 0]) < 1e-7


@pytest.mark.parametrize("c", load_golden("implicit_extra.json")["events"], ids=lambda c: c["cls"])
def test_events_with_implicit_solvers(c):
    s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), **c["kw"])
    t, y, te, ye = s.run_events(c["ts"], [onp.EVENTS[e] for e in c["events"]])
    assert rel(t, c["t"]) < 1e-12 and rel(y, c["y"]) < 5e-7  # Newton-tolerance bar of the implicit families
    assert [len(x) for x in te] == [len(x) for x in c["t_events"]]
    for a, b in zip(te, c["t_events"]):
        assert rel(a, b) < 1e-6


@pytest.mark.parametrize("seed", range(24))
def test_events_random_sets_vs_oracle(seed):
    """Seeded random combinations: solver class (adaptive and fixed-step) x right-hand side x a subset of the oracle's
    event functions (directional, terminal, time events) x a small ensemble, every trajectory against the NumPy
    restatement of event_handler.py: the same events in the same order, the same truncation by terminal events."""
    rng = np.random.default_rng(7000 + seed)
    rname = str(rng.choice(["lorenz", "vdp"]))
    names = ["y0", "y1_down", "time", "y0_up_terminal"] + (["z25_up", "x_m8_terminal"] if rname == "lorenz" else [])
    events = [onp.EVENTS[e] for e in rng.choice(names, size=int(rng.integers(1, 4)), replace=False)]
    B = int(rng.integers(2, 7))
    y0, p = (onp.lorenz_ensemble if rname == "lorenz" else onp.vdp_ensemble)(B, seed=int(rng.integers(0, 1 << 30)))
    fn = onp.RHS[rname]
    cls = str(rng.choice(["RungeKutta45", "RungeKutta23", "DOP853", "RungeKutta4", "AdamsBashforth2", "ForwardEuler"]))
    if cls in ("RungeKutta45", "RungeKutta23", "DOP853"):
        kw = dict(rtol=1e-8, atol=1e-10)
        mk = lambda b: onp.AdaptiveRK(cls, fn, 0.0, y0[b], p[b], **kw)  # noqa: E731
    elif cls == "RungeKutta4":
        kw = dict(h=0.002)
        mk = lambda b: onp.FixedRK(cls, fn, 0.0, y0[b], p[b], **kw)  # noqa: E731
    else:
        kw = dict(h=0.0005)
        mk = lambda b: onp.AdamsBashforth(1 if cls == "ForwardEuler" else 2, fn, 0.0, y0[b], p[b], **kw)  # noqa: E731
    ts = np.sort(rng.uniform(0.2, 3.0 if rname == "lorenz" else 6.0, int(rng.integers(2, 9))))
    s = getattr(nbk, cls)(rname, 0.0, y0, p, **kw)
    t, y, te, ye = s.run_events(ts, events)
    # Lorenz separates nearby trajectories by e^(0.9 t): event times inherit the amplified rounding differences
    tol_t, tol_y = (2e-5, 2e-3) if rname == "lorenz" else (1e-6, 1e-4)
    for i in range(B):
        to, yo, teo, yeo = onp.run_events(mk(i), ts, events)
        k = len(to)
        ti = t[i] if t.ndim == 2 else t
        assert np.isnan(y[i, k:]).all() and np.isfinite(y[i, :k]).all(), (cls, rname, i)
        assert rel(ti[:k], to) < tol_t and rel(y[i, :k], np.asarray(yo).reshape(k, -1)) < tol_y, (cls, rname, i)
        for j in range(len(events)):
            assert s.event_counts[i, j] == len(teo[j]), (cls, rname, [e.__name__ for e in events], i, j)
            if teo[j]:
                assert rel(te[j][i, :len(teo[j])], teo[j]) < tol_t
                assert rel(ye[j][i, :len(teo[j])], np.array(yeo[j])) < tol_y
