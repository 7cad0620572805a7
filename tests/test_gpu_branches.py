"""GPU parity of right-hand sides and event functions with run-time branches (`if` on the state and the time,
Python's min / max) -- the reference compiles them with numba as they are (nbkode/core.py:125-132); here the tracer
explores their control-flow paths (nbkode_b200/rhs.py, tests/test_trace_branches.py) and the branch is taken on
the device.

Checked against tests/golden/branching.json (single trajectories of the LIVE reference, make_golden.py
`branching`) and, for ensembles, against the bit-exact NumPy restatement running the same Python callable.
"""

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402

G = load_golden("branching.json")
FIXED_TOL = 1e-12


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b)))) if a.size else 0.0


def within(y, ref, rtol, atol, slack=1.0):
    return np.abs(np.asarray(y) - np.asarray(ref)) <= slack * (rtol * np.abs(ref) + atol)


@pytest.mark.parametrize("c", G["adaptive"], ids=lambda c: f"{c['cls']}-{len(c['ts'])}-{c['kw']['rtol']}")
def test_branching_adaptive_single_vs_reference(c):
    kw = dict(c["kw"])
    s = getattr(nbk, c["cls"])(onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), **kw)
    assert s.is_jit
    assert rel(float(s.h), c["h0"]) < 1e-13 and rel(s.f, c["f0"]) < 1e-14
    t, y = s.run(c["ts"])
    assert abs(s.n_accepted - c["n_accepted"]) <= 1
    assert within(y, c["ys"], kw["rtol"], kw["atol"]).all()
    assert within(s.y, c["y_end"], kw["rtol"], kw["atol"]).all()
    if s.n_accepted == c["n_accepted"]:  # same step sequence: the dense output agrees far inside the bar
        assert rel(y, c["ys"]) < 1e-9


@pytest.mark.parametrize("c", G["fixed"], ids=lambda c: c["cls"])
def test_branching_fixed_single_vs_reference(c):
    s = getattr(nbk, c["cls"])(onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"])
    t, y = s.run(c["ts"])
    assert rel(y, c["ys"]) <= FIXED_TOL and s.t == c["t_end"]
    assert rel(s.y, c["y_end"]) <= FIXED_TOL and rel(s.f, c["f_end"]) <= FIXED_TOL


@pytest.mark.parametrize("c", G["events"], ids=lambda c: c["cls"])
def test_branching_events_vs_reference(c):
    s = getattr(nbk, c["cls"])(onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), **c["kw"])
    t, y, te, ye = s.run_events(c["ts"], [onp.EVENTS[e] for e in c["events"]])
    assert rel(t, c["t"]) < 1e-7 and rel(y, c["y"]) < 1e-7
    assert [len(x) for x in te] == [len(x) for x in c["t_events"]]
    for a, b in zip(te, c["t_events"]):
        assert rel(a, b) < 1e-6


def _ensemble(B, seed=7):
    rng = np.random.default_rng(seed)
    y0 = np.stack([rng.uniform(-2, 2, B), rng.uniform(-1, 1, B)], 1)
    p = np.stack([rng.uniform(1, 3, B), rng.uniform(0.3, 1.2, B), rng.uniform(0.2, 1.5, B)], 1)
    return y0, p


def test_branching_ensemble_vs_oracle():
    """1024 trajectories take different arms at the same time inside one warp; 48 of them against the restatement.

    Adaptive classes to t = 1 only: every kink a step crosses amplifies a last-bit difference of the step sequence
    about tenfold (tests/golden/make_golden.py, branching_section), so the north star's bar is meaningful for a few
    crossings, not for twenty; 95 % of the rows must meet it, the rest stay within 20 x.  The fixed-step classes
    have no controller: 1e-12 to t = 10."""
    B = 1024
    y0, p = _ensemble(B)
    rows = np.r_[0:24, B - 24:B]
    ts = np.linspace(0.1, 1.0, 10)
    for cls, kw in (("RungeKutta45", dict(rtol=1e-6, atol=1e-9)), ("RungeKutta23", dict(rtol=1e-6, atol=1e-9))):
        s = getattr(nbk, cls)(onp.rhs_pwl, 0.0, y0, p, **kw)
        t, y = s.run(ts)
        assert y.shape == (B, ts.size, 2) and np.isfinite(y).all()
        nacc = np.asarray(s.n_accepted)
        ratio = []
        for i in rows:
            o = onp.AdaptiveRK(cls, onp.rhs_pwl, 0.0, y0[i], p[i], **kw)
            _, yo = onp.run(o, ts)
            assert abs(int(nacc[i]) - o.n_accepted) <= 1, (cls, i)
            ratio.append(float(np.max(np.abs(y[i] - yo) / (kw["rtol"] * np.abs(yo) + kw["atol"]))))
        ratio = np.array(ratio)
        print(f"branching ensemble {cls}: err / tol median {np.median(ratio):.2e}, max {ratio.max():.2e}")
        assert np.mean(ratio <= 1.0) >= 0.95 and ratio.max() <= 20.0, (cls, np.sort(ratio)[-5:])
    ts = np.linspace(0.5, 10, 20)
    for cls, h in (("RungeKutta4", 0.01), ("AdamsBashforth4", 0.005)):
        s = getattr(nbk, cls)(onp.rhs_pwl, 0.0, y0, p, h=h)
        t, y = s.run(ts)
        for i in rows[::6]:
            o = onp.FixedRK(cls, onp.rhs_pwl, 0.0, y0[i], p[i], h=h) if cls == "RungeKutta4" else \
                onp.AdamsBashforth(4, onp.rhs_pwl, 0.0, y0[i], p[i], h=h)
            _, yo = onp.run(o, ts)
            assert rel(y[i], yo) <= FIXED_TOL, (cls, i)


def test_branching_vector_rhs_in_the_componentwise_regimes():
    """Scalar decisions (time, parameters, single elements) around a whole-vector expression become selects of
    nbk_rhs_i: sub-warp groups (n = 24) and one CTA per trajectory (n = 300) against the restatement."""

    def switched(t, y, p):
        d = p[0] if t < 0.5 else 2 * p[0]
        out = d * (np.roll(y, -1) - 2 * y + np.roll(y, 1))
        if p[1] > 1.0 and t >= 0.25:
            out = out + p[1] * y * (1 - y)
        if y[0] > y[-1]:           # one element against another; the extra term joins continuously
            out = out - 0.25 * (y[0] - y[-1]) * y
        return out

    rng = np.random.default_rng(2)
    for n, B in ((24, 96), (300, 24)):
        x = np.arange(n) / n
        y0 = 0.5 + 0.4 * np.sin(2 * np.pi * x[None, :] + rng.uniform(0, 6, B)[:, None])
        p = np.stack([rng.uniform(0.5, 2, B), rng.uniform(0.5, 2, B)], 1)
        s = nbk.RungeKutta45(switched, 0.0, y0, p, rtol=1e-6, atol=1e-9)
        t, y = s.run([0.3, 0.6, 1.0])
        nacc = np.asarray(s.n_accepted)
        for i in (0, 1, B // 2, B - 1):
            o = onp.AdaptiveRK("RungeKutta45", switched, 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
            _, yo = onp.run(o, [0.3, 0.6, 1.0])
            assert abs(int(nacc[i]) - o.n_accepted) <= 1
            assert within(y[i], yo, 1e-6, 1e-9).all(), (n, i)


def test_events_with_a_traced_componentwise_rhs():
    """run_events when BOTH the right-hand side (whole-vector Python callable -> nbk_rhs_i) and the events are user
    code, in the sub-warp regime (n = 24), as sub-warp teams (DOP853) and with one CTA per trajectory (n = 300): the
    events section of the team kernels has to see the event prototype although the user right-hand side pulls the
    kernel header in first (csrc/nbk_jit.cpp).  Same events as with the built-in `rd1d`, and against the restatement."""

    def rd(t, y, p):
        return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)

    def gap(t, y):
        return y[5] - y[17]

    def level(t, y):
        if t < 0.3:                      # a branching event function on top
            return y[2] - 0.6
        return y[2] - 0.7

    ts = [0.25, 0.5]
    seen = 0
    for cls, n, B in (("RungeKutta45", 24, 200), ("DOP853", 24, 64), ("RungeKutta23", 300, 48)):
        y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
        p[:, 0] *= 0.05 * (100.0 / n) ** 2
        kw = dict(rtol=1e-6, atol=1e-9)
        u = getattr(nbk, cls)(rd, 0.0, y0, p, **kw)
        b = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
        assert u._regime == b._regime != 0 and u.is_jit
        tu, yu, teu, yeu = u.run_events(ts, [gap, level])
        tb, yb, teb, yeb = b.run_events(ts, [gap, level])
        assert np.array_equal(u.event_counts, b.event_counts) and (n > 24 or u.event_counts.sum() > 0)
        seen += int(u.event_counts.sum())
        assert np.nanmax(np.abs(yu - yb)) < 1e-7
        for j in range(2):
            assert np.array_equal(np.isnan(teu[j]), np.isnan(teb[j]))
            if teu[j].size and np.isfinite(teu[j]).any():
                assert np.nanmax(np.abs(teu[j] - teb[j])) < 1e-6
        for i in range(3):
            o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[i], p[i], **kw)
            t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, [gap, level])
            for j in range(2):
                assert u.event_counts[i, j] == len(te_ref[j]), (cls, i, j)
                if len(te_ref[j]):
                    assert np.abs(teu[j][i, :len(te_ref[j])] - np.array(te_ref[j])).max() < 1e-6
    assert seen > 20


def _relu_loop(t, y, p):
    n = len(y)
    dy = np.empty(n)
    for i in range(n):
        if y[i] > 0:
            dy[i] = -p[0] * y[i]
        else:
            dy[i] = -p[1] * y[i] + 0.05 * (y[(i + 1) % n] - y[i])
    return dy


def test_if_converted_elementwise_branches_on_the_device():
    """Twelve independent element-wise branches (4096 control-flow paths) are traced if-converted (nbkode_b200/_ifconv.py,
    one select each): per-thread kernels (regime="thread") and the default regime against the NumPy restatement running the
    Python function -- the pieces join continuously at y = 0, fixed step to 1e-12."""
    from nbkode_b200 import rhs as nrhs

    n, B = 12, 256
    assert nrhs.trace(_relu_loop, n, (2,), True).if_converted
    rng = np.random.default_rng(31)
    y0 = rng.uniform(-1, 1, (B, n))
    p = np.stack([rng.uniform(0.5, 2, B), rng.uniform(0.5, 2, B)], 1)
    ts = np.linspace(0.2, 2, 10)
    for regime in ("thread", "auto"):
        s = nbk.RungeKutta4(_relu_loop, 0.0, y0, p, h=0.01, regime=regime)
        _, y = s.run(ts)
        for i in (0, 100, 255):
            o = onp.FixedRK("RungeKutta4", _relu_loop, 0.0, y0[i], p[i], h=0.01)
            _, yo = onp.run(o, ts)
            assert rel(y[i], yo) <= FIXED_TOL, (regime, i)
    s = nbk.RungeKutta45(_relu_loop, 0.0, y0, p, rtol=1e-6, atol=1e-9, regime="thread")
    _, y = s.run(ts[:3])
    for i in (0, 255):
        o = onp.AdaptiveRK("RungeKutta45", _relu_loop, 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
        _, yo = onp.run(o, ts[:3])
        assert within(y[i], yo, 1e-6, 1e-9, slack=2.0).all()
