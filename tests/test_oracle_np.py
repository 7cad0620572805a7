"""Pin the NumPy oracle (oracle/nbk_oracle_np.py) against golden vectors produced by the
live reference (tests/golden/make_golden.py).  The bar here is BIT-EXACT equality: the
restatement uses the reference's own NumPy primitives in the same association order."""

import numpy as np
import pytest

from oracle import nbk_oracle_np as onp

from conftest import load_golden


def _eq(a, b, what=""):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, what
    assert np.array_equal(a, b), f"{what}: max|d|={np.max(np.abs(a - b))}"


def _mk_adaptive(c):
    return onp.AdaptiveRK(c["cls"], onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), **c["kw"])


@pytest.mark.parametrize("c", load_golden("adaptive_single.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_adaptive_single_bitexact(c):
    s = _mk_adaptive(c)
    assert float(s.h) == c["h0"]
    _eq(s.f, c["f0"], "f0")
    t, y = onp.run(s, c["ts"])
    _eq(y, c["ys"], "run(ts)")
    assert s.t == c["t_end"]
    _eq(s.y, c["y_end"], "y_end")
    _eq(s.f, c["f_end"], "f_end")
    assert float(s.h) == c["h_end"]
    _eq(s.K, c["K_end"], "K")
    assert s.cache.ts[0] == c["t_old"]
    _eq(s.cache.ys[0], c["y_old"])
    assert s.n_accepted == c["n_accepted"]


def _first_stepper(kw):
    fs = kw.get("first_stepper_cls", "auto")
    if isinstance(fs, str) and fs.startswith("AdamsBashforth"):
        return int(fs[-1])
    return fs


def _order(cls):
    return 1 if cls == "ForwardEuler" else int(cls[-1])


@pytest.mark.parametrize("c", load_golden("fixed_single.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{c['kw']}")
def test_fixed_single_bitexact(c):
    s = onp.AdamsBashforth(
        _order(c["cls"]), onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"],
        first_stepper=_first_stepper(c["kw"]),
    )
    _eq(s.cache.ts, c["init_ts"], "start-up ts")
    _eq(s.cache.ys, c["init_ys"], "start-up ys")
    _eq(s.cache.fs, c["init_fs"], "start-up fs")
    t, y = onp.run(s, c["ts"])
    _eq(y, c["ys"], "run(ts)")
    assert s.t == c["t_end"]
    _eq(s.y, c["y_end"])
    _eq(s.f, c["f_end"])
    _eq(s.cache.ts, c["end_ts"])


def test_c1_readme():
    for c in load_golden("c1_readme.json"):
        kw = dict(c["kw"])
        if "Runge" in c["cls"]:
            s = onp.AdaptiveRK(c["cls"], onp.rhs_decay, 0.0, 1.0, np.array([0.1]))
        else:
            s = onp.AdamsBashforth(_order(c["cls"]), onp.rhs_decay, 0.0, 1.0, np.array([0.1]), h=kw.get("h"))
        assert float(s.h) == c["h0"]
        t, y = onp.run(s, [0, 5, 10])
        _eq(y, c["ys"], c["cls"])
        assert s.t == c["t_end"] and float(s.h) == c["h_end"]


def test_lorenz_ensemble_rows():
    g = load_golden("lorenz_ensemble_rk45.json")
    y0, p = onp.lorenz_ensemble(g["B_total"], g["seed"])
    for r in g["rows"]:
        i = r["row"]
        _eq(y0[i], r["y0"])
        _eq(p[i], r["p"])
        s = onp.AdaptiveRK("RungeKutta45", onp.rhs_lorenz, 0.0, y0[i], p[i], rtol=g["rtol"], atol=g["atol"])
        assert float(s.h) == r["h0"]
        t, y = onp.run(s, [g["T"]])
        _eq(y, r["ys"])
        assert s.t == r["t_end"] and s.n_accepted == r["n_accepted"] and float(s.h) == r["h_end"]
        _eq(s.y, r["y_end"])


def test_vdp_ensemble_rows():
    g = load_golden("vdp_ensemble.json")
    y0, p = onp.vdp_ensemble(g["B_total"], g["seed"])
    ts = np.linspace(0, 20, g["n_ts"])
    for r in g["rk23"][:4]:
        i = r["row"]
        _eq(y0[i], r["y0"])
        s = onp.AdaptiveRK("RungeKutta23", onp.rhs_vdp, 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
        t, y = onp.run(s, ts)
        _eq(y[g["keep"]], r["ys"])
        assert s.t == r["t_end"] and s.n_accepted == r["n_accepted"]
    for r in g["ab4"][:4]:
        i = r["row"]
        s = onp.AdamsBashforth(4, onp.rhs_vdp, 0.0, y0[i], p[i], h=0.01)
        t, y = onp.run(s, ts)
        _eq(y[g["keep"]], r["ys"])
        assert s.t == r["t_end"]


@pytest.mark.parametrize("c", load_golden("api_semantics.json"), ids=lambda c: c["cls"])
def test_api_semantics(c):
    def mk():
        if "Runge" in c["cls"]:
            # the reference ignores h= for adaptive solvers (core.py:134-137 then :706 overrides it)
            return onp.AdaptiveRK(c["cls"], onp.rhs_decay, 0.0, [1.0], np.array([0.01]))
        return onp.AdamsBashforth(_order(c["cls"]), onp.rhs_decay, 0.0, [1.0], np.array([0.01]), h=0.1, first_stepper=None)

    s = mk()
    ts, ys = onp.steps(s, n=20)
    _eq(ts, c["ts20"])
    _eq(ys, c["ys20"])
    s = mk()
    t, y = onp.steps(s, n=1)
    _eq(t, c["seq"][0][1])
    t, y = onp.steps(s, n=2)
    _eq(y, c["seq"][1][2])
    t, y = onp.steps(s, upto_t=ts[5])
    _eq(t, c["seq"][2][1])
    t, y = onp.steps(s, n=5, upto_t=ts[7])
    _eq(t, c["seq"][3][1])


@pytest.mark.parametrize("c", load_golden("erk_fixed.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_fixed_rk_bitexact(c):
    """Fixed-step explicit Runge-Kutta classes."""
    s = onp.FixedRK(c["cls"], onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"])
    _eq(s.cache.ys, c["init_ys"])
    t, y = onp.run(s, c["ts"])
    _eq(y, c["ys"], "run(ts)")
    assert s.t == c["t_end"]
    _eq(s.y, c["y_end"])
    _eq(s.f, c["f_end"])
    if "step10_t" in c:
        s = onp.FixedRK(c["cls"], onp.rhs_vdp, 0.0, np.array([2.0, 0.0]), np.array([1.0]), h=0.05)
        tt, yy = onp.steps(s, n=10)
        _eq(tt, c["step10_t"]); _eq(yy, c["step10_y"]); _eq(s.K, c["step10_K"])


# ---------------------------------------------------------------- DOP853
_DOP = load_golden("dop853.json")


@pytest.mark.parametrize("c", _DOP["single"], ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_dop853_single_bitexact(c):
    test_adaptive_single_bitexact(c)


def test_dop853_ensemble_rows():
    g = _DOP["ensemble"]
    y0, p = onp.lorenz_ensemble(g["B_total"], g["seed"])
    for r in g["rows"]:
        i = r["row"]
        _eq(y0[i], r["y0"])
        s = onp.AdaptiveRK("DOP853", onp.rhs_lorenz, 0.0, y0[i], p[i], rtol=g["rtol"], atol=g["atol"])
        assert float(s.h) == r["h0"]
        t, y = onp.run(s, [g["T"]])
        _eq(y, r["ys"])
        assert s.t == r["t_end"] and s.n_accepted == r["n_accepted"] and float(s.h) == r["h_end"]
        _eq(s.y, r["y_end"])


def test_dop853_api_semantics():
    c = _DOP["api"]
    mk = lambda: onp.AdaptiveRK("DOP853", onp.rhs_decay, 0.0, [1.0], np.array([0.01]))  # noqa: E731
    s = mk()
    ts, ys = onp.steps(s, n=20)
    _eq(ts, c["ts20"])
    _eq(ys, c["ys20"])
    s = mk()
    t, y = onp.run(s, np.linspace(0, 4, 9))
    _eq(y, c["resume_a"])
    t, y = onp.run(s, c["resume_b_t"])
    _eq(y, c["resume_b"])


# ---------------------------------------------------------------- events
def _mk_any(c):
    cls, rhs, y0, p, kw = c["cls"], onp.RHS[c["rhs"]], np.array(c["y0"]), np.array(c["p"]), c["kw"]
    if cls in ("RungeKutta45", "RungeKutta23", "DOP853"):
        return onp.AdaptiveRK(cls, rhs, 0.0, y0, p, **kw)
    if cls in onp.ERK_TABLEAUX:
        return onp.FixedRK(cls, rhs, 0.0, y0, p, **kw)
    return onp.AdamsBashforth(_order(cls), rhs, 0.0, y0, p, **kw)


@pytest.mark.parametrize("c", load_golden("events.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{'+'.join(c['events'])}")
def test_events_bitexact(c):
    s = _mk_any(c)
    t, y, te, ye = onp.run_events(s, c["ts"], [onp.EVENTS[e] for e in c["events"]])
    _eq(t, c["t"], "t")
    _eq(y, c["y"], "y")
    assert len(te) == len(c["t_events"])
    for a, b, ya, yb, al in zip(te, c["t_events"], ye, c["y_events"], c["y_events_aliased"]):
        _eq(a, b, "t_events")
        for u, v, aliased in zip(ya, yb, al):
            if not aliased:  # see make_golden.py: the reference's own aliasing bug
                _eq(u, v, "y_events")
    assert s.t == c["t_end"]


# ---------------------------------------------------------------- implicit multistep
@pytest.mark.parametrize("c", load_golden("implicit_fixed.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}-{c['kw']}")
def test_implicit_multistep_bitexact(c):
    s = onp.ImplicitMultistep(c["cls"], onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"],
                              first_stepper=_first_stepper(c["kw"]))
    _eq(s.cache.ts, c["init_ts"], "start-up ts")
    _eq(s.cache.ys, c["init_ys"], "start-up ys")
    t, y = onp.run(s, c["ts"])
    _eq(y, c["ys"], "run(ts)")
    assert s.t == c["t_end"]
    _eq(s.y, c["y_end"])
    _eq(s.f, c["f_end"])
    _eq(s.cache.ts, c["end_ts"])


_IMX = load_golden("implicit_extra.json")


@pytest.mark.parametrize("c", _IMX["events"], ids=lambda c: c["cls"])
def test_implicit_events_bitexact(c):
    s = onp.ImplicitMultistep(c["cls"], onp.RHS[c["rhs"]], 0.0, np.array(c["y0"]), np.array(c["p"]), **c["kw"])
    t, y, te, ye = onp.run_events(s, c["ts"], [onp.EVENTS[e] for e in c["events"]])
    _eq(t, c["t"]); _eq(y, c["y"])
    for a, b, ya, yb, al in zip(te, c["t_events"], ye, c["y_events"], c["y_events_aliased"]):
        _eq(a, b, "t_events")
        for u, v, aliased in zip(ya, yb, al):
            if not aliased:
                _eq(u, v, "y_events")


@pytest.mark.parametrize("c", _IMX["api"], ids=lambda c: c["cls"])
def test_implicit_api_semantics(c):
    mk = lambda: onp.ImplicitMultistep(c["cls"], onp.rhs_decay, 0.0, [1.0], np.array([0.01]), h=0.1, first_stepper=None)  # noqa: E731
    s = mk()
    ts, ys = onp.steps(s, n=20)
    _eq(ts, c["ts20"]); _eq(ys, c["ys20"])
    s = mk()
    t, y = onp.run(s, np.linspace(0, 4, 9))
    _eq(y, c["resume_a"])
    t, y = onp.run(s, c["resume_b_t"])
    _eq(y, c["resume_b"])


# ---------------------------------------------------------------- right-hand sides / events with run-time branches
_BR = load_golden("branching.json")


@pytest.mark.parametrize("c", _BR["adaptive"], ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_branching_adaptive_bitexact(c):
    test_adaptive_single_bitexact(c)


@pytest.mark.parametrize("c", _BR["fixed"], ids=lambda c: f"{c['cls']}-{c['rhs']}")
def test_branching_fixed_bitexact(c):
    s = _mk_any({**c, "kw": {"h": c["h"]}})
    _eq(s.cache.ys, c["init_ys"])
    t, y = onp.run(s, c["ts"])
    _eq(y, c["ys"], "run(ts)")
    assert s.t == c["t_end"]
    _eq(s.y, c["y_end"])
    _eq(s.f, c["f_end"])


@pytest.mark.parametrize("c", _BR["events"], ids=lambda c: f"{c['cls']}-{'+'.join(c['events'])}")
def test_branching_events_bitexact(c):
    test_events_bitexact(c)
