"""Generate tests/golden/*.json by running the LIVE reference (hgrecco/numbakit-ode).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference in its pure-Python mode
(``NBKODE_NONUMBA=1``, nbkode/nbcompat/nb_to_import.py:16; SURVEY.md Appendix C notes it
agrees bit-for-bit with the numba mode for n>=2 systems).  The only shim is the scipy
``select_initial_step`` signature adapter described in SURVEY.md Appendix C (scipy>=1.11
added two positional arguments; with infinite bounds the formula is unchanged).

Floats are written with ``repr`` precision, i.e. they round-trip exactly.
"""

import json
import os
import sys
import warnings

os.environ["NBKODE_NONUMBA"] = "1"
sys.path.insert(0, "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402
import scipy.integrate._ivp.common as _sc  # noqa: E402

import nbkode.core as _core  # noqa: E402

_new = _sc.select_initial_step
_core.select_initial_step = lambda fun, t0, y0, f0, direction, order, rtol, atol: _new(
    fun, t0, y0, np.inf, np.inf, f0, direction, order, rtol, atol
)
import nbkode  # noqa: E402

from oracle import nbk_oracle_np as onp  # noqa: E402  (RHS catalogue + ensemble generators only)


def L(a):
    return np.asarray(a, dtype=float).tolist()


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as fh:
        json.dump(obj, fh, indent=0, separators=(",", ":"))
    print("wrote", name, os.path.getsize(path), "bytes")


def count_to(sol, T):
    n = 0
    while sol.t < T:
        sol.step()
        n += 1
    return n


def adaptive_case(cls, rhs, y0, p, ts, **kw):
    """ctor values, run(ts), end state, and the accepted-step count."""
    rhsf = onp.RHS[rhs]
    s = getattr(nbkode, cls)(rhsf, 0.0, np.array(y0, float), params=np.array(p, float), **kw)
    out = {"cls": cls, "rhs": rhs, "y0": L(y0), "p": L(p), "kw": kw, "ts": L(ts)}
    out["h0"] = float(s.h)
    out["f0"] = L(s.f)
    t, y = s.run(np.array(ts, float))
    out["ys"] = L(y)
    out["t_end"] = float(s.t)
    out["y_end"] = L(s.y)
    out["f_end"] = L(s.f)
    out["h_end"] = float(s.h)
    out["K_end"] = L(s.K)
    out["t_old"] = float(s.cache.ts[0])
    out["y_old"] = L(s.cache.ys[0])
    s2 = getattr(nbkode, cls)(rhsf, 0.0, np.array(y0, float), params=np.array(p, float), **kw)
    out["n_accepted"] = count_to(s2, ts[-1])
    assert s2.t == s.t
    return out


def fixed_case(cls, rhs, y0, p, ts, h, **kw):
    rhsf = onp.RHS[rhs]
    s = getattr(nbkode, cls)(rhsf, 0.0, np.array(y0, float), params=np.array(p, float), h=h, **kw)
    kw_js = {k: (v if isinstance(v, (str, type(None))) else str(v)) for k, v in kw.items()}
    out = {"cls": cls, "rhs": rhs, "y0": L(y0), "p": L(p), "h": h, "kw": kw_js, "ts": L(ts)}
    out["init_ts"] = L(s.cache.ts)
    out["init_ys"] = L(s.cache.ys)
    out["init_fs"] = L(s.cache.fs)
    t, y = s.run(np.array(ts, float))
    out["ys"] = L(y)
    out["t_end"] = float(s.t)
    out["y_end"] = L(s.y)
    out["f_end"] = L(s.f)
    out["end_ts"] = L(s.cache.ts)
    return out


def api_record(cls, kw):
    """Value flow of nbkode/testsuite/test_solver.py::test_f1_public_api for one class."""
    mk = lambda **extra: getattr(nbkode, cls)(onp.rhs_decay, 0.0, np.array([1.0]), params=np.array([0.01]), **kw, **extra)  # noqa: E731
    s = mk()
    ts, ys = s.step(n=20)
    rec = {"cls": cls, "ts20": L(ts), "ys20": L(ys)}
    s = mk()
    seq = []
    t, y = s.step()
    seq.append(["step", L(t), L(y)])
    t, y = s.step(n=2)
    seq.append(["step_n2", L(t), L(y)])
    t, y = s.step(upto_t=ts[5])
    seq.append(["step_upto5", L(t), L(y)])
    t, y = s.step(n=5, upto_t=ts[7])
    seq.append(["step_n5_upto7", L(t), L(y)])
    t, y = s.step(n=1, upto_t=ts[-1])
    seq.append(["step_n1_uptolast", L(t), L(y)])
    s.skip()
    seq.append(["skip", float(s.t)])
    s.skip(n=2)
    seq.append(["skip_n2", float(s.t)])
    s.skip(upto_t=0.5 * (ts[13] + ts[14]))
    seq.append(["skip_upto", float(s.t)])
    s.skip(n=5, upto_t=0.5 * (ts[15] + ts[16]))
    seq.append(["skip_n5_upto", float(s.t)])
    s.skip(n=1, upto_t=ts[-1])
    seq.append(["skip_n1_upto", float(s.t)])
    rec["seq"] = seq
    s = mk(t_bound=10_000_000)
    trev = np.linspace(0, 10, 20)[::-1]
    t, y = s.run(trev)
    rec["run_rev_t"] = L(t)
    rec["run_rev_y"] = L(y)
    # resumed run: two calls give the same numbers as one
    s = mk()
    ta, ya = s.run(np.linspace(0, 4, 9))
    past = 0.5 * (float(s.cache.ts[0]) + float(s.t))  # lies inside the history window
    tb, yb = s.run(np.concatenate(([past], np.linspace(4.5, 8, 8))))
    rec["resume_a"] = L(ya)
    rec["resume_b_t"] = L(tb)
    rec["resume_b"] = L(yb)
    return rec


def dop853_section():
    """DOP853: single trajectories, rows of the C2 Lorenz ensemble, API value flow."""
    from scipy.integrate._ivp import dop853_coefficients as sdc

    from nbkode.runge_kutta import dop853_coefficients as rdc

    # the oracle and the CUDA tableau header read SciPy's copy of Hairer's coefficients: identical arrays
    for name in ("A", "B", "C", "E3", "E5", "D"):
        assert np.array_equal(getattr(sdc, name), getattr(rdc, name)), name
    lor = ([1.0, 1.0, 1.0], [10.0, 28.0, 8 / 3])
    cls = "DOP853"
    single = [
        adaptive_case(cls, "lorenz", *lor, [2.5, 5.0, 10.0], rtol=1e-6, atol=1e-9),
        adaptive_case(cls, "lorenz", *lor, np.linspace(0, 3, 31), rtol=1e-3, atol=1e-6),
        adaptive_case(cls, "lorenz", *lor, np.linspace(0, 5, 501), rtol=1e-9, atol=1e-12),
        adaptive_case(cls, "vdp", [2.0, 0.0], [1.0], [5.0, 10.0, 20.0], rtol=1e-6, atol=1e-9),
        adaptive_case(cls, "vdp", [2.0, 0.0], [5.0], np.linspace(0, 20, 41), rtol=1e-6, atol=1e-9),
        adaptive_case(cls, "decay", [1.0, 2.0], [0.01, 0.05], np.linspace(0, 300, 16)),
        adaptive_case(cls, "decay", [1.0], [0.01], [30.0], first_step=0.5),
        adaptive_case(cls, "vdp", [2.0, 0.0], [2.0], [10.0], rtol=1e-5, atol=1e-8, min_factor=0.3, max_factor=4.0,
                      safety_factor=0.8),
    ]
    rng = np.random.default_rng(11)
    A8 = rng.standard_normal((8, 8)) / np.sqrt(8) - np.eye(8)
    single.append(adaptive_case(cls, "linear", rng.standard_normal(8), A8, [1.0, 10.0], rtol=1e-6, atol=1e-9))
    y0, p = onp.lorenz_ensemble(1_000_000, seed=0)
    rows = []
    for i in list(range(16)) + list(range(999_992, 1_000_000)):
        r = adaptive_case(cls, "lorenz", y0[i], p[i], [10.0], rtol=1e-6, atol=1e-9)
        r["row"] = i
        for k in ("cls", "rhs", "kw", "ts", "K_end", "f0", "f_end", "t_old", "y_old"):
            r.pop(k)
        rows.append(r)
    dump("dop853.json", {"single": single, "ensemble": {"B_total": 1_000_000, "seed": 0, "rtol": 1e-6, "atol": 1e-9,
                                                        "T": 10.0, "rows": rows},
                         "api": api_record(cls, {"h": 0.1})})


def events_section():
    """run_events: the reference's own event handling on its dense outputs."""
    lor = ([1.0, 1.0, 1.0], [10.0, 28.0, 8 / 3])
    cases = []
    for cls in ("RungeKutta45", "RungeKutta23", "DOP853", "RungeKutta4", "Heun3", "ForwardEuler", "AdamsBashforth2",
                "AdamsBashforth3"):
        adaptive = cls in ("RungeKutta45", "RungeKutta23", "DOP853")
        for rhs, y0, p, ts, evs, kw in (
            ("lorenz", *lor, [1.0, 2.0, 3.0, 4.0], ["y0", "z25_up"], dict(rtol=1e-6, atol=1e-9) if adaptive else dict(h=0.002)),
            ("lorenz", *lor, [1.0, 2.0, 3.0, 4.0], ["y0", "x_m8_terminal", "time"], dict(rtol=1e-6, atol=1e-9) if adaptive else dict(h=0.002)),
            ("vdp", [2.0, 0.0], [1.0], np.linspace(0.5, 20, 40), ["y0", "y1_down"], dict(rtol=1e-6, atol=1e-9) if adaptive else dict(h=0.01)),
            ("vdp", [2.0, 0.0], [1.0], np.linspace(0.5, 20, 40), ["y1_down", "y0_up_terminal"], dict(rtol=1e-6, atol=1e-9) if adaptive else dict(h=0.01)),
        ):
            s = getattr(nbkode, cls)(onp.RHS[rhs], 0.0, np.array(y0, float), params=np.array(p, float), **kw)
            rec = {"cls": cls, "rhs": rhs, "y0": L(y0), "p": L(p), "kw": kw, "ts": L(ts), "events": evs}
            try:
                t, y, te, ye = s.run_events(np.array(ts, float), [onp.EVENTS[e] for e in evs])
            except ValueError as exc:
                rec["raises"] = str(exc)
                cases.append(rec)
                continue
            # the reference appends interpolate(root) without copying: when the root is a step end point that is
            # a VIEW of the history buffer and later steps overwrite it (aliasing bug, event_handler.py:160 +
            # explicit.py:372-375).  Such entries are flagged and excluded from the parity checks.
            alias = [[bool(getattr(x, "base", None) is not None) for x in lst] for lst in ye]
            rec.update(t=L(t), y=L(y), t_events=[L(x) for x in te], y_events=[L(x) for x in ye], y_events_aliased=alias,
                       t_end=float(s.t))
            cases.append(rec)
    dump("events.json", cases)


def implicit_section():
    """Implicit multistep families: BackwardEuler, AdamsMoulton1..5, BDF1..6."""
    lor = ([1.0, 1.0, 1.0], [10.0, 28.0, 8 / 3])
    cases = []
    for cls in ("BackwardEuler", "AdamsMoulton1", "AdamsMoulton2", "AdamsMoulton3", "AdamsMoulton4", "AdamsMoulton5",
                "BDF1", "BDF2", "BDF3", "BDF4", "BDF5", "BDF6"):
        cases.append(fixed_case(cls, "vdp", [2.0, 0.0], [1.0], [1.0, 2.0, 5.0], 0.01))
        cases.append(fixed_case(cls, "vdp", [1.0, -0.5], [5.0], np.linspace(0, 2, 21), 0.0125))   # mildly stiff
        cases.append(fixed_case(cls, "lorenz", *lor, np.linspace(0, 1, 11), 0.001))
        cases.append(fixed_case(cls, "decay", [1.0], [0.1], [1.0, 2.0, 10.0], 0.1))               # n = 1: secant method
        cases.append(fixed_case(cls, "decay", [1.0, 2.0], [0.01, 0.05], [1.0, 10.0], 0.25))
    cases.append(fixed_case("BDF3", "vdp", [2.0, 0.0], [1.0], [1.0, 2.0], 0.01, first_stepper_cls=None))
    cases.append(fixed_case("AdamsMoulton5", "vdp", [2.0, 0.0], [1.0], [1.0, 2.0], 0.01, first_stepper_cls="AdamsBashforth2"))
    dump("implicit_fixed.json", cases)
    # events and the step/skip value flow with implicit solvers
    extra = {"events": [], "api": []}
    for cls in ("BackwardEuler", "AdamsMoulton3", "BDF2"):
        s = getattr(nbkode, cls)(onp.rhs_vdp, 0.0, np.array([2.0, 0.0]), params=np.array([1.0]), h=0.01)
        ts = np.linspace(0.5, 8, 16)
        evs = ["y0", "y1_down"]
        t, y, te, ye = s.run_events(ts, [onp.EVENTS[e] for e in evs])
        alias = [[bool(getattr(x, "base", None) is not None) for x in lst] for lst in ye]
        extra["events"].append({"cls": cls, "rhs": "vdp", "y0": [2.0, 0.0], "p": [1.0], "kw": {"h": 0.01}, "ts": L(ts), "events": evs,
                                "t": L(t), "y": L(y), "t_events": [L(x) for x in te], "y_events": [L(x) for x in ye],
                                "y_events_aliased": alias, "t_end": float(s.t)})
        extra["api"].append(api_record(cls, {"h": 0.1, "first_stepper_cls": None}))
    dump("implicit_extra.json", extra)


def branching_section():
    """Right-hand sides and event functions with run-time branches (`if` on the state and the time, Python's
    min / max): the reference compiles them as they are; nbkode_b200 explores their control-flow paths
    (tests/test_trace_branches.py).  Single trajectories of the live reference.

    Horizons: a kink of the right-hand side makes the error estimate of the step that crosses it depend on where
    the step starts, so an adaptive step sequence amplifies a last-bit difference by about 10 x per crossing (the
    NumPy restatement with the right-hand side scaled by 1 + 2.3e-16 leaves the golden run by 80 x the tolerance at
    t = 10, with the same number of accepted steps).  The adaptive cases therefore end at t = 1.2 (two crossings:
    the spring's stop and the damper's saturation), where one ulp moves the result by 0.005 of the tolerance; the
    fixed-step cases, which have no controller to amplify anything, run to t = 12 through some twenty crossings."""
    pw = ([1.6, -0.4], [2.0, 0.8, 0.7])
    short = np.linspace(0.1, 1.2, 12)
    adaptive = [
        adaptive_case("RungeKutta45", "pwl", *pw, short, rtol=1e-6, atol=1e-9),
        adaptive_case("RungeKutta23", "pwl", *pw, short, rtol=1e-6, atol=1e-9),
        adaptive_case("DOP853", "pwl", *pw, [0.4, 0.8, 1.2], rtol=1e-4, atol=1e-7),  # (at 1e-6 one ulp already moves it by 1.1 x the bar)
        adaptive_case("RungeKutta45", "pwl", [0.9, 0.6], [1.0, 0.5, 2.5], [0.5, 1.0], rtol=1e-5, atol=1e-8),
    ]
    fixed = [
        fixed_case("RungeKutta4", "pwl", *pw, np.linspace(0.5, 12, 24), 0.01),
        fixed_case("AdamsBashforth3", "pwl", *pw, np.linspace(0.5, 12, 24), 0.005),
        fixed_case("ForwardEuler", "pwl", *pw, [1.0, 2.0], 0.001),
    ]
    events = []
    for cls, kw, ts in (("RungeKutta45", dict(rtol=1e-6, atol=1e-9), short), ("RungeKutta23", dict(rtol=1e-6, atol=1e-9), short),
                        ("RungeKutta4", dict(h=0.01), np.linspace(0.5, 12, 24))):
        s = getattr(nbkode, cls)(onp.RHS["pwl"], 0.0, np.array(pw[0]), params=np.array(pw[1]), **kw)
        t, y, te, ye = s.run_events(ts, [onp.EVENTS["switch"], onp.EVENTS["y0"]])
        alias = [[bool(getattr(x, "base", None) is not None) for x in lst] for lst in ye]
        events.append({"cls": cls, "rhs": "pwl", "y0": L(pw[0]), "p": L(pw[1]), "kw": kw, "ts": L(ts),
                       "events": ["switch", "y0"], "t": L(t), "y": L(y), "t_events": [L(x) for x in te],
                       "y_events": [L(x) for x in ye], "y_events_aliased": alias, "t_end": float(s.t)})
        assert len(te[0]) >= 2 and len(te[1]) >= 1, [len(x) for x in te]
    dump("branching.json", {"adaptive": adaptive, "fixed": fixed, "events": events})


def main():
    if sys.argv[1:] == ["branching"]:
        return branching_section()
    branching_section()
    if sys.argv[1:] == ["implicit"]:
        return implicit_section()
    implicit_section()
    if sys.argv[1:] == ["dop853"]:
        return dop853_section()
    if sys.argv[1:] == ["events"]:
        return events_section()
    events_section()
    dop853_section()
    # ---------------------------------------------------------------- C1: README example
    c1 = []
    for cls, kw in [
        ("ForwardEuler", {}),
        ("RungeKutta45", {}),
        ("RungeKutta23", {}),
        ("AdamsBashforth2", {"h": 0.5}),
        ("AdamsBashforth4", {"h": 0.5}),
    ]:
        s = getattr(nbkode, cls)(onp.rhs_decay, 0.0, 1.0, params=np.array([0.1]), **kw)
        rec = {"cls": cls, "kw": kw, "h0": float(s.h)}
        t, y = s.run([0, 5, 10])
        rec.update(ts=L(t), ys=L(y), t_end=float(s.t), h_end=float(s.h))
        c1.append(rec)
    dump("c1_readme.json", c1)

    # ---------------------------------------------------------------- single trajectories
    single = []
    lor = ([1.0, 1.0, 1.0], [10.0, 28.0, 8 / 3])
    for cls in ("RungeKutta45", "RungeKutta23"):
        single.append(adaptive_case(cls, "lorenz", *lor, [2.5, 5.0, 10.0], rtol=1e-6, atol=1e-9))
        single.append(adaptive_case(cls, "lorenz", *lor, np.linspace(0, 3, 31), rtol=1e-3, atol=1e-6))
        single.append(adaptive_case(cls, "vdp", [2.0, 0.0], [1.0], [5.0, 10.0, 20.0], rtol=1e-6, atol=1e-9))
        single.append(adaptive_case(cls, "vdp", [2.0, 0.0], [5.0], np.linspace(0, 20, 41), rtol=1e-6, atol=1e-9))
        single.append(adaptive_case(cls, "decay", [1.0, 2.0], [0.01, 0.05], np.linspace(0, 300, 16)))
        single.append(adaptive_case(cls, "decay", [1.0], [0.01], [30.0], first_step=0.5))
        single.append(
            adaptive_case(
                cls, "vdp", [2.0, 0.0], [2.0], [10.0], rtol=1e-5, atol=1e-8,
                min_factor=0.3, max_factor=4.0, safety_factor=0.8,
            )
        )
    rng = np.random.default_rng(11)
    A8 = rng.standard_normal((8, 8)) / np.sqrt(8) - np.eye(8)
    single.append(adaptive_case("RungeKutta45", "linear", rng.standard_normal(8), A8, [1.0, 10.0], rtol=1e-6, atol=1e-9))
    y16, p16 = onp.rd1d_ensemble(2, n=16, seed=3)
    single.append(adaptive_case("RungeKutta45", "rd1d", y16[0], p16[0] / 64.0, [0.5, 1.0], rtol=1e-6, atol=1e-9))
    single.append(adaptive_case("RungeKutta23", "rd1d", y16[1], p16[1] / 64.0, [0.5, 1.0], rtol=1e-6, atol=1e-9))
    dump("adaptive_single.json", single)

    fixed = []
    vdp = ([2.0, 0.0], [1.0])
    for cls, kw in [
        ("ForwardEuler", {}),
        ("AdamsBashforth1", {}),
        ("AdamsBashforth2", {}),
        ("AdamsBashforth3", {}),
        ("AdamsBashforth4", {}),
        ("AdamsBashforth5", {}),
        ("AdamsBashforth4", {"first_stepper_cls": None}),
        ("AdamsBashforth5", {"first_stepper_cls": None}),
        ("AdamsBashforth3", {"first_stepper_cls": "AdamsBashforth1"}),
        ("AdamsBashforth5", {"first_stepper_cls": "AdamsBashforth2"}),
        ("AdamsBashforth4", {"first_stepper_cls": "RungeKutta23"}),
    ]:
        fixed.append(fixed_case(cls, "vdp", *vdp, [5.0, 10.0, 20.0], 0.01, **kw))
        fixed.append(fixed_case(cls, "vdp", [1.0, -0.5], [3.0], np.linspace(0, 4, 41), 0.0125, **kw))
    fixed.append(fixed_case("AdamsBashforth4", "lorenz", *lor, np.linspace(0, 2, 21), 0.001))
    fixed.append(fixed_case("AdamsBashforth2", "decay", [1.0, 2.0], [0.01, 0.05], [1.0, 10.0], 0.01))
    dump("fixed_single.json", fixed)

    # ---------------------------------------------------------------- fixed-step explicit RK
    erk = []
    for cls in ("Runge2", "Runge3", "Heun3", "RungeKutta4", "RungeKutta3_8"):
        erk.append(fixed_case(cls, "vdp", *vdp, [5.0, 10.0, 20.0], 0.01))
        erk.append(fixed_case(cls, "vdp", [1.0, -0.5], [3.0], np.linspace(0, 4, 41), 0.0125))
        erk.append(fixed_case(cls, "lorenz", *lor, np.linspace(0, 2, 21), 0.001))
        erk.append(fixed_case(cls, "decay", [1.0, 2.0], [0.01, 0.05], [1.0, 10.0], 0.25))
        s_ = getattr(nbkode, cls)(onp.rhs_vdp, 0.0, np.array([2.0, 0.0]), params=np.array([1.0]), h=0.05)
        ts_, ys_ = s_.step(n=10)
        erk[-1]["step10_t"], erk[-1]["step10_y"], erk[-1]["step10_K"] = L(ts_), L(ys_), L(s_.K)
    dump("erk_fixed.json", erk)

    # ---------------------------------------------------------------- ensembles
    y0, p = onp.lorenz_ensemble(1_000_000, seed=0)
    ens = {"B_total": 1_000_000, "seed": 0, "rows": [], "rtol": 1e-6, "atol": 1e-9, "T": 10.0}
    for i in list(range(48)) + list(range(999_984, 1_000_000)):
        r = adaptive_case("RungeKutta45", "lorenz", y0[i], p[i], [10.0], rtol=1e-6, atol=1e-9)
        r["row"] = i
        for k in ("cls", "rhs", "kw", "ts", "K_end", "f0", "f_end", "t_old", "y_old"):
            r.pop(k)
        ens["rows"].append(r)
    dump("lorenz_ensemble_rk45.json", ens)

    y0, p = onp.vdp_ensemble(4_000_000, seed=1)
    ts = np.linspace(0, 20, 1000)
    keep = list(range(0, 1000, 37)) + [999]
    ens = {"B_total": 4_000_000, "seed": 1, "n_ts": 1000, "keep": keep, "rk23": [], "ab4": []}
    rows = [0, 1, 2, 1_000_003, 2_000_000, 2_718_281, 3_141_592, 3_999_998, 3_999_999]
    for i in rows:
        s = nbkode.RungeKutta23(onp.rhs_vdp, 0.0, y0[i], params=p[i], rtol=1e-6, atol=1e-9)
        t, y = s.run(ts)
        s2 = nbkode.RungeKutta23(onp.rhs_vdp, 0.0, y0[i], params=p[i], rtol=1e-6, atol=1e-9)
        ens["rk23"].append(
            {"row": i, "y0": L(y0[i]), "p": L(p[i]), "ys": L(y[keep]), "t_end": float(s.t),
             "y_end": L(s.y), "n_accepted": count_to(s2, 20.0)}
        )
        s = nbkode.AdamsBashforth4(onp.rhs_vdp, 0.0, y0[i], params=p[i], h=0.01)
        t, y = s.run(ts)
        ens["ab4"].append(
            {"row": i, "y0": L(y0[i]), "p": L(p[i]), "ys": L(y[keep]), "t_end": float(s.t), "y_end": L(s.y)}
        )
    dump("vdp_ensemble.json", ens)

    # ---------------------------------------------------------------- API semantics
    # mirrors nbkode/testsuite/test_solver.py::test_f1_public_api value flow
    api = []
    for cls in ("ForwardEuler", "AdamsBashforth3", "RungeKutta23", "RungeKutta45"):
        kw = {"h": 0.1}
        if "Runge" not in cls:
            kw["first_stepper_cls"] = None
        rec = api_record(cls, kw)
        api.append(rec)
    dump("api_semantics.json", api)


if __name__ == "__main__":
    main()
