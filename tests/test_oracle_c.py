"""Pin the C oracle (oracle/nbk_oracle.c) against the reference's golden vectors and the
bit-exact NumPy restatement.  Tolerance: 1e-12 relative (scaled by the trajectory's
magnitude) for non-chaotic systems, 1e-9 for Lorenz over t<=10 (chaos amplifies ulp-level
summation-order differences ~1e4x); accepted-step counts must be identical."""

import numpy as np
import pytest

from oracle import nbk_oracle_c as oc
from oracle import nbk_oracle_np as onp

from conftest import load_golden


def _close(a, b, rel, what=""):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, what
    scale = max(1.0, float(np.max(np.abs(b))))
    err = float(np.max(np.abs(a - b))) / scale
    assert err <= rel, f"{what}: rel err {err:.3e} > {rel}"


def _tol(c):
    return 1e-12


@pytest.mark.parametrize("c", load_golden("adaptive_single.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_adaptive_single(c):
    kw = dict(c["kw"])
    shared = c["rhs"] == "linear"
    r = oc.run_ensemble(c["cls"], c["rhs"], [c["y0"]], c["p"], c["ts"], params_shared=shared, **kw)
    tol = _tol(c)
    _close(r["h0"][0], c["h0"], 1e-14, "h0")
    assert r["n_accepted"][0] == c["n_accepted"]
    _close(r["ys"][0], c["ys"], tol, "ys")
    # the step sequence itself (t_end, h_end and the state at t_end) is more sensitive than the
    # dense output at fixed times: E@K cancels 4-5 digits, so ulp-level summation-order changes
    # move h by ~1e-12 per step and that accumulates along the controller's feedback.
    _close(r["t_end"][0], c["t_end"], 1e-8, "t_end")
    _close(r["y_end"][0], c["y_end"], 1e-8, "y_end")
    _close(r["f_end"][0], c["f_end"], 1e-8, "f_end")
    _close(r["h_end"][0], c["h_end"], 1e-8, "h_end")
    assert r["status"][0] == 0 and r["n_out"][0] == len(c["ts"])


@pytest.mark.parametrize("c", load_golden("fixed_single.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{c['kw']}")
def test_fixed_single(c):
    fs = c["kw"].get("first_stepper_cls", "auto")
    r = oc.run_ensemble(c["cls"], c["rhs"], [c["y0"]], c["p"], c["ts"], h=c["h"], first_stepper=fs)
    tol = 1e-12
    _close(r["ys"][0], c["ys"], tol, "ys")
    assert r["t_end"][0] == c["t_end"]
    _close(r["y_end"][0], c["y_end"], tol)
    _close(r["f_end"][0], c["f_end"], tol)


def test_lorenz_ensemble_rows_and_threads():
    g = load_golden("lorenz_ensemble_rk45.json")
    rows = g["rows"]
    y0 = np.array([r["y0"] for r in rows])
    p = np.array([r["p"] for r in rows])
    for nt in (1, 4):
        r = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [g["T"]], rtol=g["rtol"], atol=g["atol"], nthreads=nt)
        assert np.array_equal(r["n_accepted"], [q["n_accepted"] for q in rows])
        _close(r["ys"][:, 0], [q["ys"][0] for q in rows], 1e-9)
        _close(r["t_end"], [q["t_end"] for q in rows], 1e-8)
        _close(r["h0"], [q["h0"] for q in rows], 1e-14)


def test_vdp_ensemble_rows():
    g = load_golden("vdp_ensemble.json")
    ts = np.linspace(0, 20, g["n_ts"])
    rows = g["rk23"]
    y0 = np.array([r["y0"] for r in rows])
    p = np.array([r["p"] for r in rows])
    r = oc.run_ensemble("RungeKutta23", "vdp", y0, p, ts, rtol=1e-6, atol=1e-9, nthreads=2)
    assert np.array_equal(r["n_accepted"], [q["n_accepted"] for q in rows])
    _close(r["ys"][:, g["keep"]], [q["ys"] for q in rows], 1e-10)
    r = oc.run_ensemble("AdamsBashforth4", "vdp", y0, p, ts, h=0.01, nthreads=2)
    _close(r["ys"][:, g["keep"]], [q["ys"] for q in g["ab4"]], 1e-12)
    assert np.array_equal(r["t_end"], [q["t_end"] for q in g["ab4"]])


def test_against_numpy_oracle_random_lorenz():
    y0, p = onp.lorenz_ensemble(24, seed=5)
    r = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [1.0, 4.0], rtol=1e-6, atol=1e-9)
    for i in range(len(y0)):
        s = onp.AdaptiveRK("RungeKutta45", onp.rhs_lorenz, 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
        t, y = onp.run(s, [1.0, 4.0])
        assert s.n_accepted == r["n_accepted"][i] and s.n_rejected == r["n_rejected"][i]
        _close(r["ys"][i], y, 1e-10)


def test_t_bound_prefix():
    # t_bound guard: the prefix of ts that was reached is emitted, the rest is NaN (core.py:611-617)
    r = oc.run_ensemble("AdamsBashforth2", "decay", [[1.0]], [0.1], [0.5, 1.0, 1.5, 2.0], h=0.1, t_bound=1.25, first_stepper=None)
    assert r["status"][0] == 1 and r["n_out"][0] == 2
    assert np.isnan(r["ys"][0, 2:]).all() and np.isfinite(r["ys"][0, :2]).all()


@pytest.mark.parametrize("c", load_golden("erk_fixed.json"), ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_fixed_rk(c):
    r = oc.run_ensemble(c["cls"], c["rhs"], [c["y0"]], c["p"], c["ts"], h=c["h"])
    tol = 1e-12
    _close(r["ys"][0], c["ys"], tol, "ys")
    assert r["t_end"][0] == c["t_end"]
    _close(r["y_end"][0], c["y_end"], tol)


# ---------------------------------------------------------------- DOP853
_DOP = load_golden("dop853.json")


@pytest.mark.parametrize("c", _DOP["single"], ids=lambda c: f"{c['cls']}-{c['rhs']}-{len(c['ts'])}")
def test_dop853_single(c):
    # the 8(5,3) tableau has coefficients up to ~43 that cancel, so summation-order differences between
    # BLAS and the C loops are ~20x those of RK45: 1e-10 (non-chaotic) / 1e-8 (Lorenz) on the dense output
    kw = dict(c["kw"])
    r = oc.run_ensemble(c["cls"], c["rhs"], [c["y0"]], c["p"], c["ts"], params_shared=c["rhs"] == "linear", **kw)
    _close(r["h0"][0], c["h0"], 1e-14, "h0")
    assert r["n_accepted"][0] == c["n_accepted"]
    _close(r["ys"][0], c["ys"], 1e-8 if c["rhs"] == "lorenz" else 1e-10, "ys")
    _close(r["t_end"][0], c["t_end"], 1e-8, "t_end")
    _close(r["y_end"][0], c["y_end"], 1e-7, "y_end")
    _close(r["h_end"][0], c["h_end"], 1e-6, "h_end")  # E5.K cancels ~8 digits at rtol 1e-5..1e-6
    assert r["status"][0] == 0 and r["n_out"][0] == len(c["ts"])


def test_dop853_ensemble_rows():
    g = _DOP["ensemble"]
    rows = g["rows"]
    y0 = np.array([r["y0"] for r in rows])
    p = np.array([r["p"] for r in rows])
    r = oc.run_ensemble("DOP853", "lorenz", y0, p, [g["T"]], rtol=g["rtol"], atol=g["atol"], nthreads=4)
    assert np.array_equal(r["n_accepted"], [q["n_accepted"] for q in rows])
    _close(r["ys"][:, 0], [q["ys"][0] for q in rows], 1e-8)
    _close(r["t_end"], [q["t_end"] for q in rows], 1e-8)
    _close(r["h0"], [q["h0"] for q in rows], 1e-14)
