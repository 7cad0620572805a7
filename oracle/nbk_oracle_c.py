"""ctypes wrapper of the C oracle (oracle/nbk_oracle.c).  TEST INFRASTRUCTURE: only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""

from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libnbk_oracle.so")

RHS_ID = {"decay": 0, "lorenz": 1, "vdp": 2, "linear": 3, "rd1d": 4}
METHOD_ID = {
    "RungeKutta23": 23, "RungeKutta45": 45, "DOP853": 853, "ForwardEuler": 1, "AdamsBashforth1": 1,
    "AdamsBashforth2": 2, "AdamsBashforth3": 3, "AdamsBashforth4": 4, "AdamsBashforth5": 5,
    "Runge2": 202, "Runge3": 203, "Heun3": 213, "RungeKutta4": 204, "RungeKutta3_8": 238,
}


class Config(C.Structure):
    _fields_ = [
        ("method", C.c_int), ("rhs", C.c_int), ("n", C.c_int), ("np", C.c_int),
        ("params_shared", C.c_int), ("first_stepper", C.c_int), ("max_rejects", C.c_long),
        ("t0", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
        ("min_factor", C.c_double), ("max_factor", C.c_double), ("safety", C.c_double),
        ("first_step", C.c_double), ("h", C.c_double), ("t_bound", C.c_double),
    ]


_dp, _lp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_long), C.POINTER(C.c_int)


class Result(C.Structure):
    _fields_ = [
        ("ys", _dp), ("t_end", _dp), ("y_end", _dp), ("f_end", _dp), ("h0", _dp), ("h_end", _dp),
        ("n_accepted", _lp), ("n_rejected", _lp), ("n_out", _ip), ("status", _ip),
    ]


def build(force=False):
    src = os.path.join(HERE, "nbk_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_run_ensemble.restype = C.c_int
        _lib.orc_run_ensemble.argtypes = [
            C.POINTER(Config), C.c_long, _dp, _dp, _dp, C.c_int, C.POINTER(Result), C.c_int,
        ]
    return _lib


def _first_stepper_id(fs):
    if fs is None:
        return 0
    if fs == "auto":
        return 45
    if isinstance(fs, int):
        return fs
    return METHOD_ID[fs]


def run_ensemble(method, rhs, y0, params, ts, *, t0=0.0, rtol=1e-3, atol=1e-6, h=None, first_step=None,
                 first_stepper="auto", t_bound=math.inf, min_factor=0.2, max_factor=10.0,
                 safety_factor=0.9, params_shared=False, dense=True, nthreads=1, max_rejects=100000):
    """Integrate every row of y0 [B, n] with the reference's algorithm; returns a dict of
    NumPy arrays (ys [B, m, n], t_end, y_end, f_end, h0, h_end, n_accepted, n_rejected,
    n_out, status)."""
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, float)))
    B, n = y0.shape
    params = np.ascontiguousarray(np.asarray(params, float))
    if params_shared:
        npar = params.size
    else:
        params = np.ascontiguousarray(params.reshape(B, -1))
        npar = params.shape[1]
    ts = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, float)))
    m = ts.size
    cfg = Config(
        METHOD_ID[method] if isinstance(method, str) else method, RHS_ID[rhs], n, npar, int(params_shared),
        _first_stepper_id(first_stepper), max_rejects, t0, rtol, atol, min_factor, max_factor, safety_factor,
        math.nan if first_step is None else first_step, 1.0 if h is None else h, t_bound,
    )
    out = {
        "ys": np.empty((B, m, n)) if dense else None,
        "t_end": np.empty(B), "y_end": np.empty((B, n)), "f_end": np.empty((B, n)),
        "h0": np.empty(B), "h_end": np.empty(B),
        "n_accepted": np.empty(B, np.int64), "n_rejected": np.empty(B, np.int64),
        "n_out": np.empty(B, np.int32), "status": np.empty(B, np.int32),
    }

    def ptr(a, t):
        return a.ctypes.data_as(t) if a is not None else t()

    res = Result(
        ptr(out["ys"], _dp), ptr(out["t_end"], _dp), ptr(out["y_end"], _dp), ptr(out["f_end"], _dp),
        ptr(out["h0"], _dp), ptr(out["h_end"], _dp), ptr(out["n_accepted"], _lp), ptr(out["n_rejected"], _lp),
        ptr(out["n_out"], _ip), ptr(out["status"], _ip),
    )
    rc = lib().orc_run_ensemble(C.byref(cfg), B, ptr(y0, _dp), ptr(params, _dp), ptr(ts, _dp), m, C.byref(res), nthreads)
    if rc != 0:
        raise ValueError(f"orc_run_ensemble failed with {rc}")
    return out
