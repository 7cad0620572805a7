"""Right-hand sides the way users of the reference write them (chemistry, n-body loops with np.linalg.norm, gating
functions with removable singularities, switched parameters, np.sign friction, clamped populations, all-to-all coupling
as a broadcast and as a double loop, small matrix forms, tuple unpacking): traced into CUDA C by both tracers, the
emitted source compiled for the HOST and compared with the Python function on random inputs.  CPU only."""

import ctypes as C

import pytest

from nbkode_b200 import rhs
from test_trace_branches import _call_rhs, _host_build

import math
import numpy as np

def robertson(t, y, p):
    return np.array([-0.04*y[0] + 1e4*y[1]*y[2], 0.04*y[0] - 1e4*y[1]*y[2] - 3e7*y[1]**2, 3e7*y[1]**2])

def nbody(t, y, p):
    pos = y[:6].reshape(2, 3); vel = y[6:].reshape(2, 3)
    acc = np.zeros((2, 3))
    for i in range(2):
        for j in range(2):
            if i != j:
                d = pos[j] - pos[i]
                r = np.linalg.norm(d)
                acc[i] += p[j] * d / r**3
    return np.concatenate([vel.ravel(), acc.ravel()])

def hh_gate(t, y, p):
    v = y[0]
    if abs(v + 40.0) < 1e-6:
        am = 1.0
    else:
        am = 0.1 * (v + 40.0) / (1.0 - math.exp(-(v + 40.0) / 10.0))
    bm = 4.0 * math.exp(-(v + 65.0) / 18.0)
    m = y[1]
    return np.array([p[0] - 120.0 * m**3 * (v - 50.0) - 0.3 * (v + 54.4), am * (1.0 - m) - bm * m])

def sir(t, y, p):
    beta = p[0] if t < p[2] else p[1]
    s, i, r = y
    n = s + i + r
    return [-beta*s*i/n, beta*s*i/n - 0.1*i, 0.1*i]

def friction_sign(t, y, p):
    return np.array([y[1], -p[0]*y[0] - p[1]*np.sign(y[1])])

def lv_clamped(t, y, p):
    yy = np.maximum(y, 0.0)
    return np.array([p[0]*yy[0] - p[1]*yy[0]*yy[1], -p[2]*yy[1] + p[3]*yy[0]*yy[1]])

def kuramoto(t, y, p):
    n = len(y)
    return p[0] + (p[1] / n) * np.sin(y[None, :] - y[:, None]).sum(axis=1)

def kuramoto_loop(t, y, p):
    n = y.shape[0]
    out = np.zeros_like(y)
    for i in range(n):
        s = 0.0
        for j in range(n):
            s += math.sin(y[j] - y[i])
        out[i] = p[0] + p[1] / n * s
    return out

def matrix_form(t, y, p):
    A = np.array([[-p[0], p[1]], [p[1], -p[0]]])
    return A.dot(y) + np.array([math.cos(t), 0.0])

def with_tuple_unpack(t, y, p):
    x, v = y
    k, c = p
    return np.array((v, -k*x - c*v))

def power_and_where(t, y):
    return np.where(y > 1.0, y**-0.5, np.sqrt(np.abs(y)) + y**3)

def heaviside_drive(t, y, p):
    return -y + p[0] * (t > 1.0)


CASES = [
    (robertson, 3, (1,), "both"), (nbody, 12, (2,), "both"), (hh_gate, 2, (1,), "both"), (sir, 3, (3,), "both"),
    (friction_sign, 2, (2,), "both"), (lv_clamped, 2, (4,), "both"), (kuramoto, 5, (2,), "thread"),
    (kuramoto_loop, 5, (2,), "both"), (matrix_form, 2, (2,), "both"), (with_tuple_unpack, 2, (2,), "both"),
    (heaviside_drive, 2, (1,), "both"),
]


@pytest.mark.parametrize("func,n,pshape,forms", CASES, ids=[c[0].__name__ for c in CASES])
def test_realistic_right_hand_sides(tmp_path, func, n, pshape, forms):
    rng = np.random.default_rng(n + len(func.__name__))
    samples = []
    for _ in range(60):
        t = float(rng.uniform(0.0, 3.0))
        y = rng.uniform(0.1, 1.5, n) * rng.choice([-1.0, 1.0], n) if func not in (sir, nbody) else rng.uniform(0.2, 2.0, n)
        if func is hh_gate:
            y = np.array([rng.uniform(-80.0, 20.0), rng.uniform(0.0, 1.0)])
        p = rng.uniform(0.2, 2.0, pshape)
        samples.append((t, y, p, np.asarray(func(t, y, p), dtype=float)))
    src = rhs.trace(func, n, pshape, True).source
    lib = _host_build(tmp_path, src, func.__name__)
    for t, y, p, want in samples:
        np.testing.assert_allclose(_call_rhs(lib, t, y, p), want, rtol=1e-12, atol=1e-14)
    if forms == "thread":
        with pytest.raises(rhs.TraceError):  # (resolve() falls back to the per-thread form for these)
            rhs.trace_cta(func, n, pshape, True)
        assert "nbk_rhs(" in rhs.resolve(func, n, pshape, True)[1]
        return
    src = rhs.trace_cta(func, n, pshape, True).source
    lib = _host_build(tmp_path, src, func.__name__ + "_vec")
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs_i.restype = C.c_double
    lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
    for t, y, p, want in samples:
        pp = np.ascontiguousarray(p, dtype=np.float64).ravel()
        got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), pp.ctypes.data_as(dp), i, n) for i in range(n)])
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)


def test_scalar_on_the_left_of_a_constant_array(tmp_path):
    """`p[0] * grid`, `y[-1] * grid`, `p[0] ** grid`, `2.0 ** y` in the whole-vector form: the array side builds the vector
    whichever side the traced scalar stands on."""
    n = 8
    G = np.linspace(0.5, 1.0, n)
    forms = [
        lambda t, y, p: p[0] * G + y, lambda t, y, p: y[-1] * G - y, lambda t, y, p: (t * p[0]) * G * y,
        lambda t, y, p: p[0] / G[::-1] + y, lambda t, y, p: p[0] ** G + y, lambda t, y, p: p[0] ** p[1] + y,
        lambda t, y, p: p[0] ** y, lambda t, y, p: 2.0 ** p[0] + y, lambda t, y, p: 2.0 ** y + G ** y,
    ]
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(3)
    for k, f in enumerate(forms):
        lib = _host_build(tmp_path, rhs.trace_cta(f, n, (2,), True).source, f"left{k}")
        lib.nbk_rhs_i.restype = C.c_double
        lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
        for _ in range(10):
            t, y, p = float(rng.uniform(0.1, 2)), rng.uniform(0.2, 1.5, n), rng.uniform(0.5, 2, 2)
            got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, n) for i in range(n)])
            np.testing.assert_allclose(got, f(t, y, p), rtol=1e-13, atol=1e-15)
