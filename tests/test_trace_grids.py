"""Method-of-lines right-hand sides on 2-D / 3-D grids in the whole-vector tracer (nbkode_b200/rhs.py, `Grid`):
``y.reshape(N, N)``, ``np.roll(u, 1, axis=0)``, N-d slices, slice assignment into ``np.zeros_like(u)`` / ``np.zeros((N, N))``,
transposes, ``ravel()`` -- the emitted nbk_rhs_i compiled for the HOST equals NumPy, and NVRTC compiles it into the
one-CTA-per-trajectory kernels at n = 4096.  CPU only."""

import ctypes as C

import numpy as np
import pytest

from nbkode_b200 import _lib, rhs
from test_trace_branches import _host_build

N = 12


def heat2d(t, y, p):
    u = y.reshape(N, N)
    lap = np.roll(u, 1, axis=0) + np.roll(u, -1, axis=0) + np.roll(u, 1, axis=1) + np.roll(u, -1, axis=1) - 4 * u
    return (p[0] * lap + p[1] * u * (1 - u)).ravel()


def heat2d_dirichlet(t, y, p):
    u = y.reshape(N, N)
    out = np.zeros_like(u)
    out[1:-1, 1:-1] = p[0] * (u[2:, 1:-1] + u[:-2, 1:-1] + u[1:-1, 2:] + u[1:-1, :-2] - 4 * u[1:-1, 1:-1])
    out[0, :] = -u[0, :]
    out[:, -1] = np.sin(t) - u[:, -1]
    return out.ravel()


def brusselator2d(t, y, p):
    u = y[:N * N].reshape(N, N)
    v = y[N * N:].reshape((N, N))

    def lap(w):
        return np.roll(w, 1, 0) + np.roll(w, -1, 0) + np.roll(w, 1, 1) + np.roll(w, -1, 1) - 4.0 * w

    du = p[0] * lap(u) + 1.0 + u * u * v - 4.4 * u
    dv = p[0] * lap(v) + 3.4 * u - u * u * v
    return np.concatenate([du.ravel(), dv.ravel()])


def three_d_and_friends(t, y, p):
    u = y.reshape(4, 6, 6)
    w = np.where(u > 0.5, u.transpose(0, 2, 1), np.roll(u, 2, axis=2)) + u[1].sum() * 0.01
    z = np.zeros((4, 6, 6))
    z[1:3, :, 2:5] = w[0:2, :, 1:4] ** 2
    z[0] = z[0] + np.arange(36.0).reshape(6, 6) * p[0]
    if t > 0.5:
        z = z - 0.5 * u.T.T
    return np.maximum(z, -1.0).ravel() - y * np.abs(u).ravel() + (u[:, 2, :] * 0.0).sum()


CASES = [(heat2d, N * N), (heat2d_dirichlet, N * N), (brusselator2d, 2 * N * N), (three_d_and_friends, 144)]


@pytest.mark.parametrize("func,n", CASES, ids=[c[0].__name__ for c in CASES])
def test_grid_right_hand_sides_equal_numpy(tmp_path, func, n):
    src = rhs.trace_cta(func, n, (2,), True).source
    lib = _host_build(tmp_path, src, func.__name__)
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs_i.restype = C.c_double
    lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
    rng = np.random.default_rng(n)
    for _ in range(8):
        t, y, p = float(rng.uniform(0, 1)), rng.uniform(0, 1, n), rng.uniform(0.1, 1, 2)
        got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, n) for i in range(n)])
        np.testing.assert_allclose(got, func(t, y, p), rtol=1e-13, atol=1e-15)


def test_grid_right_hand_side_compiles_into_the_cta_kernels():
    """A 64 x 64 periodic reaction-diffusion grid (n = 4096): one CTA per trajectory, RungeKutta45 and AdamsBashforth3."""
    M = 64

    def rd2d(t, y, p):
        u = y.reshape(M, M)
        lap = np.roll(u, 1, axis=0) + np.roll(u, -1, axis=0) + np.roll(u, 1, axis=1) + np.roll(u, -1, axis=1) - 4 * u
        return (p[0] * lap + p[1] * u * (1 - u)).ravel()

    name, src = rhs.resolve(rd2d, M * M, (2,), True)
    assert name is None and "nbk_rhs_i" in src
    sz = C.c_longlong(0)
    for method in (45, 3):
        cfg = _lib.NbkConfig(method, M * M, 2, 0, 16, None, src.encode(), 1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)
        _lib.check(_lib.lib().nbk_jit_compile_only(C.byref(cfg), C.byref(sz)))
        assert sz.value > 10_000


def test_grid_shape_errors_are_trace_errors():
    with pytest.raises(rhs.TraceError):
        rhs.trace_cta(lambda t, y: y.reshape(5, 5), 24, (), False)
    with pytest.raises(rhs.TraceError):
        rhs.trace_cta(lambda t, y: (y.reshape(4, 6) + y.reshape(6, 4)).ravel(), 24, (), False)
    with pytest.raises(rhs.TraceError):
        rhs.trace_cta(lambda t, y: y.reshape(4, 6).sum(axis=0).repeat(4), 24, (), False)


# ------------------------------------------------------------------ random grid programs
SHAPES = [(6, 8), (8, 6), (4, 3, 4), (2, 4, 6)]
def _gen_grid_program(seed):
    import random

    r = random.Random(seed)
    shape = r.choice(SHAPES); nd = len(shape); n = int(np.prod(shape))
    L = [f"    u = y.reshape{shape}"]
    cur = "u"
    for step in range(r.randint(2, 5)):
        k = r.choice(["roll", "arith", "where", "slice_set", "transpose", "scal", "row", "pad"])
        if k == "roll":
            ax = r.randrange(nd); sh = r.randint(-3, 3)
            L.append(f"    {cur} = np.roll({cur}, {sh}, axis={ax}) + 0.5 * {cur}")
        elif k == "arith":
            L.append(f"    {cur} = {cur} * {cur} - p[0] * np.abs({cur}) + t")
        elif k == "where":
            L.append(f"    {cur} = np.where({cur} > 0.3, {cur}, -0.5 * np.roll({cur}, 1, axis={r.randrange(nd)}))")
        elif k == "slice_set":
            sl = ", ".join(r.choice(["1:-1", ":", "0:2", "1:"]) for _ in range(nd))
            L.append(f"    z = np.zeros_like({cur})")
            L.append(f"    z[{sl}] = ({cur} * 2.0)[{sl}]")
            L.append(f"    {cur} = z + 0.1 * {cur}")
        elif k == "transpose" and shape[0] == shape[-1]:
            L.append(f"    {cur} = {cur} + {cur}.T")
        elif k == "scal":
            idx = ", ".join(str(r.randrange(d)) for d in shape)
            L.append(f"    {cur} = {cur} + {cur}[{idx}] * p[1]")
        elif k == "row":
            L.append(f"    {cur} = {cur} - 0.25 * np.linspace(0, 1, {shape[-1]})")
        elif k == "pad":
            L.append(f"    w = np.zeros({shape})")
            L.append(f"    w[{', '.join(['0'] + [':'] * (nd - 1))}] = {cur}[{', '.join(['-1'] + [':'] * (nd - 1))}]")
            L.append(f"    {cur} = {cur} + w")
    if r.random() < 0.5:
        L.append(f"    if p[0] > p[1]:\n        {cur} = {cur} * 0.5\n    else:\n        {cur} = {cur} + 1.0")
    L.append(f"    return {cur}.ravel() - 0.1 * y")
    return "def f(t, y, p):\n" + "\n".join(L) + "\n", n


def test_random_grid_programs(tmp_path):
    """25 generated programs on 2-D / 3-D views (rolls along axes, N-d slice assignment, transposes, row broadcasts, single
    elements, a scalar branch): the emitted nbk_rhs_i compiled for the host equals NumPy."""
    dp = C.POINTER(C.c_double)
    for seed in range(25):
        src, n = _gen_grid_program(seed)
        ns = {"np": np}
        exec(src, ns)
        f = ns["f"]
        lib = _host_build(tmp_path, rhs.trace_cta(f, n, (2,), True).source, f"g{seed}")
        lib.nbk_rhs_i.restype = C.c_double
        lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
        rng = np.random.default_rng(seed)
        for _ in range(5):
            t, y, p = float(rng.uniform(0, 1)), rng.uniform(-1, 1, n), rng.uniform(0, 1, 2)
            got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, n) for i in range(n)])
            np.testing.assert_allclose(got, np.asarray(f(t, y, p), float), rtol=1e-12, atol=1e-13, err_msg=f"seed {seed}\n{src}")
