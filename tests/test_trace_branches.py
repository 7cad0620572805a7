"""Data-dependent control flow in Python right-hand sides and event functions (path exploration in
nbkode_b200/rhs.py).

The reference compiles the user's callable with numba (nbkode/core.py:125-132), so an ``if y[0] > 0:`` inside it is
a run-time branch.  The tracer re-runs the callable once per control-flow path and emits nested ``if`` / ``else``
(per-thread form, ``nbk_rhs``) or selects (component-wise form, ``nbk_rhs_i``).  CPU-only checks:

 * the emitted source, compiled for the HOST with g++ behind a three-line shim, returns bit for bit what the Python
   callable returns on random inputs that visit every path (the arithmetic is + - * / fabs sqrt only, so the
   comparison is exact; ``-ffp-contract=off`` keeps g++ from fusing what Python does not);
 * NVRTC compiles the same source into the sm_100a integrator kernels (no GPU needed to compile);
 * the bounds (paths, depth) and the determinism check raise ``TraceError``.
"""

import ctypes as C
import math
import subprocess
from math import exp, sin  # noqa: F401  (the numba spelling: names imported from math, used by the callables below)

import numpy as np
import pytest
from numpy import zeros

from nbkode_b200 import _lib, rhs

SHIM = r"""
#include <cmath>
namespace nbk { inline double d_nan() { return NAN; } inline double d_inf() { return INFINITY; } }
#define NBK_RHS extern "C" void
#define NBK_EVENT extern "C" double
#define NBK_RHS_I extern "C" double
#define __device__
#define __forceinline__ inline
typedef const double* nbk_vec;
"""


def _host_build(tmp_path, source, tag):
    cpp = tmp_path / f"{tag}.cpp"
    so = tmp_path / f"{tag}.so"
    cpp.write_text(SHIM + source)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)], check=True)
    return C.CDLL(str(so))


def _call_rhs(lib, t, y, p):
    dy = np.empty_like(y)
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs.restype = None
    lib.nbk_rhs.argtypes = [C.c_double, dp, dp, dp]
    pp = np.ascontiguousarray(p, dtype=np.float64).ravel() if p is not None else np.zeros(1)
    lib.nbk_rhs(float(t), y.ctypes.data_as(dp), pp.ctypes.data_as(dp), dy.ctypes.data_as(dp))
    return dy


# ------------------------------------------------------------------ right-hand sides with branches
def piecewise(t, y, p):
    # if / elif / else with `and`, Python's min / max, abs, an early return
    if y[0] > 1.0:
        a = -p[0] * y[0]
    elif y[0] < -1 and t > 2:
        a = p[0] * y[1]
    else:
        a = min(y[0], y[1])
    if t >= 4.5:
        return np.array([a, -y[1]])
    return np.array([a, max(abs(y[1]), 0.5) - t])


def friction(t, y, p):
    # Coulomb friction: the sign of the velocity picks the arm; `==` must be a run-time test as well
    v = y[1]
    if v == 0.0:
        f = 0.0
    elif v > 0:
        f = -p[1]
    else:
        f = p[1]
    return np.array([v, -p[0] * y[0] + f])


def helper_clip(x, lo, hi):
    if x < lo:
        return lo
    if x > hi:
        return hi
    return x


def with_helper_and_loop(t, y):
    # a nested helper, a chained comparison, a bounded `while` on the state, truth of a float, `not`
    a = helper_clip(y[0], -0.5, 0.5)
    b = y[1]
    k = 0
    while k < 2 and b > 1.0:
        b = b * 0.5
        k += 1
    c = 1.0 if 0.0 < y[2] < 1.0 else -1.0
    if not (y[2] * 0.0 + t):   # t == 0
        c = c * 2.0
    return np.array([a - b, b * c, (y[0] > 0) * y[2] + ((y[1] > 0) & (y[2] > 0)) * 1.5])


CASES = [
    (piecewise, 2, (1,), True),
    (friction, 2, (2,), True),
    (with_helper_and_loop, 3, (), False),
]


@pytest.mark.parametrize("func,n,pshape,has_p", CASES, ids=[c[0].__name__ for c in CASES])
def test_branching_rhs_host_compiled_source_equals_python(tmp_path, func, n, pshape, has_p):
    src = rhs.trace(func, n, pshape, has_p).source
    assert "if (" in src and "else" in src
    lib = _host_build(tmp_path, src, func.__name__)
    rng = np.random.default_rng(5)
    seen = set()
    for k in range(600):
        y = rng.uniform(-3, 3, n)
        if k % 7 == 0:
            y[rng.integers(n)] = 0.0  # the `==` arm
        t = float(rng.choice([0.0, rng.uniform(0, 5)]))
        p = rng.uniform(0.1, 2, pshape) if has_p else None
        want = np.asarray(func(t, y, p) if has_p else func(t, y), dtype=np.float64)
        got = _call_rhs(lib, t, y, p)
        assert np.array_equal(got, want), (t, y, p, got, want)
        seen.add(tuple(np.sign(want)))
    assert len(seen) >= 3  # the sample visits different arms


def test_branching_rhs_compiles_into_the_kernels():
    lib = _lib.lib()
    sz = C.c_longlong(0)

    def cfg(method, n, npar, src):
        return _lib.NbkConfig(method, n, npar, 0, 16, None, src, 1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)

    src = rhs.trace(piecewise, 2, (1,), True).source.encode()
    for method in (45, 23, 853, 4, 204):
        _lib.check(lib.nbk_jit_compile_only(C.byref(cfg(method, 2, 1, src)), C.byref(sz)))
        assert sz.value > 10_000
    src = rhs.trace(with_helper_and_loop, 3, (), False).source.encode()
    _lib.check(lib.nbk_jit_compile_only(C.byref(cfg(45, 3, 0, src)), C.byref(sz)))


def test_path_count_depth_and_determinism_are_checked():
    def many(t, y):
        return np.array([(e if e > 0 else -e) for e in y])   # 2 ** 7 element-wise paths

    assert rhs.trace(many, 6, (), False).source.count("dy[0]") == 64
    # beyond the budget: retried if-converted (test_if_conversion_beyond_the_path_budget) -- one select per element
    seven = rhs.trace(many, 7, (), False)
    assert seven.if_converted and seven.paths == 1 and seven.source.count("?") == 7
    with pytest.raises(rhs.TraceError, match="control-flow paths"):  # the same without source to rewrite
        rhs.trace(eval("lambda t, y: np.array([(e if e > 0 else -e) for e in y])", {"np": np}), 7, (), False)

    def endless(t, y):
        x = y[0]
        while x > 0:       # never settles when every comparison is taken as True
            x = x - 1.0
        return np.array([x])

    with pytest.raises(rhs.TraceError, match="decisions along one path"):
        rhs.trace(endless, 1, (), False)

    calls = []

    def flaky(t, y):
        calls.append(1)
        if len(calls) > 1 and t > 1:
            return -y
        if y[0] > 0:
            return y
        return -y

    with pytest.raises(rhs.TraceError, match="not deterministic"):
        rhs.trace(flaky, 1, (), False)

    # float() of a traced value is refused, not silently evaluated (math.* functions are wrapped while tracing, see
    # test_math_functions_and_float_containers; anything else that insists on a real float ends here)
    with pytest.raises(rhs.TraceError):
        rhs.trace(lambda t, y: np.array([float(y[0]) * 2.0]), 1, (), False)
    with pytest.raises(rhs.TraceError):
        rhs.trace(lambda t, y: np.array([math.floor(y[0])]), 1, (), False)
    with pytest.raises(TypeError):
        hash(rhs.trace.__globals__["Cond"](None, "(a < b)"))


def test_array_on_the_left_of_a_traced_scalar(tmp_path):
    """`y * p[0]`, `-y + t`, `y / np.float64(2) - y.sum()`: NumPy broadcasts the traced scalar over the object array
    whichever side it stands on (an Expr is an object scalar to NumPy, it claims no array priority)."""
    A = np.arange(9.0).reshape(3, 3) / 7.0

    def f(t, y, p):
        a = y * p[0] + (-y + t) * p[1]
        b = y / np.float64(2.0) - y.sum() + np.float64(0.5) * y[0]
        return a - b + A @ y * (t > 1) + (y > 0) * y[::-1]

    src = rhs.trace(f, 3, (2,), True).source
    lib = _host_build(tmp_path, src, "left")
    rng = np.random.default_rng(9)
    for _ in range(100):
        t, y, p = float(rng.uniform(0, 2)), rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 2)
        # (A @ y and y.sum() add in BLAS / pairwise order in NumPy, left to right in the trace: last-bit differences)
        np.testing.assert_allclose(_call_rhs(lib, t, y, p), f(t, y, p), rtol=1e-13, atol=1e-15)
    for g in (lambda t, y: y * t, lambda t, y: y - y.sum(), lambda t, y: y / np.sin(t), lambda t, y: -y + t):
        assert rhs.trace(g, 4, (), False).source.count("dy[") == 4


# ------------------------------------------------------------------ events
def test_branching_event_functions(tmp_path):
    def ev0(t, y):
        return y[0] - 1.0 if t < 3 else y[0] + 1.0

    def ev1(t, y, p):
        if p[0] > 1.0 or y[1] < 0:
            return y[1] * p[0]
        return y[1] - p[0]

    ev0.terminal = True
    ev1.direction = -1
    ev = rhs.trace_events([ev0, ev1], 2, (1,), True)
    assert ev.terminal == [True, False] and ev.direction == [0, -1] and ev.n_events == 2
    lib = _host_build(tmp_path, ev.source, "events")
    dp = C.POINTER(C.c_double)
    lib.nbk_event.restype = C.c_double
    lib.nbk_event.argtypes = [C.c_int, C.c_double, dp, dp]
    rng = np.random.default_rng(11)
    for _ in range(300):
        t, y, p = float(rng.uniform(0, 6)), rng.uniform(-2, 2, 2), rng.uniform(0.2, 2, 1)
        assert lib.nbk_event(0, t, y.ctypes.data_as(dp), p.ctypes.data_as(dp)) == ev0(t, y)
        assert lib.nbk_event(1, t, y.ctypes.data_as(dp), p.ctypes.data_as(dp)) == ev1(t, y, p)


# ------------------------------------------------------------------ component-wise form (CTA / sub-warp regimes)
def switched_diffusion(t, y, p):
    d = p[0] if t < 1.0 else 2 * p[0]
    out = d * (np.roll(y, -1) - 2 * y + np.roll(y, 1))
    if p[1] > 0 and t >= 0.5:
        out = out + p[1] * y * (1 - y)
    if y[0] > y[-1]:           # one element against another: a per-trajectory decision
        out = out - 0.25 * y
    return np.where(t < 0.1, out, np.maximum(out, -1.0))


def test_branching_vector_rhs_host_compiled_source_equals_python(tmp_path):
    n = 48
    src = rhs.trace_cta(switched_diffusion, n, (2,), True).source
    assert "nbk_rhs_i" in src and "?" in src
    lib = _host_build(tmp_path, src, "vec")
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs_i.restype = C.c_double
    lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
    rng = np.random.default_rng(3)
    for _ in range(60):
        t = float(rng.uniform(0, 1.5))
        y = rng.uniform(0, 1, n)
        p = np.array([rng.uniform(0.5, 2), rng.uniform(-1, 1)])
        want = switched_diffusion(t, y, p)
        got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, n) for i in range(n)])
        assert np.array_equal(got, want)
    # captured constant arrays are emitted once, however many paths capture them
    grid = np.linspace(0, 1, n)
    src = rhs.trace_cta(lambda t, y: (grid * y if t < 1 else (grid * y) * -1.0), n, (), False).source
    assert src.count("nbk_tab") == 3 and "nbk_tab1" not in src  # one definition + one use per path
    sz = C.c_longlong(0)
    cfg = _lib.NbkConfig(45, n, 2, 0, 16, None, rhs.trace_cta(switched_diffusion, n, (2,), True).source.encode(),
                         1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)
    _lib.check(_lib.lib().nbk_jit_compile_only(C.byref(cfg), C.byref(sz)))
    assert sz.value > 10_000


def test_resolve_prefers_selects_for_elementwise_branches():
    relu = lambda t, y, p: np.where(y > 0, -y, p[0] * y)  # noqa: E731
    # 2 ** 6 paths per thread: within the bound, but the whole-vector tracer says the same with one select
    name, src = rhs.resolve(relu, 6, (1,), True)
    assert name is None and "nbk_rhs_i" in src and src.count("?") == 1
    # a solver class without the component-wise regimes still gets the branching per-thread form
    name, src = rhs.resolve(relu, 3, (1,), True, "auto", False)
    assert "nbk_rhs(" in src and "if (" in src
    # scalar branches stay per thread for small systems
    name, src = rhs.resolve(friction, 2, (2,), True)
    assert "nbk_rhs(" in src and "if (" in src


# ------------------------------------------------------------------ the numba spelling of a right-hand side
def pendulum(t, y, p):
    dy = np.empty(2)                       # a float container, filled element by element
    dy[0] = y[1]
    dy[1] = -p[0] * sin(y[0]) - p[1] * y[1] + math.cos(t) * exp(-t / 10)
    return dy


def ring_loop(t, y, p):
    n = len(y)
    dy = zeros(n)                          # `from numpy import zeros`
    for i in range(n):
        left = y[i - 1] if i > 0 else y[n - 1]
        right = y[i + 1] if i < n - 1 else y[0]
        dy[i] = p[0] * (left - 2.0 * y[i] + right) + math.atan2(y[i], 1.0 + math.fabs(t)) + math.log(1.0 + y[i] * y[i], 2.0)
    if math.hypot(y[0], y[1]) > 1.5:       # a branch on a math.* value
        dy[0] = dy[0] - math.sqrt(y[0] * y[0] + 1.0)
    return dy


def assembled(t, y, p):
    out = np.zeros(len(y))
    out[1:-1] = p[0] * (y[2:] - 2 * y[1:-1] + y[:-2])
    out[0] = -y[0] * math.exp(-t)
    out[-1] = math.sin(t) - y[-1]
    return out + np.full(len(y), 0.25) * p[1] + np.ones(len(y))


def test_math_functions_and_float_containers(tmp_path):
    """`from math import sin`, `math.cos(t)`, `np.empty(2)` / `zeros(n)` filled element by element, `np.zeros(n)` with
    slice assembly in the whole-vector form: the way right-hand sides are written for numba.  While a callable is traced
    the math functions accept traced values and the float constructors return containers that hold them
    (rhs._tracing); afterwards every patched name is the original again -- also when tracing fails."""
    for func, n, pshape in ((pendulum, 2, (2,)), (ring_loop, 5, (1,))):
        src = rhs.trace(func, n, pshape, True).source
        lib = _host_build(tmp_path, src, func.__name__)
        rng = np.random.default_rng(4)
        for _ in range(200):
            t, y, p = float(rng.uniform(0, 3)), rng.uniform(-2, 2, n), rng.uniform(0.1, 2, pshape)
            np.testing.assert_allclose(_call_rhs(lib, t, y, p), func(t, y, p), rtol=1e-14, atol=1e-16)
    n = 40
    src = rhs.trace_cta(assembled, n, (2,), True).source
    lib = _host_build(tmp_path, src, "assembled")
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs_i.restype = C.c_double
    lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
    rng = np.random.default_rng(6)
    for _ in range(20):
        t, y, p = float(rng.uniform(0, 3)), rng.uniform(-1, 1, n), rng.uniform(0.1, 2, 2)
        got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, n) for i in range(n)])
        np.testing.assert_allclose(got, assembled(t, y, p), rtol=1e-14, atol=1e-16)
    # a constant profile filled element by element stays ONE table (not a chain of selects)
    def profiled(t, y):
        w = np.zeros(n)
        for i in range(n):
            w[i] = math.sin(i / n)
        return w * y + np.ones(n) * 0.5

    src = rhs.trace_cta(profiled, n, (), False).source
    assert src.count("nbk_tab0[") == 2 and "?" not in src and len(src) < 2500
    # an event function in the same spelling
    ev = rhs.trace_events([lambda t, y: sin(y[0]) - 0.5 * math.exp(-t)], 2, (), False)
    assert "sin(y[0])" in ev.source and "exp(" in ev.source

    def failing(t, y):
        np.zeros(2)[0] = y[0]
        raise RuntimeError("boom")

    with pytest.raises(rhs.TraceError, match="boom"):
        rhs.trace(failing, 2, (), False)
    g = globals()
    assert g["sin"] is math.sin and g["exp"] is math.exp and g["zeros"] is np.zeros
    assert math.sin(0.5) == 0.479425538604203 and np.zeros(2).dtype == np.float64 and np.empty(3).dtype == np.float64
    assert np.full(2, 1.5).dtype == np.float64 and math.log(8.0, 2.0) == 3.0 and math.atan2(1.0, 1.0) == math.pi / 4
    # another thread that calls np.zeros while a callable is traced (the shards of a devices=[...] solver do) gets floats
    import threading
    import time

    seen, stop = [], []

    def other():
        while not stop:
            seen.append(np.zeros(3).dtype)
            seen.append(np.full(2, 1.0).dtype)

    def slow(t, y):
        dy = np.empty(2)
        time.sleep(0.02)
        dy[0], dy[1] = y[1], -y[0]
        return dy

    th = threading.Thread(target=other)
    th.start()
    try:
        for _ in range(3):
            rhs.trace(slow, 2, (), False)
    finally:
        stop.append(1)
        th.join()
    assert len(seen) > 10 and set(seen) == {np.dtype(np.float64)}
    # integer / explicit dtypes and non-traced arguments pass through while tracing
    seen = {}

    def probe(t, y):
        seen["int"] = np.zeros(3, dtype=int).dtype
        seen["sin"] = math.sin(0.5)
        seen["two_d"] = np.zeros((2, 2)).shape
        return -y

    rhs.trace(probe, 2, (), False)
    assert seen == {"int": np.dtype(int), "sin": math.sin(0.5), "two_d": (2, 2)}


# ------------------------------------------------------------------ callables that numba already compiled
numba = pytest.importorskip("numba")


@numba.njit
def _grav(r2):
    return -1.0 / (r2 * math.sqrt(r2))


@numba.njit
def kepler_njit(t, y, p):
    dy = np.empty(4)
    r = np.linalg.norm(y[:2])
    g = _grav(r * r) * p[0]
    dy[0] = y[2]
    dy[1] = y[3]
    dy[2] = g * y[0]
    dy[3] = g * y[1] + 0.01 * np.arctan2(y[1], y[0])
    return dy


@numba.njit
def _crossing_njit(t, y):
    return y[1] if t < 2.0 else y[1] - 0.1


def test_numba_compiled_callables_are_traced_through_their_python_functions(tmp_path):
    """The reference accepts a right-hand side that is already an njit dispatcher (core.py:125-132).  Here the dispatcher
    -- and every njit helper of its module that it calls -- is traced through the Python function it wraps; afterwards
    the module's names are the dispatchers again."""
    src = rhs.trace(kepler_njit, 4, (1,), True).source
    assert "atan2(y[1], y[0])" in src and src.count("sqrt(") == 2
    lib = _host_build(tmp_path, src, "kepler")
    rng = np.random.default_rng(8)
    for _ in range(100):
        t, y, p = float(rng.uniform(0, 3)), rng.uniform(0.3, 2, 4), rng.uniform(0.5, 2, 1)
        np.testing.assert_allclose(_call_rhs(lib, t, y, p), kepler_njit(t, y, p), rtol=1e-14, atol=1e-16)
    ev = rhs.trace_events([_crossing_njit], 4, (1,), True)
    assert "if (t < 2.0)" in ev.source
    assert type(_grav).__name__ == "CPUDispatcher" and globals()["_grav"] is _grav and np.linalg.norm([3.0, 4.0]) == 5.0
    assert rhs.resolve(kepler_njit, 4, (1,), True)[1] == src
    # whole-vector form: binary ufuncs and the norm of the traced vector
    v = rhs.trace_cta(lambda t, y: np.arctan2(y, 1.0 + t) + np.linalg.norm(y) * np.hypot(y, np.roll(y, 1)), 32, (), False).source
    lib = _host_build(tmp_path, v, "vecnorm")
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs_i.restype = C.c_double
    lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
    y = rng.uniform(-1, 1, 32)
    p = np.zeros(1)
    got = np.array([lib.nbk_rhs_i(0.5, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, 32) for i in range(32)])
    np.testing.assert_allclose(got, np.arctan2(y, 1.5) + np.linalg.norm(y) * np.hypot(y, np.roll(y, 1)), rtol=1e-13)


# ------------------------------------------------------------------ random branching programs
def _gen_expr(r, depth, names):
    if depth == 0 or r.random() < 0.25:
        c = r.random()
        if c < 0.6: return r.choice(names)
        return repr(round(r.uniform(-2, 2), 3))
    op = r.choice(["+", "-", "*", "abs", "min", "max", "neg", "sq"])
    a = _gen_expr(r, depth-1, names); b = _gen_expr(r, depth-1, names)
    if op in "+-*": return f"({a} {op} {b})"
    if op == "abs": return f"abs({a})"
    if op == "neg": return f"(-{a})"
    if op == "sq": return f"({a} * {a})"
    return f"{op}({a}, {b})"

def _gen_cond(r, names):
    a = _gen_expr(r, 1, names); b = _gen_expr(r, 1, names)
    c = f"{a} {r.choice(['<','>','<=','>='])} {b}"
    k = r.random()
    if k < 0.2: return f"({c}) and ({_gen_expr(r,1,names)} > 0.1)"
    if k < 0.4: return f"({c}) or ({_gen_expr(r,1,names)} < -0.3)"
    if k < 0.5: return f"not ({c})"
    return c

def _gen_block(r, indent, names, nvars, budget):
    lines = []
    pad = "    " * indent
    for _ in range(r.randint(1, 3)):
        k = r.random()
        if k < 0.45 or budget[0] <= 0 or indent > 3:
            v = f"v{nvars[0]}"; nvars[0] += 1
            lines.append(f"{pad}{v} = {_gen_expr(r, 2, names)}")
            names = names + [v]
        elif k < 0.8:
            budget[0] -= 1
            tgt = r.choice([n for n in names if n.startswith('v')] or [None])
            if tgt is None:
                tgt = f"v{nvars[0]}"; nvars[0] += 1
                lines.append(f"{pad}{tgt} = 0.5"); names = names + [tgt]
            lines.append(f"{pad}if {_gen_cond(r, names)}:")
            lines.append(f"{pad}    {tgt} = {_gen_expr(r, 2, names)}")
            if r.random() < 0.6:
                lines.append(f"{pad}else:")
                lines.append(f"{pad}    {tgt} = {_gen_expr(r, 2, names)}")
        else:
            budget[0] -= 1
            lines.append(f"{pad}if {_gen_cond(r, names)}:")
            sub, _ = _gen_block(r, indent+1, names, nvars, budget)
            lines += sub
            if r.random() < 0.3:
                lines.append(f"{pad}    return np.array([{_gen_expr(r,1,names)}, {_gen_expr(r,1,names)}])")
    return lines, names

def _gen_func(seed):
    import random

    r = random.Random(seed)
    names = ["t", "y[0]", "y[1]", "p[0]"]
    nvars = [0]; budget = [4]
    lines, names = _gen_block(r, 1, names, nvars, budget)
    src = "def f(t, y, p):\n" + "\n".join(lines) + f"\n    return np.array([{_gen_expr(r,2,names)}, {_gen_expr(r,2,names)}])\n"
    return src



def test_random_branching_programs(tmp_path):
    """40 generated right-hand sides (nested if / else, early returns, and / or / not, min / max / abs, variables
    reassigned inside branches): the emitted source compiled for the host equals the Python function bit for bit on 60
    random inputs each."""
    checked = 0
    for seed in range(40):
        ns = {"np": np}
        exec(_gen_func(seed), ns)
        f = ns["f"]
        try:
            traced = rhs.trace(f, 2, (1,), True)
        except rhs.TraceError as exc:
            assert "control-flow paths" in str(exc), (seed, exc)
            continue
        lib = _host_build(tmp_path, traced.source, f"f{seed}")
        rng = np.random.default_rng(seed)
        for _ in range(60):
            t, y, p = float(rng.uniform(-1, 2)), rng.uniform(-2, 2, 2), rng.uniform(-1, 1, 1)
            assert np.array_equal(_call_rhs(lib, t, y, p), np.asarray(f(t, y, p), float)), (seed, t, y, p)
        checked += 1
    assert checked >= 30


# ------------------------------------------------------------------ random whole-vector programs
N = 16
def _vgen_vec(r, d):
    if d == 0 or r.random() < 0.2:
        return r.choice(["y", "np.roll(y, 1)", "np.roll(y, -2)", "y[::-1]", "G", "np.ones(N)", "np.concatenate([y[3:], y[:3]])"])
    k = r.choice(["+","-","*","where","max","min","abs","sc*","neg","sq","clip","diff","sin","sum*","sl"])
    a = _vgen_vec(r, d-1); b = _vgen_vec(r, d-1)
    if k in "+-*": return f"({a} {k} {b})"
    if k == "where": return f"np.where({a} {r.choice(['<','>','>='])} {_vgen_sc(r)}, {a}, {b})"
    if k == "max": return f"np.maximum({a}, {b})"
    if k == "min": return f"np.minimum({a}, {_vgen_sc(r)})"
    if k == "abs": return f"np.abs({a})"
    if k == "sc*": return f"({_vgen_sc(r)} * {a})"
    if k == "neg": return f"(-{a})"
    if k == "sq": return f"({a} ** 2)"
    if k == "clip": return f"np.clip({a}, -1.0, 1.5)"
    if k == "diff": return f"np.append(np.diff({a}), 0.0)"
    if k == "sin": return f"np.tanh({a})"
    if k == "sum*": return f"({a} * ({b}).sum() * 0.01)"
    return f"np.concatenate([({a})[:5], ({b})[5:]])"
def _vgen_sc(r):
    return r.choice(["t", "p[0]", "p[1]", "0.5", "(t * p[0])", "(p[1] - 0.3)", "y[2]", "y[-1]"])
def _vgen_func(seed):
    import random

    r = random.Random(seed)
    body = []
    body.append(f"    a = {_vgen_vec(r, 2)}")
    if r.random() < 0.6:
        body.append(f"    if {_vgen_sc(r)} {r.choice(['<','>'])} {_vgen_sc(r)}:")
        body.append(f"        a = a + {_vgen_vec(r, 1)}")
        if r.random() < 0.5:
            body.append("    else:"); body.append(f"        a = {_vgen_vec(r, 1)} - a")
    if r.random() < 0.4:
        body.append("    out = np.zeros(N)"); body.append("    out[1:-1] = a[2:] - a[:-2]"); body.append("    out[0] = a[0]"); body.append("    a = out + a")
    body.append(f"    return a + {_vgen_vec(r, 1)}")
    return "def f(t, y, p):\n" + "\n".join(body) + "\n"


def test_random_whole_vector_programs(tmp_path):
    """30 generated whole-vector right-hand sides (np.roll, reversed / shifted slices, concatenate, where / maximum /
    minimum / clip, diff, reductions, scalar branches on t / p / single elements, slice assembly into np.zeros): the
    emitted nbk_rhs_i compiled for the host equals NumPy on 25 random inputs each."""
    G = np.linspace(-1, 1, N)
    dp = C.POINTER(C.c_double)
    for seed in range(30):
        ns = {"np": np, "N": N, "G": G}
        exec(_vgen_func(seed), ns)
        f = ns["f"]
        lib = _host_build(tmp_path, rhs.trace_cta(f, N, (2,), True).source, f"v{seed}")
        lib.nbk_rhs_i.restype = C.c_double
        lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
        rng = np.random.default_rng(seed)
        for _ in range(25):
            t, y, p = float(rng.uniform(-1, 2)), rng.uniform(-2, 2, N), rng.uniform(-1, 1, 2)
            got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, N) for i in range(N)])
            np.testing.assert_allclose(got, np.asarray(f(t, y, p), float), rtol=1e-12, atol=1e-13, err_msg=f"seed {seed}")


# ------------------------------------------------------------------ if-conversion (beyond the path budget)
def relu_loop(t, y, p):
    n = len(y)
    dy = np.empty(n)
    for i in range(n):
        if y[i] > 0:
            dy[i] = -y[i]
        elif y[i] < -1 and t > 0.5:
            dy[i] = p[0] * y[i] * y[i]
        else:
            dy[i] = p[0] * y[i]
            dy[i] = dy[i] + 0.1
    s = 0.0
    for i in range(n):
        s = s + (y[i] if 0.0 < y[i] < 1.0 else 0.25 * min(y[i], 0.5))
    dy[0] = dy[0] + s
    return dy


def staged_switches(t, y, p):
    out = p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1))
    for k in range(8):                      # eight independent scalar switches: 256 paths
        if t > 0.1 * k and not y[k] < 0.0:
            out = out + 0.01 * (k + 1) * y
        else:
            out = out - p[1] * 0.001
    return out


def _many_sided_event(t, y):
    g = -2.5
    for i in range(8):
        if y[i] > 0.25:
            g = g + y[i]
        else:
            g = g + 0.25
    return g


def one_armed(t, y):
    z = 0.0 * y[0]
    for i in range(8):
        if y[i] > 0:
            half = 0.5 * y[i]               # a temporary of this arm only
            z = z - half - half
    return np.array([z] * 8)


def index_switch(t, y):
    k = 0
    for i in range(8):
        if y[i] > 0:
            k = i                           # the arms disagree on an INDEX: nothing a select can express
    return np.array([y[k]] * 8)


def side_effects(t, y):
    terms = []
    for i in range(8):
        if y[i] > 0:
            terms.append(y[i])              # a call for its side effect: running both arms would append twice
        else:
            terms.append(-y[i])
    return np.array(terms)


def test_if_conversion_beyond_the_path_budget(tmp_path):
    """n independent element-wise branches are 2 ** n paths; beyond the budget the tracers retry with the if-converted
    callable (nbkode_b200/_ifconv.py): both arms run, the assigned values merge into selects.  Bit for bit the Python
    function, per thread and in the whole-vector form; what a select cannot express reports the path budget."""
    from nbkode_b200 import _ifconv

    n = 12
    traced = rhs.trace(relu_loop, n, (1,), True)
    assert traced.if_converted and traced.paths == 1 and "if (" not in traced.source and traced.source.count("?") >= 3 * n
    lib = _host_build(tmp_path, traced.source, "relu_loop")
    rng = np.random.default_rng(21)
    for _ in range(300):
        t, y, p = float(rng.uniform(0, 1)), rng.uniform(-2, 2, n), rng.uniform(0.1, 2, 1)
        assert np.array_equal(_call_rhs(lib, t, y, p), relu_loop(t, y, p))
    sz = C.c_longlong(0)
    cfg = _lib.NbkConfig(45, n, 1, 0, 16, None, traced.source.encode(), 1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)
    _lib.check(_lib.lib().nbk_jit_compile_only(C.byref(cfg), C.byref(sz)))
    # small systems keep the exact branches (no retry below the budget)
    one = rhs.trace(relu_loop, 1, (1,), True)
    assert not getattr(one, "if_converted", False) and one.paths > 8 and "if (" in one.source

    m = 24
    v = rhs.trace_cta(staged_switches, m, (2,), True)
    assert v.if_converted
    lib = _host_build(tmp_path, v.source, "staged")
    dp = C.POINTER(C.c_double)
    lib.nbk_rhs_i.restype = C.c_double
    lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
    for _ in range(40):
        t, y, p = float(rng.uniform(0, 1)), rng.uniform(-1, 1, m), rng.uniform(0.1, 2, 2)
        got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, m) for i in range(m)])
        np.testing.assert_allclose(got, staged_switches(t, y, p), rtol=1e-13, atol=1e-15)

    # the element-by-element loop in the whole-vector form (what regime="auto" picks for n = 12): the container is merged
    # element by element, one select per element -- linear in n, not one copy of the vector per arm
    for nn in (12, 40):
        w = rhs.trace_cta(relu_loop, nn, (1,), True)
        assert w.if_converted and len(w.source) < 700 * nn
        lib = _host_build(tmp_path, w.source, f"relu_vec{nn}")
        lib.nbk_rhs_i.restype = C.c_double
        lib.nbk_rhs_i.argtypes = [C.c_double, dp, dp, C.c_int, C.c_int]
        for _ in range(20):
            t, y, p = float(rng.uniform(0, 1)), rng.uniform(-2, 2, nn), rng.uniform(0.1, 2, 1)
            got = np.array([lib.nbk_rhs_i(t, y.ctypes.data_as(dp), p.ctypes.data_as(dp), i, nn) for i in range(nn)])
            assert np.array_equal(got, relu_loop(t, y, p))
    cfg = _lib.NbkConfig(45, 12, 1, 0, 16, None, rhs.trace_cta(relu_loop, 12, (1,), True).source.encode(), 1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)
    _lib.check(_lib.lib().nbk_jit_compile_only(C.byref(cfg), C.byref(sz)))

    # event functions take the same second chance
    ev = rhs.trace_events([_many_sided_event], 8, (), False)
    lib = _host_build(tmp_path, ev.source, "ev_many")
    lib.nbk_event.restype = C.c_double
    lib.nbk_event.argtypes = [C.c_int, C.c_double, dp, dp]
    for _ in range(50):
        y = rng.uniform(-1, 1, 8)
        assert lib.nbk_event(0, 0.5, y.ctypes.data_as(dp), y.ctypes.data_as(dp)) == _many_sided_event(0.5, y)

    # a temporary bound in one arm only is that arm's business
    lib = _host_build(tmp_path, rhs.trace(one_armed, 8, (), False).source, "one_armed")
    for _ in range(50):
        y = rng.uniform(-1, 1, 8)
        assert np.array_equal(_call_rhs(lib, 0.0, y, None), one_armed(0.0, y))
    # what a select cannot express, and calls made for their side effect, stay under path exploration: the budget error
    for refused in (index_switch, side_effects):
        with pytest.raises(rhs.PathBudgetError):
            rhs.trace(refused, 8, (), False)
    with pytest.raises(rhs.PathBudgetError):  # no source to rewrite: a lambda built at run time
        rhs.trace(eval("lambda t, y: np.array([(e if e > 0 else -e) for e in y])", {"np": np}), 7, (), False)
    assert _ifconv.convert(pendulum) is None  # nothing to convert


def test_if_converted_random_programs(tmp_path):
    """The generated branching programs again, forced through the if-conversion: every program it accepts equals the Python
    function bit for bit; what it refuses (a name bound in one arm only, early returns mixed in) is refused loudly."""
    from nbkode_b200 import _ifconv

    checked = 0
    for seed in range(60):
        path = tmp_path / f"prog{seed}.py"
        path.write_text("import numpy as np\n" + _gen_func(seed))
        ns = {}
        exec(compile(path.read_text(), str(path), "exec"), ns)  # a real file, so that inspect.getsource finds it
        f = ns["f"]
        twin = _ifconv.convert(f)
        if twin is None:
            continue
        try:
            traced = rhs._trace_paths(twin, 2, (1,), True)
        except rhs.TraceError as exc:
            assert "NoConvert" in str(exc) or "control-flow paths" in str(exc), (seed, exc)
            continue
        lib = _host_build(tmp_path, traced.source, f"c{seed}")
        rng = np.random.default_rng(seed)
        for _ in range(60):
            t, y, p = float(rng.uniform(-1, 2)), rng.uniform(-2, 2, 2), rng.uniform(-1, 1, 1)
            assert np.array_equal(_call_rhs(lib, t, y, p), np.asarray(f(t, y, p), float)), (seed, t, y, p)
        checked += 1
    assert checked >= 30
