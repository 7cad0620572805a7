"""Host-side tests (no GPU): registry (port of nbkode/testsuite/test_introspection.py and
test_util.py for the explicit subset), the Python->CUDA tracer, the C-ABI library's exports,
NVRTC compilation of user right-hand sides, and the loud failure without a device."""

import ctypes as C
import os
import re

import numpy as np
import pytest

import nbkode_b200 as nbkode
from nbkode_b200 import _lib, rhs
from nbkode_b200.util import CaseInsensitiveDict

from conftest import ROOT

N_SOLVERS = 26  # every solver class of the reference (nbkode/testsuite/test_introspection.py:18)


# ------------------------------------------------------------------ registry
def test_get_solvers():
    assert len(nbkode.get_solvers()) == N_SOLVERS


def test_get_solvers_groups():
    assert len(nbkode.get_solvers("Euler")) == 2
    assert len(nbkode.get_solvers("euler")) == 2
    assert len(nbkode.get_solvers("Adams-Moulton")) == 5 and len(nbkode.get_solvers("BDF")) == 6
    assert len(nbkode.get_solvers("adams-bashforth")) == 5
    assert len(nbkode.get_solvers("Runge-Kutta")) == 8
    assert nbkode.get_groups() == ("Adams-Bashforth", "Adams-Moulton", "BDF", "Euler", "Runge-Kutta")
    with pytest.raises(KeyError):
        nbkode.get_solvers("not-a-group")


def test_get_solvers_filters():
    for solver in nbkode.get_solvers(implicit=False):
        assert solver.IMPLICIT is False
    assert len(nbkode.get_solvers(implicit=True)) == 12
    for solver in nbkode.get_solvers(implicit=True):
        assert solver.IMPLICIT is True and solver.FIXED_STEP is True
    for solver in nbkode.get_solvers(fixed_step=True):
        assert solver.FIXED_STEP is True
    for solver in nbkode.get_solvers(fixed_step=False):
        assert solver.FIXED_STEP is False
    assert set(nbkode.get_solvers(fixed_step=False) + nbkode.get_solvers(fixed_step=True)) == set(nbkode.get_solvers())
    assert {nbkode.RungeKutta23, nbkode.RungeKutta45, nbkode.RungeKutta4} <= set(nbkode.get_solvers(runge_kutta=True))
    assert set(nbkode.get_solvers(runge_kutta=True, fixed_step=False)) == {nbkode.RungeKutta23, nbkode.RungeKutta45, nbkode.DOP853}
    assert len(nbkode.get_solvers(multistep=True)) == 18
    # README.rst:63 -- with the corrected IMPLICIT flag the explicit fixed-step solvers are listed
    assert nbkode.ForwardEuler in nbkode.get_solvers(implicit=False, fixed_step=True)


def test_get_solver():
    assert nbkode.get_solver("forwardeuler") is nbkode.get_solver("ForwardEuler")
    assert nbkode.get_solver("euler") is nbkode.get_solver("ForwardEuler")
    with pytest.raises(ValueError):
        nbkode.get_solver("not_a_solver")


def test_module_facade():
    assert len(dir(nbkode)) == N_SOLVERS + 4
    with pytest.warns(UserWarning):
        sol = nbkode.__getattr__("euler")
    assert sol is nbkode.get_solver("ForwardEuler")
    with pytest.raises(AttributeError):
        nbkode.not_a_solver
    assert getattr(nbkode, "ForwardEuler") is nbkode.get_solver("ForwardEuler")
    assert "Euler (alias of ForwardEuler)" in nbkode.list_solvers()


def test_class_attributes_match_reference_tableaux():
    from oracle import nbk_oracle_np as onp

    for cls, tab in ((nbkode.RungeKutta23, onp.BS23), (nbkode.RungeKutta45, onp.DP45)):
        for name in "ABCEP":
            assert np.array_equal(getattr(cls, name), getattr(tab, name)), (cls.__name__, name)
        assert cls.error_estimator_order == tab.q and cls.stages == len(tab.A) + 1
        assert cls.LEN_HISTORY == 2 and cls.FIXED_STEP is False
    for k in range(1, 6):
        cls = getattr(nbkode, f"AdamsBashforth{k}")
        assert np.array_equal(cls.B, onp.AB_B[k]) and cls.ORDER == k and cls.LEN_HISTORY == max(k, 2)
        assert cls.A[-1] == -1 and np.count_nonzero(cls.A) == 1
    for name, (A, B, Bn) in onp.IMPLICIT_TABLEAUX.items():
        cls = getattr(nbkode, name)
        assert np.array_equal(cls.A, A) and np.array_equal(cls.B, B) and cls.Bn == Bn, name
        assert cls.ORDER == A.size and cls.LEN_HISTORY == max(A.size, 2) and cls.IMPLICIT is True
    assert nbkode.ForwardEuler.GROUP == "Euler" and nbkode.ForwardEuler.ALIASES == ("Euler",)
    for name, (A, B, C) in onp.ERK_TABLEAUX.items():
        cls = getattr(nbkode, name)
        assert np.array_equal(cls.A, A) and np.array_equal(cls.B, B) and np.array_equal(cls.C, C)
        assert cls.FIXED_STEP is True and cls.stages == len(A) and cls.LEN_HISTORY == 2


# ------------------------------------------------------------------ CaseInsensitiveDict (test_util.py)
def test_case_insensitive_dict():
    cid = CaseInsensitiveDict({"Accept": "application/json"}, user="x")
    assert cid["aCCEPT"] == "application/json" and "ACCEPT" in cid and list(cid) == ["Accept", "user"]
    cid["ACCEPT"] = "text/plain"
    assert len(cid) == 2 and list(cid) == ["ACCEPT", "user"] and cid.get("accept") == "text/plain"
    assert cid == {"accept": "text/plain", "USER": "x"} and cid != {"accept": "other", "user": "x"}
    del cid["accept"]
    assert "Accept" not in cid and dict(cid.lower_items()) == {"user": "x"}
    cp = cid.copy()
    cp["z"] = 1
    assert "z" not in cid and repr(cid) == "{'user': 'x'}"
    with pytest.raises(KeyError):
        cid["missing"]


# ------------------------------------------------------------------ tracer
def test_trace_arithmetic_and_ufuncs():
    def f(t, y, p):
        return np.array([p[0] * (y[1] - y[0]) + np.sin(t), -y[0] ** 2 / p[1], abs(y[2]) ** 0.5 + 2])

    src = rhs.trace(f, 3, (2,), True).source
    assert "nbk_rhs" in src and "sin(t)" in src and "fabs(" in src and "pow(" in src and src.count("dy[") == 3

    def g(t, y):
        return -0.5 * y

    assert "dy[0]" in rhs.trace(g, 1, (), False).source

    def mat(t, y, A):
        return A @ y

    src = rhs.trace(mat, 4, (4, 4), True).source
    assert "p[15]" in src


def test_trace_rejects_untraceable():
    # comparisons are run-time branches of the emitted source (tests/test_trace_branches.py) ...
    def branching(t, y):
        return y if y[0] > 0 else -y

    src = rhs.trace(branching, 1, (), False).source
    assert "if (y[0] > 0.0) {" in src and "} else {" in src and src.count("dy[0]") == 2

    # ... and so are == / !=: they must not fall back to Python's identity comparison (that would freeze one arm of
    # the branch into the CUDA source at trace time, silently diverging from the reference, which branches at run time)
    def eq_branch(t, y):
        if y[0] == 0.0:
            return np.array([1.0 + 0 * y[0]])
        return -y

    def ne_time(t, y):
        if t != 0:
            return -y
        return y

    for fn, cond in ((eq_branch, "y[0] == 0.0"), (ne_time, "t != 0.0")):
        src = rhs.trace(fn, 1, (), False).source
        assert f"if ({cond})" in src and src.count("dy[0]") == 2
        src = rhs.trace_cta(fn, 1, (), False).source
        assert f"({cond}) ?" in src
    with pytest.raises(TypeError):
        hash(rhs.Sc("t"))
    with pytest.raises(TypeError):
        hash(rhs.Expr(None, "t"))
    # vectors: == builds a mask for np.where, like the other comparisons; the truth of a whole vector is refused
    src = rhs.trace_cta(lambda t, y: np.where(y == 0.0, 1.0, -y), 128, (), False).source
    assert "==" in src
    with pytest.raises(rhs.TraceError):
        rhs.trace_cta(lambda t, y: (y if y > 0 else -y), 128, (), False)

    def wrong_size(t, y):
        return np.array([y[0], y[0]])

    with pytest.raises(rhs.TraceError):
        rhs.trace(wrong_size, 1, (), False)
    with pytest.raises(ValueError):
        rhs.resolve("no_such_builtin", 1, (), False)
    with pytest.raises(ValueError):
        rhs.CudaRHS("void f() {}")


# ------------------------------------------------------------------ C-ABI
def _declared_functions():
    text = open(os.path.join(ROOT, "include", "nbkode_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.EXPORTS)
    assert lib.nbk_version() == 1


def _cfg(method, n, npar, rhs_name=None, src=None, **kw):
    return _lib.NbkConfig(method, n, npar, 0, 16, rhs_name, src, kw.get("rtol", 1e-6), kw.get("atol", 1e-9),
                          kw.get("min_factor", 0.2), kw.get("max_factor", 10.0), 0.9, kw.get("first_stepper", 0), 0)


def test_nvrtc_compiles_user_and_builtin_programs_for_sm100a():
    lib = _lib.lib()
    sz = C.c_longlong(0)

    def f(t, y, p):
        return np.array([y[1], p[0] * (1 - y[0] ** 2) * y[1] - y[0]])

    src = rhs.trace(f, 2, (1,), True).source.encode()
    for method in (45, 23, 853, 4, 1, 204, 11, 15, 31, 36):
        _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(method, 2, 1, src=src)), C.byref(sz)))
        assert sz.value > 10_000
    _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(45, 5, 5, rhs_name=b"decay")), C.byref(sz)))
    with pytest.raises(_lib.CompileError) as ei:
        _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(45, 1, 0, src=b"NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy) { dy[0] = undefined_symbol; }")), C.byref(sz)))
    assert "undefined_symbol" in str(ei.value)
    with pytest.raises(KeyError):
        _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(45, 3, 3, rhs_name=b"nonexistent")), C.byref(sz)))
    with pytest.raises(ValueError):
        _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(99, 3, 3, rhs_name=b"lorenz")), C.byref(sz)))
    with pytest.raises(ValueError):
        _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(45, 3, 3, rhs_name=b"lorenz", rtol=-1.0)), C.byref(sz)))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.NbkError):
        nbkode.RungeKutta45("lorenz", 0.0, [1.0, 1.0, 1.0], [10.0, 28.0, 8 / 3])
    handle = C.c_void_p()
    rc = _lib.lib().nbk_create(C.byref(_cfg(45, 3, 3, rhs_name=b"lorenz")), C.byref(handle))
    assert rc == _lib.NBK_ECUDA and not handle.value
    with pytest.raises(_lib.NbkError):
        nbkode.fp64_peak(0, 10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "numbakit-ode_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "build.py", os.path.join(dirpath, f)


def test_nvrtc_compiles_event_programs_and_traces_events():
    from oracle import nbk_oracle_np as onp

    lib = _lib.lib()
    sz = C.c_longlong(0)
    ev = rhs.trace_events([onp.ev_y0, onp.ev_z25_up, onp.ev_x_m8_terminal], 3, (3,), True)
    assert ev.n_events == 3 and ev.terminal == [False, False, True] and ev.direction == [0, 1, -1]
    assert "nbk_event" in ev.source
    for method in (45, 853, 2, 204, 13, 32):
        _lib.check(lib.nbk_jit_compile_events_only(C.byref(_cfg(method, 3, 3, rhs_name=b"lorenz")), 3, ev.source.encode(), C.byref(sz)))
        assert sz.value > 10_000
    with pytest.raises(ValueError):
        _lib.check(lib.nbk_jit_compile_events_only(C.byref(_cfg(45, 3, 3, rhs_name=b"lorenz")), 9, ev.source.encode(), C.byref(sz)))
    # component-wise right-hand sides: the adaptive classes have run_events as sub-warp teams and with one CTA per trajectory
    for method, n in ((45, 2048), (45, 24), (853, 40), (853, 300), (23, 200)):
        _lib.check(lib.nbk_jit_compile_events_only(C.byref(_cfg(method, n, 2, rhs_name=b"rd1d")), 3, ev.source.encode(), C.byref(sz)))
        assert sz.value > 10_000
    # ... also when the right-hand side is user code too (it pulls the kernel header in ahead of the events source)
    rd = lambda t, y, p: p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)  # noqa: E731
    ev2 = rhs.trace_events([lambda t, y: y[0] - 0.5, lambda t, y, p: (y[3] - p[1] if t < 1 else y[3])], 24, (2,), True)
    for method, n in ((45, 24), (853, 40), (23, 300)):
        cfg = _cfg(method, n, 2, src=rhs.trace_cta(rd, n, (2,), True).source.encode())
        _lib.check(lib.nbk_jit_compile_events_only(C.byref(cfg), 2, ev2.source.encode(), C.byref(sz)))
        assert sz.value > 10_000
    # ... and the fixed-step classes (sub-warp teams and one CTA per trajectory; built-in and user right-hand sides)
    for method, n in ((4, 200), (204, 24), (1, 2048), (2, 40), (213, 300)):
        _lib.check(lib.nbk_jit_compile_events_only(C.byref(_cfg(method, n, 2, rhs_name=b"rd1d")), 3, ev.source.encode(), C.byref(sz)))
        assert sz.value > 10_000
    cfg = _cfg(204, 24, 2, src=rhs.trace_cta(rd, 24, (2,), True).source.encode())
    _lib.check(lib.nbk_jit_compile_events_only(C.byref(cfg), 2, ev2.source.encode(), C.byref(sz)))
    # a branching event function is a run-time branch of nbk_event (tests/test_trace_branches.py)
    ev = rhs.trace_events(lambda t, y: 1.0 if y[0] > 0 else -1.0, 3, (3,), True)
    assert "if (y[0] > 0.0) {" in ev.source and "return 1.0;" in ev.source and "return -1.0;" in ev.source
    _lib.check(lib.nbk_jit_compile_events_only(C.byref(_cfg(45, 3, 3, rhs_name=b"lorenz")), 1, ev.source.encode(), C.byref(sz)))
    with pytest.raises(rhs.TraceError):
        rhs.trace_events(lambda t, y: float(y[0]), 3, (3,), True)


def test_vector_tracer_emits_componentwise_cuda_and_compiles():
    """NumPy-style whole-vector right-hand sides of large systems -> nbk_rhs_i of the CTA regime."""
    lib = _lib.lib()
    sz = C.c_longlong(0)

    def rd(t, y, p):
        d, r = p
        return d * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + r * y * (1 - y)

    src = rhs.trace_cta(rd, 2048, (2,), True).source
    assert "nbk_rhs_i" in src and "% 2048" in src
    _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(45, 2048, 2, src=src.encode())), C.byref(sz)))
    assert sz.value > 10_000

    x = np.linspace(0, 1, 100)

    def heat(t, y, p):  # Dirichlet-type boundaries assembled by slice assignment, a captured grid, a time-dependent source
        dy = np.zeros_like(y)
        dy[1:-1] = p[0] * (y[2:] - 2 * y[1:-1] + y[:-2]) + np.sin(x[1:-1]) * np.exp(-t)
        dy[0] = -y[0]
        dy[-1] = 0.0
        return dy

    src = rhs.trace_cta(heat, 100, (1,), True).source
    assert "nbk_tab0[98]" in src
    _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(23, 100, 1, src=src.encode())), C.byref(sz)))
    # resolve(): the per-thread form up to n = 16, the component-wise (CTA) form above when the solver has that
    # regime; "thread" / "cta" force one; a function the vector tracer cannot handle falls back while n <= 64
    name, s2 = rhs.resolve(rd, 2048, (2,), True)
    assert name is None and "nbk_rhs_i" in s2
    assert "nbk_rhs(" in rhs.resolve(rd, 16, (2,), True)[1]
    assert "nbk_rhs_i" in rhs.resolve(rd, 17, (2,), True)[1]
    assert "nbk_rhs(" in rhs.resolve(rd, 40, (2,), True, "auto", False)[1]      # solver without a CTA regime
    assert "nbk_rhs(" in rhs.resolve(rd, 40, (2,), True, "thread")[1]
    assert "nbk_rhs_i" in rhs.resolve(rd, 8, (2,), True, "cta")[1]
    assert "nbk_rhs(" in rhs.resolve(lambda t, y: np.cumsum(y) * 0 + y, 30, (), False)[1]  # cumsum: scalar tracer only
    # comparisons / np.where are selects in the whole-vector tracer only: a small system that uses them is traced
    # that way (CTA regime) instead of failing in the component-by-component tracer
    relu = lambda t, y, p: np.where(y > 0, -y, p[0] * y) + np.diff(np.append(y, 0.0)).sum()  # noqa: E731
    assert rhs.trace(relu, 6, (1,), True).paths == 64  # per thread: one branch per element, 2 ** 6 paths
    with pytest.raises(rhs.TraceError):
        rhs.trace(relu, 7, (1,), True)
    s3 = rhs.resolve(relu, 6, (1,), True)[1]
    assert "nbk_rhs_i" in s3 and "?" in s3
    _lib.check(lib.nbk_jit_compile_only(C.byref(_cfg(45, 6, 1, src=s3.encode())), C.byref(sz)))
    with pytest.raises(ValueError):
        rhs.resolve(rd, 100, (2,), True, "auto", False)
    with pytest.raises(rhs.TraceError):
        rhs.trace_cta(lambda t, y: np.cumsum(y), 100, (), False)
    with pytest.raises(rhs.TraceError):
        rhs.trace_cta(lambda t, y: y[:50], 100, (), False)


def test_vector_tracer_values_match_numpy(tmp_path):
    """The C expressions the vector tracer emits, compiled for the host, reproduce the NumPy right-hand side."""
    import subprocess

    x = np.linspace(0, 1, 100)

    def heat(t, y, p):
        dy = np.zeros_like(y)
        dy[1:-1] = p[0] * (y[2:] - 2 * y[1:-1] + y[:-2]) + np.sin(x[1:-1]) * np.exp(-t)
        dy[0] = -y[0]
        dy[-1] = 0.0
        return dy

    def rd(t, y, p):
        d, r = p
        return d * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + r * y * (1 - y) ** 2 / (1 + np.abs(y))

    def upwind(t, y, p):
        flux = y * y * 0.5
        return np.concatenate(([0.0], -(flux[1:] - flux[:-1]) * p[0])) + np.maximum(y, 0.25) * np.cos(t)

    rng = np.random.default_rng(3)
    A = rng.standard_normal((33, 33))

    def lin_const(t, y, p):  # a captured matrix
        return A @ y - p[0] * y

    def lin_param(t, y, p):  # the parameter block as the matrix
        return p.reshape(33, 33) @ y

    def lorenz_scalar_style(t, y, p):  # written out component by component, as for a small system
        x, yy, z = y
        return np.array([p[0] * (yy - x), x * (p[1] - z) - yy, x * yy - p[2] * z])

    w = rng.standard_normal(40)

    def limiter(t, y, p):  # comparisons / np.where / np.diff / clip / sign / reductions / reversed and strided slices
        d = np.diff(y)
        flux = np.where(d > 0, d, 0.5 * d) * (np.abs(d) <= 1.5)
        out = np.hstack(([0.0], flux)) * p[0] + np.clip(y, -0.5, 0.5) - y.mean() + (y @ y) * 1e-3
        out = out + y[::-1] * 0.1 + np.sign(y) * 0.01 + np.dot(w, y) * 1e-2 - np.sum(y * y) * p[1] + np.flip(y) * np.exp(-t)
        return out + np.append(np.diff(y, 2), [1.0, 2.0]) * 0.3 + np.concatenate((y[::2], y[1::2])) * 0.2

    for k, (fn, n, par) in enumerate(((heat, 100, [2.5]), (rd, 257, [0.7, 1.3]), (upwind, 90, [3.0]), (lin_const, 33, [0.3]),
                                      (lin_param, 33, list(rng.standard_normal(33 * 33))), (lorenz_scalar_style, 3, [10.0, 28.0, 2.6]),
                                      (limiter, 40, [0.8, 0.05]))):
        src = rhs.trace_cta(fn, n, (33, 33) if fn is lin_param else (len(par),), True).source
        cfile = tmp_path / f"t{k}.cpp"
        cfile.write_text(
            "#include <cmath>\n#define __device__\n#define __forceinline__ inline\n#define NBK_RHS_I static inline double\n"
            "struct nbk_vec { const double* b; double operator[](int i) const { return b[i]; } };\n"
            + src +
            "\nextern \"C\" void eval_all(double t, const double* y, const double* p, int n, double* out) {\n"
            "    for (int i = 0; i < n; ++i) out[i] = nbk_rhs_i(t, nbk_vec{y}, p, i, n);\n}\n")
        so = tmp_path / f"t{k}.so"
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", str(cfile), "-o", str(so)])
        lib = C.CDLL(str(so))
        y = rng.uniform(-1, 1, n)
        p = np.array(par)
        out = np.empty(n)
        dp = C.POINTER(C.c_double)
        lib.eval_all(C.c_double(0.3), y.ctypes.data_as(dp), p.ctypes.data_as(dp), n, out.ctypes.data_as(dp))
        ref = fn(0.3, y, p.reshape(33, 33) if fn is lin_param else p)
        assert np.allclose(out, ref, rtol=1e-13, atol=1e-13), (fn.__name__, np.max(np.abs(out - ref)))


def test_cubin_disk_cache(tmp_path):
    """User programs are compiled once per machine: a second process finds the cubin under NBK_CACHE_DIR."""
    import subprocess
    import sys

    code = (
        "import sys, time, ctypes as C; sys.path[:0] = [%r, %r]\n"
        "from nbkode_b200 import _lib\n"
        "src = b'NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy) { dy[0] = -p[0] * y[0] + 0.125; }'\n"
        "cfg = _lib.NbkConfig(45, 1, 1, 0, 16, None, src, 1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)\n"
        "sz = C.c_longlong(0); t0 = time.perf_counter()\n"
        "_lib.check(_lib.lib().nbk_jit_compile_only(C.byref(cfg), C.byref(sz)))\n"
        "print(sz.value, time.perf_counter() - t0)\n" % (ROOT, os.path.join(ROOT, "numbakit-ode_b200")))
    env = dict(os.environ, NBK_CACHE_DIR=str(tmp_path))
    a = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    files = list(tmp_path.glob("*.cubin"))
    assert len(files) == 1 and files[0].stat().st_size == int(a[0])
    b = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert b[0] == a[0] and float(b[1]) < 0.5 * float(a[1])  # no NVRTC run the second time
    off = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, NBK_CACHE_DIR=""), capture_output=True, text=True,
                         check=True).stdout.split()
    assert off[0] == a[0] and len(list(tmp_path.glob("*.cubin"))) == 1
