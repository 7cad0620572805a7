"""Right-hand sides for the GPU integrators.

The reference accepts a Python callable ``rhs(t, y[, p])`` and JIT-compiles it with numba
(nbkode/core.py:125-132).  On the GPU the right-hand side must become device code that is
inlined into the integrator kernel, so it can be given as

* the name of a built-in device function (``"lorenz"``, ``"vdp"``, ``"decay"``),
* CUDA C source (a ``str`` containing ``nbk_rhs`` or a :class:`CudaRHS`) defining::

      NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy) { ... }

  (``NBK_N`` / ``NBK_NP`` are predefined), compiled by NVRTC into the kernels, or
* a Python callable with the reference's signature: it is *traced* once with symbolic
  arguments and translated to CUDA C (no Python runs on the device, nothing falls back to
  the CPU).  Tracing supports arithmetic, ``**``, indexing/slicing, ``@`` and the NumPy
  ufuncs listed in ``_UNARY``; data-dependent Python control flow is not traceable.
"""

from __future__ import annotations

import math
import numbers

import numpy as np

BUILTINS = {
    # name -> (n or None for "from y0", n_params rule)
    "lorenz": (3, 3),
    "vdp": (2, 1),
    "decay": (None, None),   # n from y0; n_params 1 or n
    # CTA-per-trajectory regime (large systems, csrc/nbk_cta.cuh)
    "rd1d": (None, 2),       # periodic 1-D reaction-diffusion, params (d, r)
    "linear": (None, None),  # y' = A y, params = the shared n x n matrix
}
MAX_TRACED_N = 64


class CudaRHS:
    """CUDA C source of a right-hand side (see module docstring)."""

    def __init__(self, source: str, name: str = "user"):
        if "nbk_rhs" not in source:
            raise ValueError("CUDA right-hand side source must define nbk_rhs(t, y, p, dy) or nbk_rhs_i(t, y, p, i, n)")
        self.source = source
        self.name = name

    def __repr__(self):
        return f"CudaRHS({self.name!r}, {len(self.source)} chars)"


class TraceError(TypeError):
    pass


# --------------------------------------------------------------------------------------
# symbolic tracer
# --------------------------------------------------------------------------------------

_UNARY = {
    "sin": "sin", "cos": "cos", "tan": "tan", "exp": "exp", "log": "log", "sqrt": "sqrt", "tanh": "tanh",
    "sinh": "sinh", "cosh": "cosh", "arctan": "atan", "arcsin": "asin", "arccos": "acos", "log10": "log10",
    "log2": "log2", "exp2": "exp2", "expm1": "expm1", "log1p": "log1p", "fabs": "fabs", "absolute": "fabs",
}


def _lit(v) -> str:
    v = float(v)
    if math.isnan(v):
        return "nbk::d_nan()"
    if math.isinf(v):
        return "nbk::d_inf()" if v > 0 else "(-nbk::d_inf())"
    s = repr(v)
    if "e" not in s and "." not in s:
        s += ".0"
    return s


class _Trace:
    def __init__(self):
        self.stmts = []
        self.count = 0

    def new(self, code: str) -> "Expr":
        name = f"v{self.count}"
        self.count += 1
        self.stmts.append(f"    const double {name} = {code};")
        return Expr(self, name)


class Expr:
    """A scalar double value inside the traced right-hand side."""

    __slots__ = ("tr", "name")
    __array_priority__ = 1000.0

    def __init__(self, tr, name):
        self.tr, self.name = tr, name

    @staticmethod
    def _c(x):
        if isinstance(x, Expr):
            return x.name
        if isinstance(x, (numbers.Real, np.floating, np.integer, np.bool_)):
            return _lit(x)
        raise TraceError(f"cannot use {type(x).__name__} inside a traced right-hand side")

    def _bin(self, other, op, swap=False):
        if isinstance(other, np.ndarray):
            return NotImplemented
        a, b = (Expr._c(other), self.name) if swap else (self.name, Expr._c(other))
        return self.tr.new(f"{a} {op} {b}")

    def __add__(self, o): return self._bin(o, "+")
    def __radd__(self, o): return self._bin(o, "+", True)
    def __sub__(self, o): return self._bin(o, "-")
    def __rsub__(self, o): return self._bin(o, "-", True)
    def __mul__(self, o): return self._bin(o, "*")
    def __rmul__(self, o): return self._bin(o, "*", True)
    def __truediv__(self, o): return self._bin(o, "/")
    def __rtruediv__(self, o): return self._bin(o, "/", True)
    def __neg__(self): return self.tr.new(f"-{self.name}")
    def __pos__(self): return self
    def __abs__(self): return self.tr.new(f"fabs({self.name})")

    def __pow__(self, e):
        if isinstance(e, (numbers.Integral, np.integer)) or (isinstance(e, float) and e.is_integer() and abs(e) <= 8):
            k = int(e)
            if k == 0:
                return self.tr.new("1.0")
            out = self
            for _ in range(abs(k) - 1):
                out = out * self
            return out if k > 0 else 1.0 / out
        return self.tr.new(f"pow({self.name}, {Expr._c(e)})")

    def __rpow__(self, base):
        return self.tr.new(f"pow({Expr._c(base)}, {self.name})")

    def __bool__(self):
        raise TraceError("data-dependent control flow (if/while on the state) cannot be traced; "
                         "write the right-hand side as CUDA C (CudaRHS) instead")

    def __float__(self):
        raise TraceError("a traced value cannot be converted to float")

    def _cmp(self, *_):
        raise TraceError("comparisons on traced values are not supported; use CudaRHS for branching right-hand sides")

    __lt__ = __le__ = __gt__ = __ge__ = _cmp

    def __getattr__(self, item):
        # NumPy's object-dtype ufunc loops call a method named like the ufunc (np.sin -> .sin())
        if item in _UNARY:
            return lambda: self.tr.new(f"{_UNARY[item]}({self.name})")
        raise AttributeError(item)


def trace(func, n: int, params_shape, has_params: bool) -> CudaRHS:
    """Run ``func`` once on symbolic (t, y, p) and emit the equivalent ``nbk_rhs`` CUDA C."""
    tr = _Trace()
    t = Expr(tr, "t")
    y = np.empty(n, dtype=object)
    for i in range(n):
        y[i] = Expr(tr, f"y[{i}]")
    try:
        if has_params:
            p = np.empty(int(np.prod(params_shape, dtype=int)), dtype=object)
            for i in range(p.size):
                p[i] = Expr(tr, f"p[{i}]")
            out = func(t, y, p.reshape(params_shape))
        else:
            out = func(t, y)
    except TraceError:
        raise
    except Exception as exc:  # noqa: BLE001
        raise TraceError(
            f"could not trace {getattr(func, '__name__', func)!r} into CUDA C ({type(exc).__name__}: {exc}); "
            "pass a built-in name or CUDA C source (nbkode_b200.CudaRHS) instead"
        ) from exc
    out = np.asarray(out, dtype=object).ravel()
    if out.size != n:
        raise TraceError(f"right-hand side returned {out.size} values for a state of size {n}")
    lines = ["NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy) {"]
    lines += tr.stmts
    for i, e in enumerate(out):
        lines.append(f"    dy[{i}] = {Expr._c(e)};")
    lines.append("}")
    return CudaRHS("\n".join(lines) + "\n", name=getattr(func, "__name__", "traced"))


class CudaEvents:
    """CUDA C source of ``n_events`` event functions for :meth:`Solver.run_events`::

        NBK_EVENT nbk_event(int k, double t, const double* y, const double* p) { ... return g_k; }

    ``terminal`` / ``direction`` follow the SciPy convention the reference uses (nbkode/core.py:343-366)."""

    def __init__(self, source: str, n_events: int, terminal=None, direction=None):
        if "nbk_event" not in source:
            raise ValueError("CUDA event source must define nbk_event(k, t, y, p)")
        self.source, self.n_events = source, int(n_events)
        self.terminal = [bool(x) for x in (terminal if terminal is not None else [False] * self.n_events)]
        self.direction = [int(np.sign(x)) for x in (direction if direction is not None else [0] * self.n_events)]
        if len(self.terminal) != self.n_events or len(self.direction) != self.n_events:
            raise ValueError("terminal / direction must have one entry per event")


def trace_events(events, n: int, params_shape, has_params: bool) -> CudaEvents:
    """Standardise ``events`` (a callable, an iterable of callables ``event(t, y)`` -- or ``event(t, y, p)`` --
    with optional ``terminal`` / ``direction`` attributes, or a :class:`CudaEvents`) the way
    ``event_handler.build_handler`` does (nbkode/event_handler.py:236-262) and emit ``nbk_event``."""
    import inspect

    if isinstance(events, CudaEvents):
        return events
    if callable(events):
        events = (events,)
    events = list(events)
    lines = ["NBK_EVENT nbk_event(int k, double t, const double* y, const double* p) {"]
    terminal, direction = [], []
    for k, ev in enumerate(events):
        terminal.append(bool(getattr(ev, "terminal", False)))
        direction.append(int(np.sign(getattr(ev, "direction", 0))))
        tr = _Trace()
        t = Expr(tr, "t")
        y = np.empty(n, dtype=object)
        for i in range(n):
            y[i] = Expr(tr, f"y[{i}]")
        try:
            nargs = len(inspect.signature(ev).parameters)
        except (TypeError, ValueError):
            nargs = 2
        try:
            if nargs >= 3 and has_params:
                p = np.empty(int(np.prod(params_shape, dtype=int)), dtype=object)
                for i in range(p.size):
                    p[i] = Expr(tr, f"p[{i}]")
                out = ev(t, y, p.reshape(params_shape))
            else:
                out = ev(t, y)
        except TraceError:
            raise
        except Exception as exc:  # noqa: BLE001
            raise TraceError(f"could not trace event {getattr(ev, '__name__', ev)!r} into CUDA C "
                             f"({type(exc).__name__}: {exc}); pass nbkode_b200.CudaEvents instead") from exc
        out = np.asarray(out, dtype=object).ravel()
        if out.size != 1:
            raise TraceError("an event function must return one float")
        lines.append(f"  if (k == {k}) {{")
        lines += tr.stmts
        lines.append(f"    return {Expr._c(out[0])};")
        lines.append("  }")
    lines.append("  return 0.0;")
    lines.append("}")
    return CudaEvents("\n".join(lines) + "\n", len(events), terminal, direction)


def resolve(rhs, n: int, params_shape, has_params: bool):
    """-> (builtin_name or None, cuda_source or None)."""
    if isinstance(rhs, CudaRHS):
        return None, rhs.source
    if isinstance(rhs, str):
        if rhs.lower() in BUILTINS:
            return rhs.lower(), None
        if "nbk_rhs" in rhs:
            return None, rhs
        raise ValueError(f"unknown built-in right-hand side {rhs!r}; known: {sorted(BUILTINS)}")
    if callable(rhs):
        return None, trace(rhs, n, params_shape, has_params).source
    raise TypeError("rhs must be a built-in name, CUDA C source, a CudaRHS or a traceable Python callable")
