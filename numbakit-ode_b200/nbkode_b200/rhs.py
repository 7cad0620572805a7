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
  ufuncs listed in ``_UNARY``.  Data-dependent Python control flow (``if`` / ``elif`` / ``while`` /
  ``and`` / ``or`` / ``min`` / ``max`` on the state, the time or the parameters) is traced by
  *path exploration*: the callable is re-run once per control-flow path with the outcome of every
  comparison dictated by the tracer, and the paths are emitted as nested ``if`` / ``else`` (per-thread
  form) or selects (component-wise form) -- the branch is taken at run time on the device, as numba
  takes it on the CPU.  A callable with more than ``MAX_TRACED_PATHS`` paths is refused.
"""

from __future__ import annotations

import math
import numbers
import threading

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


class PathBudgetError(TraceError):
    """More control-flow paths than MAX_TRACED_PATHS: the tracers retry with the if-converted callable (_ifconv.py)."""


# --------------------------------------------------------------------------------------
# symbolic tracer
# --------------------------------------------------------------------------------------

_UNARY = {
    "sin": "sin", "cos": "cos", "tan": "tan", "exp": "exp", "log": "log", "sqrt": "sqrt", "tanh": "tanh",
    "sinh": "sinh", "cosh": "cosh", "arctan": "atan", "arcsin": "asin", "arccos": "acos", "log10": "log10",
    "log2": "log2", "exp2": "exp2", "expm1": "expm1", "log1p": "log1p", "fabs": "fabs", "absolute": "fabs",
}


_BINARY = {"arctan2": "atan2", "hypot": "hypot", "fmod": "fmod", "copysign": "copysign", "fmax": "fmax", "fmin": "fmin"}
# math.<name> -> C function (the numba-style spelling of a scalar right-hand side: `from math import sin`)
_MATH_UNARY = {"sin": "sin", "cos": "cos", "tan": "tan", "asin": "asin", "acos": "acos", "atan": "atan", "sinh": "sinh",
               "cosh": "cosh", "tanh": "tanh", "asinh": "asinh", "acosh": "acosh", "atanh": "atanh", "exp": "exp",
               "expm1": "expm1", "log10": "log10", "log2": "log2", "log1p": "log1p", "sqrt": "sqrt", "fabs": "fabs",
               "erf": "erf", "erfc": "erfc", "cbrt": "cbrt", "exp2": "exp2"}
_MATH_BINARY = {"atan2": "atan2", "hypot": "hypot", "pow": "pow", "fmod": "fmod", "copysign": "copysign"}


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


MAX_TRACED_PATHS = 64   # control-flow paths of one traced callable (each is one re-run and one emitted leaf)
MAX_TRACED_DEPTH = 256  # decisions along one path (a `while` on the state that never settles ends here)


class _Paths:
    """The decisions of ONE run of a callable under path exploration: comparisons that are asked for their truth
    value take the outcome the plan dictates, and True beyond the plan (`_explore` comes back for the False arms)."""

    def __init__(self, plan):
        self.plan = list(plan)
        self.taken = []  # (condition as C code, outcome, mark = position in the statement list)

    def decide(self, code: str, mark: int = 0) -> bool:
        k = len(self.taken)
        if k >= MAX_TRACED_DEPTH:
            raise TraceError(f"more than {MAX_TRACED_DEPTH} data-dependent decisions along one path (a loop on the state "
                             "that does not terminate under tracing?); write the right-hand side as CUDA C")
        outcome = self.plan[k] if k < len(self.plan) else True
        self.taken.append((code, outcome, mark))
        return outcome


class _Node:
    """Decision tree of a traced callable: a leaf carries the payload of its run, a branch the condition."""

    __slots__ = ("cond", "mark", "true", "false", "payload")

    def __init__(self, payload, cond=None, mark=0, true=None, false=None):
        self.payload, self.cond, self.mark, self.true, self.false = payload, cond, mark, true, false

    @property
    def is_leaf(self):
        return self.cond is None


def _explore(run, what="right-hand side"):
    """Depth-first exploration of the control-flow paths of a callable.  ``run(paths)`` executes it once under the
    decision plan of ``paths`` and returns the payload of that path; the result is the decision tree.  A run with the
    plan ``prefix`` takes True at every later decision, so it is the all-True leaf under ``prefix`` and its decisions
    name the branch nodes above that leaf; the False arm of each of them is one more run."""
    count = [0]

    def build(prefix, conds=()):
        count[0] += 1
        if count[0] > MAX_TRACED_PATHS:
            raise PathBudgetError(f"the {what} has more than {MAX_TRACED_PATHS} control-flow paths (element-wise branches on a "
                             "vector multiply them: use np.where / np.maximum on whole arrays, or CUDA C)")
        paths = _Paths(prefix)
        payload = run(paths)
        taken = paths.taken
        if [c for c, _, _ in taken[:len(prefix)]] != list(conds):  # the same questions in the same order as the parent run
            raise TraceError(f"the {what} is not deterministic under tracing (its control flow changed between runs)")
        node = _Node(payload)
        for j in range(len(taken) - 1, len(prefix) - 1, -1):
            code, _, mark = taken[j]
            other = build([o for _, o, _ in taken[:j]] + [False], [c for c, _, _ in taken[:j + 1]])
            node = _Node(payload, cond=code, mark=mark, true=node, false=other)
        return node

    return build([])


class _Trace:
    def __init__(self, paths=None):
        self.stmts = []
        self.count = 0
        self.paths = paths

    def new(self, code: str) -> "Expr":
        name = f"v{self.count}"
        self.count += 1
        self.stmts.append(f"const double {name} = {code};")
        return Expr(self, name)

    def decide(self, code: str) -> bool:
        if self.paths is None:
            raise TraceError("data-dependent control flow (if/while on the state) cannot be traced here; "
                             "write it as CUDA C instead")
        return self.paths.decide(code, len(self.stmts))


class Cond:
    """A comparison inside the traced right-hand side.  Asked for its truth value (``if``, ``and``, ``min`` ...) it
    becomes a run-time branch of the emitted code (path exploration, `_explore`); used in arithmetic it is 1.0 / 0.0
    as a Python bool would be; ``&``, ``|`` and ``~`` combine conditions without branching."""

    __slots__ = ("tr", "code")
    __hash__ = None

    def __init__(self, tr, code):
        self.tr, self.code = tr, code

    def __bool__(self):
        return self.tr.decide(self.code)

    @staticmethod
    def _code(x):
        if isinstance(x, Cond):
            return x.code
        if isinstance(x, (bool, np.bool_)):
            return "true" if x else "false"
        if isinstance(x, Expr):
            return f"({x.name} != 0.0)"
        raise TraceError(f"cannot combine a traced comparison with {type(x).__name__}")

    def __and__(self, o): return Cond(self.tr, f"({self.code} && {Cond._code(o)})")
    def __or__(self, o): return Cond(self.tr, f"({self.code} || {Cond._code(o)})")
    def __xor__(self, o): return Cond(self.tr, f"({self.code} != {Cond._code(o)})")
    __rand__, __ror__, __rxor__ = __and__, __or__, __xor__
    def __invert__(self): return Cond(self.tr, f"(!{self.code})")

    def as_expr(self) -> "Expr":
        return self.tr.new(f"({self.code} ? 1.0 : 0.0)")

    def __add__(self, o): return self.as_expr() + o
    def __radd__(self, o): return o + self.as_expr()
    def __sub__(self, o): return self.as_expr() - o
    def __rsub__(self, o): return o - self.as_expr()
    def __mul__(self, o): return self.as_expr() * o
    def __rmul__(self, o): return o * self.as_expr()
    def __neg__(self): return -self.as_expr()

    def __float__(self):
        raise TraceError("a traced comparison cannot be converted to float")


class Expr:
    """A scalar double value inside the traced right-hand side."""

    __slots__ = ("tr", "name")
    # no __array_priority__ / __array_ufunc__: to NumPy an Expr is an object scalar, so `y * p[0]` and `-y + t` (array on
    # the left) broadcast element by element through the operators below, like `p[0] * y` does

    def __init__(self, tr, name):
        self.tr, self.name = tr, name

    @staticmethod
    def _c(x):
        if isinstance(x, Expr):
            return x.name
        if isinstance(x, Cond):
            return x.as_expr().name
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
        # `if y[0]:` / `while x:` -- Python's truth of a float, decided at run time (path exploration)
        return self.tr.decide(f"({self.name} != 0.0)")

    def __float__(self):
        raise TraceError("a traced value cannot be converted to float (math.* functions need one: use the numpy ones)")

    # Comparisons are symbolic conditions (`Cond`); `==` / `!=` too: Python's default identity comparison would
    # evaluate `if y[0] == 0.0:` to a constant at trace time and freeze one arm of the branch into the CUDA source
    def _cmp(self, other, op):
        if isinstance(other, np.ndarray):
            return NotImplemented
        return Cond(self.tr, f"({self.name} {op} {Expr._c(other)})")

    def __lt__(self, o): return self._cmp(o, "<")
    def __le__(self, o): return self._cmp(o, "<=")
    def __gt__(self, o): return self._cmp(o, ">")
    def __ge__(self, o): return self._cmp(o, ">=")
    def __eq__(self, o): return self._cmp(o, "==")
    def __ne__(self, o): return self._cmp(o, "!=")
    __hash__ = None

    def __getattr__(self, item):
        # NumPy's object-dtype ufunc loops call a method named like the ufunc (np.sin -> .sin(), np.arctan2 -> .arctan2(b))
        if item in _UNARY:
            return lambda: self.tr.new(f"{_UNARY[item]}({self.name})")
        if item in _BINARY:
            return lambda other: self.tr.new(f"{_BINARY[item]}({self.name}, {Expr._c(other)})")
        if item == "sign":
            return lambda: self.tr.new(f"(({self.name}) > 0.0 ? 1.0 : (({self.name}) < 0.0 ? -1.0 : 0.0))")
        raise AttributeError(item)


def _python_function(func):
    """A callable that numba already compiled (`@numba.njit def rhs(...)`; the reference accepts those and skips its own
    njit, core.py:125-132) is traced through the Python function it wraps."""
    inner = getattr(func, "py_func", None)
    return inner if callable(inner) else func


class _tracing:
    """While a callable is traced: `math.sin(x)` & co. accept traced values (math's own functions insist on a real
    float), and `np.zeros(n)` / `np.empty(n)` / `np.ones(n)` / `np.full(n, v)` return containers that can hold them --
    the numba spelling of a right-hand side is

        from math import sin
        def rhs(t, y, p):
            dy = np.empty(2)
            dy[0] = y[1]
            dy[1] = -p[0] * sin(y[0])
            return dy

    The wrappers replace the attributes of the `math` / `numpy` modules and the names the callable's module imported
    from `math` for the duration of the trace (a few milliseconds, under a lock) and pass everything that is not a
    traced value -- and every call from another thread -- through to the originals.  ``vector``: the whole-vector tracer (containers are `Vec`s)."""

    _lock = threading.RLock()

    def __init__(self, funcs, vector=False):
        self.funcs, self.vector, self.saved = [_python_function(f) for f in funcs if f is not None], vector, []
        self.owner = None  # the tracing thread: other threads calling np.zeros & co. meanwhile get the originals

    @staticmethod
    def _traced(x):
        return isinstance(x, (Expr, Cond, Sc, Vec)) or (isinstance(x, np.ndarray) and x.dtype == object)

    def _math_wrappers(self):
        out = {}
        for name, cname in _MATH_UNARY.items():
            orig = getattr(math, name, None)
            if orig is None:
                continue

            def un(x, orig=orig, cname=cname):
                if isinstance(x, Cond):
                    x = x.as_expr()
                if isinstance(x, Expr):
                    return x.tr.new(f"{cname}({x.name})")
                if isinstance(x, Sc):
                    return Sc(f"{cname}({x.c})")
                return orig(x)

            out[name] = (orig, un)
        for name, cname in _MATH_BINARY.items():
            orig = getattr(math, name)

            def bi(a, b, orig=orig, cname=cname):
                tr = next((v.tr for v in (a, b) if isinstance(v, (Expr, Cond))), None)
                if tr is not None:
                    return tr.new(f"{cname}({Expr._c(a)}, {Expr._c(b)})")
                if isinstance(a, Sc) or isinstance(b, Sc):
                    return Sc(f"{cname}({_sc(a)}, {_sc(b)})")
                return orig(a, b)

            out[name] = (orig, bi)
        orig_log = math.log

        def log(x, *base):
            if not any(isinstance(v, (Expr, Cond, Sc)) for v in (x, *base)):
                return orig_log(x, *base)
            if isinstance(x, (Expr, Cond)):
                num = x.tr.new(f"log({Expr._c(x)})")
            elif isinstance(x, Sc):
                num = Sc(f"log({x.c})")
            else:
                num = orig_log(x)
            return num / log(base[0]) if base else num

        out["log"] = (orig_log, log)
        return out

    def _numpy_wrappers(self):
        vector = self.vector

        def container(fill):
            def make(orig):
                def f(shape, *args, **kw):
                    if threading.get_ident() != self.owner:
                        return orig(shape, *args, **kw)
                    dtype = kw.get("dtype", args[0] if args and fill != "full" else None)
                    if fill == "full":
                        value = args[0] if args else kw.get("fill_value")
                        dtype = kw.get("dtype", args[1] if len(args) > 1 else None)
                    floaty = dtype is None or dtype in (float, np.float64, "float64", "f8", "d")
                    one_d = isinstance(shape, (int, np.integer)) or (isinstance(shape, tuple) and len(shape) == 1)
                    if not floaty:
                        return orig(shape, *args, **kw)
                    if vector:
                        if not one_d:
                            if isinstance(shape, tuple) and all(isinstance(d, (int, np.integer)) for d in shape) and not (
                                    fill == "full" and isinstance(value, Sc)):
                                return Grid(_ConstVec(orig(shape, *args, **kw)), shape)
                            return orig(shape, *args, **kw)
                        length = int(shape if not isinstance(shape, tuple) else shape[0])
                        if fill == "full" and isinstance(value, Sc):
                            return Vec(length, lambda j, code=value.c: code)
                        return _ConstVec(orig(shape, *args, **kw))
                    arr = np.ndarray(shape, dtype=object)  # (np.ndarray: not one of the wrapped constructors)
                    arr[...] = value if fill == "full" else (1.0 if fill == "ones" else 0.0)
                    return arr
                return f
            return make

        return {"zeros": container("zeros"), "empty": container("empty"), "ones": container("ones"), "full": container("full")}

    def __enter__(self):
        self._lock.acquire()
        self.owner = threading.get_ident()
        try:
            wrappers = self._math_wrappers()
            originals = {id(orig): fn for orig, fn in wrappers.values()}
            for name, (orig, fn) in wrappers.items():
                self.saved.append((math, name, orig, "attr"))
                setattr(math, name, fn)
            for name, make in self._numpy_wrappers().items():
                orig = getattr(np, name)
                self.saved.append((np, name, orig, "attr"))
                setattr(np, name, make(orig))
            orig_norm = np.linalg.norm

            def norm(x, *args, **kw):  # the Euclidean norm of a traced vector (np.linalg.norm casts to float first)
                if isinstance(x, Vec) and not args and not kw:
                    return np.sqrt(_vec_sum(x * x))
                if isinstance(x, np.ndarray) and x.dtype == object and x.ndim == 1 and not args and not kw:
                    return np.sqrt((x * x).sum())
                return orig_norm(x, *args, **kw)

            self.saved.append((np.linalg, "norm", orig_norm, "attr"))
            np.linalg.norm = norm
            # `from math import sin` / `from numpy import zeros` in the callable's module: rebind those names as well;
            # helpers the callable's module decorated with numba.njit run as the Python functions they wrap
            np_originals = {id(orig): getattr(mod, name) for mod, name, orig, _ in self.saved if mod is np or mod is np.linalg}
            seen, work = set(), list(self.funcs)
            while work:
                g = getattr(work.pop(), "__globals__", None)
                if g is None or id(g) in seen or g is globals():
                    continue
                seen.add(id(g))
                for key, val in list(g.items()):
                    repl = originals.get(id(val)) or np_originals.get(id(val))
                    if repl is None and callable(val) and hasattr(val, "py_func") and callable(getattr(val, "py_func", None)):
                        repl = val.py_func
                        work.append(repl)
                    if repl is not None and callable(val):
                        self.saved.append((g, key, val, "item"))
                        g[key] = repl
        except BaseException:
            self.__exit__(None, None, None)
            raise
        return self

    def __exit__(self, *exc):
        try:
            for target, name, orig, kind in reversed(self.saved):
                if kind == "attr":
                    setattr(target, name, orig)
                else:
                    target[name] = orig
            self.saved = []
        finally:
            self._lock.release()
        return False


def _symbolic_args(tr, n, params_shape, with_params):
    t = Expr(tr, "t")
    y = np.empty(n, dtype=object)
    for i in range(n):
        y[i] = Expr(tr, f"y[{i}]")
    if not with_params:
        return (t, y)
    p = np.empty(int(np.prod(params_shape, dtype=int)), dtype=object)
    for i in range(p.size):
        p[i] = Expr(tr, f"p[{i}]")
    return (t, y, p.reshape(params_shape))


def _count_leaves(node) -> int:
    return 1 if node.is_leaf else _count_leaves(node.true) + _count_leaves(node.false)


def _emit_tree(node, leaf, lines, start=0, depth=1):
    """Statements of a decision tree: what a run computed between two decisions is emitted once, above the `if`
    whose arms continue it (all runs below a node share the statements up to the node's mark, name for name)."""
    pad = "    " * depth
    tr = node.payload[0]
    if node.is_leaf:
        lines += [pad + st for st in tr.stmts[start:]]
        lines += [pad + st for st in leaf(node.payload)]
        return
    lines += [pad + st for st in tr.stmts[start:node.mark]]
    lines.append(f"{pad}if {node.cond} {{")
    _emit_tree(node.true, leaf, lines, node.mark, depth + 1)
    lines.append(f"{pad}}} else {{")
    _emit_tree(node.false, leaf, lines, node.mark, depth + 1)
    lines.append(f"{pad}}}")


class _Explored:
    """A decision tree as the result of a tracer (what `_retry_if_converted` hands back and marks)."""

    def __init__(self, tree):
        self.tree = tree


def _retry_if_converted(tracer, func, *args):
    """``tracer(func, *args)``; when the callable has more control-flow paths than the budget, once more with its
    if-converted twin (independent branches become selects, _ifconv.py).  A failure of the retry reports the first one."""
    try:
        return tracer(func, *args)
    except PathBudgetError as first:
        from . import _ifconv

        twin = _ifconv.convert(func)
        if twin is None:
            raise
        try:
            out = tracer(twin, *args)
        except TraceError:
            raise first from None
        out.if_converted = True
        return out


def trace(func, n: int, params_shape, has_params: bool) -> CudaRHS:
    """Run ``func`` on symbolic (t, y, p) -- once per control-flow path -- and emit the equivalent ``nbk_rhs`` CUDA C."""
    return _retry_if_converted(_trace_paths, _python_function(func), n, params_shape, has_params)


def _trace_paths(func, n: int, params_shape, has_params: bool) -> CudaRHS:
    def run(paths):
        tr = _Trace(paths)
        try:
            out = func(*_symbolic_args(tr, n, params_shape, has_params))
        except TraceError:
            raise
        except Exception as exc:  # noqa: BLE001
            if isinstance(exc, TypeError) and "unorderable types" in str(exc) and paths.taken and not any(o for _, o, _ in paths.taken[-3:]):
                # np.sign on a traced scalar compares three times (< 0, > 0, == 0); the path on which all three are false
                # is the NaN input, for which NumPy's float loop returns NaN: that leaf is all-NaN
                return tr, ["nbk::d_nan()"] * n
            raise TraceError(
                f"could not trace {getattr(func, '__name__', func)!r} into CUDA C ({type(exc).__name__}: {exc}); "
                "pass a built-in name or CUDA C source (nbkode_b200.CudaRHS) instead"
            ) from exc
        out = np.asarray(out, dtype=object).ravel()
        if out.size != n:
            raise TraceError(f"right-hand side returned {out.size} values for a state of size {n}")
        return tr, [Expr._c(e) for e in out]  # (Cond outputs append their 1.0 / 0.0 statement here, before emission)

    with _tracing([func]):
        tree = _explore(run)
    lines = ["NBK_RHS nbk_rhs(double t, const double* y, const double* p, double* dy) {"]
    _emit_tree(tree, lambda payload: [f"dy[{i}] = {e};" for i, e in enumerate(payload[1])], lines)
    lines.append("}")
    out = CudaRHS("\n".join(lines) + "\n", name=getattr(func, "__name__", "traced"))
    out.paths = _count_leaves(tree)
    return out


# --------------------------------------------------------------------------------------
# vector tracer: NumPy-style right-hand sides of large systems -> component-wise nbk_rhs_i
# --------------------------------------------------------------------------------------
# A method-of-lines right-hand side is written with whole-array NumPy expressions
# (``d * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + r * y * (1 - y)``, slices, ``np.concatenate``, assignment into
# ``np.zeros_like(y)``).  Running it once on a symbolic vector yields, for every output component i, one C expression
# in i: exactly the form the CTA-per-trajectory kernels take (``nbk_rhs_i(t, y, p, i, n)``, csrc/nbk_cta.cuh).

_HANDLED = {}


def _implements(np_func):
    def deco(f):
        _HANDLED[np_func] = f
        return f
    return deco


class Vec:
    """A symbolic vector of ``length`` doubles: ``code(j)`` gives the C expression of element j (j: C int expression)."""

    __array_priority__ = 1000.0

    def __init__(self, length, code):
        self.length, self.code = int(length), code

    # ---- NumPy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        if any(isinstance(x, Grid) for x in inputs):
            return NotImplemented  # the grid's own __array_ufunc__ takes it
        name = ufunc.__name__
        binary = {"add": "+", "subtract": "-", "multiply": "*", "true_divide": "/", "divide": "/"}
        if name in binary and len(inputs) == 2:
            return _vec_binary(inputs[0], inputs[1], binary[name])
        if name == "negative":
            return _vec_unary(inputs[0], lambda a: f"(-{a})")
        if name == "positive":
            return inputs[0]
        if name in ("power", "float_power"):
            return _vec_pow(inputs[0], inputs[1])
        if name == "square":
            return _vec_pow(inputs[0], 2)
        if name in _UNARY and len(inputs) == 1:
            return _vec_unary(inputs[0], lambda a, f=_UNARY[name]: f"{f}({a})")
        if name in ("maximum", "minimum") and len(inputs) == 2:
            fn = "fmax" if name == "maximum" else "fmin"
            return _vec_binary(inputs[0], inputs[1], None, lambda a, b: f"{fn}({a}, {b})")
        if name in _BINARY and len(inputs) == 2:
            return _vec_binary(inputs[0], inputs[1], None, lambda a, b, fn=_BINARY[name]: f"{fn}({a}, {b})")
        if name == "matmul" and len(inputs) == 2:
            return _vec_matmul(inputs[0], inputs[1]) if isinstance(inputs[1], Vec) and not isinstance(inputs[0], Vec) \
                else _vec_dot(inputs[0], inputs[1])
        compare = {"less": "<", "less_equal": "<=", "greater": ">", "greater_equal": ">=", "equal": "==", "not_equal": "!="}
        if name in compare and len(inputs) == 2:
            return _vec_binary(inputs[0], inputs[1], compare[name])
        if name == "sign" and len(inputs) == 1:
            return _vec_unary(inputs[0], lambda a: f"(({a}) > 0.0 ? 1.0 : (({a}) < 0.0 ? -1.0 : 0.0))")
        if name == "reciprocal" and len(inputs) == 1:
            return _vec_unary(inputs[0], lambda a: f"(1.0 / ({a}))")
        raise TraceError(f"numpy.{name} cannot be traced into a component-wise right-hand side")

    def __array_function__(self, func, types, args, kwargs):
        if func not in _HANDLED:
            raise TraceError(f"numpy.{func.__name__} cannot be traced into a component-wise right-hand side; "
                             "write the right-hand side as CUDA C defining nbk_rhs_i(t, y, p, i, n)")
        return _HANDLED[func](*args, **kwargs)

    # ---- Python operators
    def __add__(self, o): return _vec_binary(self, o, "+")
    def __radd__(self, o): return _vec_binary(o, self, "+")
    def __sub__(self, o): return _vec_binary(self, o, "-")
    def __rsub__(self, o): return _vec_binary(o, self, "-")
    def __mul__(self, o): return _vec_binary(self, o, "*")
    def __rmul__(self, o): return _vec_binary(o, self, "*")
    def __truediv__(self, o): return _vec_binary(self, o, "/")
    def __rtruediv__(self, o): return _vec_binary(o, self, "/")
    def __neg__(self): return _vec_unary(self, lambda a: f"(-{a})")
    def __pos__(self): return self
    def __abs__(self): return _vec_unary(self, lambda a: f"fabs({a})")
    def __pow__(self, e): return _vec_pow(self, e)
    def __rpow__(self, base): return _vec_pow(base, self)
    def __rmatmul__(self, m): return _vec_matmul(m, self)
    def __matmul__(self, o): return _vec_dot(self, o)
    def __lt__(self, o): return _vec_binary(self, o, "<")
    def __le__(self, o): return _vec_binary(self, o, "<=")
    def __gt__(self, o): return _vec_binary(self, o, ">")
    def __ge__(self, o): return _vec_binary(self, o, ">=")
    def __eq__(self, o): return _vec_binary(self, o, "==")  # a mask for np.where, like the other comparisons
    def __ne__(self, o): return _vec_binary(self, o, "!=")
    __hash__ = None
    def sum(self, axis=None): return _vec_sum(self)
    def mean(self, axis=None): return _vec_sum(self) / float(self.length)
    def dot(self, o): return _vec_dot(self, o)
    def __len__(self): return self.length

    @property
    def shape(self): return (self.length,)

    @property
    def size(self): return self.length

    ndim = 1
    dtype = np.dtype(np.float64)

    def copy(self): return Vec(self.length, self.code)

    def __bool__(self):
        raise TraceError("data-dependent control flow cannot be traced; write the right-hand side as CUDA C")

    def __getitem__(self, key):
        if isinstance(key, (numbers.Integral, np.integer)):
            k = int(key) + (self.length if key < 0 else 0)
            if not 0 <= k < self.length:
                raise IndexError(key)
            return Sc(self.code(str(k)))
        if isinstance(key, slice):
            a, b, st = key.indices(self.length)
            if st != 1:
                count = len(range(a, b, st))  # any stride, e.g. y[::-1] or y[::2]
                return Vec(count, lambda j, a=a, st=st, c=self.code: c(f"({a} + ({st}) * ({j}))"))
            return Vec(max(b - a, 0), lambda j, a=a, c=self.code: c(f"({j} + {a})" if a else j))
        raise TraceError("only integer and slice indices can be traced")

    def __setitem__(self, key, value):
        """In-place assembly ``out[a:b] = expr`` / ``out[k] = scalar`` (e.g. boundary rows of a stencil)."""
        if isinstance(key, (numbers.Integral, np.integer)):
            k = int(key) + (self.length if key < 0 else 0)
            key = slice(k, k + 1)
        if not isinstance(key, slice):
            raise TraceError("only integer and slice assignments can be traced")
        a, b, st = key.indices(self.length)
        if st != 1:
            raise TraceError("only unit-stride slice assignments can be traced")
        val = _as_vec(value, b - a)
        old = self.code
        self.code = lambda j, a=a, b=b, old=old, v=val.code: (
            f"(({j}) >= {a} && ({j}) < {b} ? {v(f'(({j}) - {a})' if a else j)} : {old(j)})")


class _ConstVec(Vec):
    """What `np.zeros(n)` / `np.ones(n)` / `np.full(n, v)` return while a whole-vector callable is traced.  It stays a float
    array as long as only numbers are stored into it (a coefficient profile filled in a loop is ONE constant table);
    traced scalars stored element by element (`dy[i] = ...` in a `for` loop) are kept per element, so that the vector
    costs one expression per element however it is assembled or merged (`_ifconv.merge`); the first slice of traced
    values stored into it, or its first use in an expression, turns it into an ordinary traced vector."""

    def __init__(self, values, traced=None):
        self.values = np.array(values, dtype=np.float64).ravel()
        self.length = self.values.size
        self.traced = dict(traced or {})  # element index -> Sc
        self._code = None

    @property
    def assembling(self):
        return self._code is None

    def element(self, k):
        """The value of element k while the vector is still assembled: a float or an Sc."""
        return self.traced.get(k, float(self.values[k]))

    @property
    def code(self):
        if self._code is None:
            v = self.values
            if v.size and np.all(v == v[0]):
                base = lambda j, c=_lit(v[0]): c  # noqa: E731
            else:
                base = lambda j, name=_TABLES.add(v.copy()): f"{name}[{j}]"  # noqa: E731
            items = sorted(self.traced.items())

            def code(j, base=base, items=items):
                expr = base(j)
                for k, sc in reversed(items):
                    expr = f"(({j}) == {k} ? {sc.c} : {expr})"
                return expr

            self._code = code if items else base
        return self._code

    @code.setter
    def code(self, fn):
        self._code = fn

    def __getitem__(self, key):
        if self._code is None and isinstance(key, (numbers.Integral, np.integer)):
            k = int(key) + (self.length if key < 0 else 0)
            if not 0 <= k < self.length:
                raise IndexError(key)
            v = self.element(k)
            return v if isinstance(v, Sc) else Sc(_lit(v))
        return Vec.__getitem__(self, key)

    def __setitem__(self, key, value):
        if self._code is None and isinstance(key, (numbers.Integral, np.integer)):
            k = int(key) + (self.length if key < 0 else 0)
            if not 0 <= k < self.length:
                raise IndexError(key)
            if isinstance(value, Sc):
                self.traced[k] = value
                return
            if isinstance(value, (numbers.Real, np.floating, np.integer)):
                self.values[k] = value
                self.traced.pop(k, None)
                return
        numeric = isinstance(value, (numbers.Real, np.floating, np.integer)) or (isinstance(value, np.ndarray) and value.dtype != object)
        if self._code is None and numeric and isinstance(key, slice):
            self.values[key] = value
            for k in range(*key.indices(self.length)):
                self.traced.pop(k, None)
            return
        Vec.__setitem__(self, key, value)

    def copy(self):
        return _ConstVec(self.values, self.traced) if self._code is None else Vec(self.length, self._code)


class Sc:
    """A symbolic scalar (time, a parameter, one vector element, or arithmetic on them)."""

    __array_priority__ = 1001.0

    def __init__(self, code, boolean=False):
        self.c, self.boolean = code, boolean

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if any(isinstance(x, Grid) for x in inputs):
            return NotImplemented  # the grid's own __array_ufunc__ takes it
        nd = next((x for x in inputs if isinstance(x, np.ndarray) and x.ndim >= 2), None)
        if nd is not None:  # scalar (x) constant N-d array: a grid of constants
            const = Grid(_as_vec(np.ascontiguousarray(nd, dtype=np.float64).ravel(), nd.size), nd.shape)
            return const.__array_ufunc__(ufunc, method, *[const if x is nd else x for x in inputs], **kwargs)
        if any(isinstance(x, Vec) or (isinstance(x, np.ndarray) and x.ndim == 1) for x in inputs):
            return Vec.__array_ufunc__(None, ufunc, method, *inputs, **kwargs)  # vector (x) scalar
        if method != "__call__":
            return NotImplemented
        name = ufunc.__name__
        binary = {"add": "+", "subtract": "-", "multiply": "*", "true_divide": "/", "divide": "/"}
        if name in binary and len(inputs) == 2:
            return Sc(f"({_sc(inputs[0])} {binary[name]} {_sc(inputs[1])})")
        if name == "negative":
            return Sc(f"(-{_sc(inputs[0])})")
        if name in _UNARY and len(inputs) == 1:
            return Sc(f"{_UNARY[name]}({_sc(inputs[0])})")
        if name in _BINARY and len(inputs) == 2:
            return Sc(f"{_BINARY[name]}({_sc(inputs[0])}, {_sc(inputs[1])})")
        if name in ("maximum", "minimum") and len(inputs) == 2:
            return Sc(f"{'fmax' if name == 'maximum' else 'fmin'}({_sc(inputs[0])}, {_sc(inputs[1])})")
        if name == "sign" and len(inputs) == 1:
            a = _sc(inputs[0])
            return Sc(f"(({a}) > 0.0 ? 1.0 : (({a}) < 0.0 ? -1.0 : 0.0))")
        if name == "square" and len(inputs) == 1:
            return Sc(_pow_code(_sc(inputs[0]), 2))
        if name == "reciprocal" and len(inputs) == 1:
            return Sc(f"(1.0 / ({_sc(inputs[0])}))")
        if name in ("power", "float_power"):
            return Sc(_pow_code(_sc(inputs[0]), inputs[1]))
        raise TraceError(f"numpy.{name} cannot be traced")

    def _b(self, o, op, swap=False):
        if isinstance(o, (Vec, Grid)) or (isinstance(o, np.ndarray) and o.ndim > 0):
            return NotImplemented  # scalar (x) vector / constant array: the array side builds the vector (`p[0] * grid`)
        a, b = (_sc(o), self.c) if swap else (self.c, _sc(o))
        return Sc(f"({a} {op} {b})")

    def __add__(self, o): return self._b(o, "+")
    def __radd__(self, o): return self._b(o, "+", True)
    def __sub__(self, o): return self._b(o, "-")
    def __rsub__(self, o): return self._b(o, "-", True)
    def __mul__(self, o): return self._b(o, "*")
    def __rmul__(self, o): return self._b(o, "*", True)
    def __truediv__(self, o): return self._b(o, "/")
    def __rtruediv__(self, o): return self._b(o, "/", True)
    def __neg__(self): return Sc(f"(-{self.c})")
    def __pos__(self): return self
    def __abs__(self): return Sc(f"fabs({self.c})")
    def __pow__(self, e):
        if isinstance(e, (Vec, Grid)) or (isinstance(e, np.ndarray) and e.ndim > 0):
            return NotImplemented  # scalar ** vector: the array side builds the vector
        return Sc(_pow_code(self.c, e))

    def __rpow__(self, base): return Sc(f"pow({_sc(base)}, {self.c})")

    # Comparisons of scalars (time, parameters, single elements) are boolean scalars: a mask for np.where, 1 / 0 in
    # arithmetic, and -- asked for their truth value by `if` / `and` / `min` -- a run-time select between the
    # control-flow paths of the callable (path exploration, `_explore`).  `==` / `!=` too: Python's identity
    # comparison would freeze an `if t == 0:` at trace time.
    def _cmp(self, o, op):
        if isinstance(o, (Vec, Grid, np.ndarray)):
            return NotImplemented  # scalar (x) vector comparison: the vector builds the mask
        return Sc(f"({self.c} {op} {_sc(o)})", boolean=True)

    def __lt__(self, o): return self._cmp(o, "<")
    def __le__(self, o): return self._cmp(o, "<=")
    def __gt__(self, o): return self._cmp(o, ">")
    def __ge__(self, o): return self._cmp(o, ">=")
    def __eq__(self, o): return self._cmp(o, "==")
    def __ne__(self, o): return self._cmp(o, "!=")
    __hash__ = None

    def _logic(self, o, op):
        if isinstance(o, (Vec, Grid, np.ndarray)):
            return NotImplemented
        return Sc(f"({self._truth()} {op} {o._truth() if isinstance(o, Sc) else ('true' if o else 'false')})", boolean=True)

    def __and__(self, o): return self._logic(o, "&&")
    def __or__(self, o): return self._logic(o, "||")
    __rand__, __ror__ = __and__, __or__
    def __invert__(self): return Sc(f"(!{self._truth()})", boolean=True)

    def _truth(self) -> str:
        return self.c if self.boolean else f"({self.c} != 0.0)"

    def __bool__(self):
        if _SC_PATHS is None:
            raise TraceError("data-dependent control flow cannot be traced here; write the right-hand side as CUDA C")
        return _SC_PATHS.decide(self._truth())

    def __float__(self):
        raise TraceError("a traced value cannot be converted to float (math.* functions need one: use the numpy ones)")


_SC_PATHS: "_Paths | None" = None


def _sc(x) -> str:
    if isinstance(x, Sc):
        return x.c
    if isinstance(x, (numbers.Real, np.floating, np.integer, np.bool_)):
        return _lit(x)
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return _lit(x.item())
    raise TraceError(f"cannot use {type(x).__name__} as a scalar inside a traced right-hand side")


def _pow_code(a: str, e) -> str:
    if isinstance(e, (numbers.Integral, np.integer)) or (isinstance(e, float) and e.is_integer() and abs(e) <= 8):
        k = int(e)
        if k == 0:
            return "1.0"
        body = " * ".join([f"({a})"] * abs(k))
        return f"({body})" if k > 0 else f"(1.0 / ({body}))"
    return f"pow({a}, {_sc(e)})"


class _Tables:
    """Constant NumPy vectors captured by the right-hand side (grids, coefficient profiles): emitted as device arrays."""

    def __init__(self):
        self.arrays = []
        self.functions = []
        self._function_formats = []

    def add(self, arr) -> str:
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        for k, a in enumerate(self.arrays):  # the runs of a path exploration capture the same constants again
            if a.shape == arr.shape and a.tobytes() == arr.tobytes():
                return f"nbk_tab{k}"
        self.arrays.append(arr)
        return f"nbk_tab{len(self.arrays) - 1}"

    def source(self) -> str:
        out = []
        for k, a in enumerate(self.arrays):
            out.append(f"__device__ const double nbk_tab{k}[{a.size}] = {{" + ", ".join(_lit(v) for v in a.ravel()) + "};")
        return "\n".join(out + self.functions)

    def add_function(self, body_fmt) -> str:
        """A helper device function (used for matrix-vector products: one row per component)."""
        if not hasattr(self, "functions"):
            self.functions = []
        for k, (fmt, _) in enumerate(self._function_formats):
            if fmt == body_fmt:
                return f"nbk_fn{k}"
        name = f"nbk_fn{len(self.functions)}"
        self._function_formats.append((body_fmt, name))
        self.functions.append(body_fmt.format(name=name))
        return name


_TABLES: "_Tables | None" = None


def _as_vec(x, length) -> Vec:
    if isinstance(x, Grid):
        x = x.flat
    if isinstance(x, Vec):
        if x.length != length:
            raise TraceError(f"shape mismatch in a traced right-hand side: ({x.length},) vs ({length},)")
        return x
    if isinstance(x, np.ndarray) and x.ndim == 1:
        if x.size != length:
            raise TraceError(f"shape mismatch in a traced right-hand side: ({x.size},) vs ({length},)")
        name = _TABLES.add(x)
        return Vec(length, lambda j, name=name: f"{name}[{j}]")
    code = _sc(x)
    return Vec(length, lambda j, code=code: code)


def _len_of(x):
    if isinstance(x, Vec):
        return x.length
    if isinstance(x, np.ndarray) and x.ndim == 1:
        return x.size
    return None


def _vec_binary(a, b, op, fn=None):
    if isinstance(a, Grid):
        return a._bin(b, op, False, fn)
    if isinstance(b, Grid):
        return b._bin(a, op, True, fn)
    length = _len_of(a) if _len_of(a) is not None else _len_of(b)
    va, vb = _as_vec(a, length), _as_vec(b, length)
    if fn is None:
        return Vec(length, lambda j, x=va.code, y=vb.code: f"({x(j)} {op} {y(j)})")
    return Vec(length, lambda j, x=va.code, y=vb.code: fn(x(j), y(j)))


def _vec_unary(a, fn):
    a = _as_vec(a, _len_of(a))
    return Vec(a.length, lambda j, x=a.code: fn(x(j)))


def _vec_pow(a, e):
    if isinstance(e, (Vec, np.ndarray)) and not (isinstance(e, np.ndarray) and e.ndim == 0):
        return _vec_binary(a, e, None, lambda x, y: f"pow({x}, {y})")
    if not isinstance(a, Vec):
        return _vec_binary(a, e, None, lambda x, y: f"pow({x}, {y})")
    return Vec(a.length, lambda j, x=a.code: _pow_code(x(j), e))


def _vec_matmul(m, v):
    """``M @ v`` for a matrix that is constant (a captured 2-D NumPy array) or the parameter block itself (``p`` of
    shape (r, n)): component i is the dot product of row i with the vector, emitted as one helper function with a
    loop instead of r * n unrolled terms."""
    if not isinstance(v, Vec):
        raise TraceError("only matrix @ vector products can be traced")
    n = v.length
    if isinstance(m, np.ndarray) and m.ndim == 2 and m.shape[1] == n and m.dtype != object:
        tab = _TABLES.add(m)
        row = lambda i, tab=tab: f"{tab}[({i}) * {n} + j]"  # noqa: E731
    elif isinstance(m, np.ndarray) and m.ndim == 2 and m.shape[1] == n and m.dtype == object:
        first = m[0, 0]
        if not (isinstance(first, Sc) and first.c.startswith("p[") and all(
                isinstance(m[a, b], Sc) and m[a, b].c == f"p[{int(first.c[2:-1]) + a * n + b}]"
                for a in (0, m.shape[0] - 1) for b in (0, n - 1))):
            raise TraceError("matrix @ vector needs a NumPy matrix or a contiguous block of the parameters")
        off = int(first.c[2:-1])
        row = lambda i, off=off: f"p[{off} + ({i}) * {n} + j]"  # noqa: E731
    else:
        raise TraceError("matrix @ vector needs a 2-D matrix whose second dimension matches the vector")
    fn = _TABLES.add_function(
        "__device__ __forceinline__ double {name}(int i, double t, nbk_vec y, const double* p) {{\n"
        f"    double s = 0.0;\n    for (int j = 0; j < {n}; ++j) s = fma({row('i')}, {v.code('j')}, s);\n    return s;\n}}}}")
    return Vec(m.shape[0], lambda i, fn=fn: f"{fn}({i}, t, y, p)")


def _vec_sum(v):
    """Sum of all elements: a scalar every component can use, emitted as one helper function with a loop (each
    component re-evaluates it: O(n) work per component, like a matrix row)."""
    if not isinstance(v, Vec):
        raise TraceError("only traced vectors can be summed")
    fn = _TABLES.add_function(
        "__device__ __forceinline__ double {name}(double t, nbk_vec y, const double* p) {{\n"
        f"    double s = 0.0;\n    for (int j = 0; j < {v.length}; ++j) s += {v.code('j')};\n    return s;\n}}}}")
    return Sc(f"{fn}(t, y, p)")


def _vec_dot(a, b):
    """vector . vector (one of them may be a captured 1-D NumPy array)."""
    length = _len_of(a) if _len_of(a) is not None else _len_of(b)
    if length is None or not (isinstance(a, Vec) or isinstance(b, Vec)):
        raise TraceError("only dot products that involve a traced vector can be traced")
    return _vec_sum(_vec_binary(a, b, "*"))


@_implements(np.sum)
def _np_sum(a, axis=None, **kw):
    return _vec_sum(a)


@_implements(np.mean)
def _np_mean(a, axis=None, **kw):
    return _vec_sum(a) / float(a.length)


@_implements(np.dot)
def _np_dot(a, b, **kw):
    if isinstance(b, Vec) and isinstance(a, np.ndarray) and a.ndim == 2:
        return _vec_matmul(a, b)
    return _vec_dot(a, b)


@_implements(np.diff)
def _np_diff(a, n=1, axis=-1, **kw):
    if kw or int(n) < 0:
        raise TraceError("np.diff(a, n) without prepend / append can be traced")
    for _ in range(int(n)):
        a = a[1:] - a[:-1]
    return a


@_implements(np.where)
def _np_where(cond, a=None, b=None):
    if a is None or b is None:
        raise TraceError("np.where(cond, a, b) needs both branches to be traced")
    length = next((L for L in (_len_of(cond), _len_of(a), _len_of(b)) if L is not None), None)
    vc, va, vb = _as_vec(cond, length), _as_vec(a, length), _as_vec(b, length)
    return Vec(length, lambda j, c=vc.code, x=va.code, y=vb.code: f"({c(j)} ? {x(j)} : {y(j)})")


@_implements(np.clip)
def _np_clip(a, a_min=None, a_max=None, **kw):
    out = a
    if a_min is not None:
        out = _vec_binary(out, a_min, None, lambda x, y: f"fmax({x}, {y})")
    if a_max is not None:
        out = _vec_binary(out, a_max, None, lambda x, y: f"fmin({x}, {y})")
    return out


@_implements(np.hstack)
def _np_hstack(parts, **kw):
    return _np_concatenate(parts)


@_implements(np.append)
def _np_append(a, values, axis=None):
    return _np_concatenate((a, values))


@_implements(np.flip)
def _np_flip(a, axis=None):
    n = a.length
    return Vec(n, lambda j, c=a.code: c(f"({n - 1} - ({j}))"))


@_implements(np.roll)
def _np_roll(a, shift, axis=None):
    n, k = a.length, int(shift) % a.length
    if k == 0:
        return a
    return Vec(n, lambda j, c=a.code: c(f"((({j}) + {n - k}) % {n})"))


@_implements(np.concatenate)
def _np_concatenate(parts, axis=0):
    vecs = []
    for part in parts:
        if isinstance(part, Vec):
            vecs.append(part)
        elif isinstance(part, np.ndarray) and part.ndim == 1:
            vecs.append(_as_vec(part, part.size))
        elif isinstance(part, (list, tuple)):
            vecs.extend(Vec(1, lambda j, c=_sc(x): c) for x in part)
        else:
            vecs.append(Vec(1, lambda j, c=_sc(part): c))
    total = sum(v.length for v in vecs)

    def code(j, vecs=vecs):
        expr, start = None, total
        for v in reversed(vecs):
            start -= v.length
            piece = v.code(f"(({j}) - {start})" if start else j)
            expr = piece if expr is None else f"(({j}) < {start + v.length} ? {piece} : {expr})"
        return expr

    return Vec(total, code)


@_implements(np.zeros_like)
def _np_zeros_like(a, **kw):
    return Vec(a.length, lambda j: "0.0")


@_implements(np.empty_like)
def _np_empty_like(a, **kw):
    return Vec(a.length, lambda j: "0.0")


@_implements(np.ones_like)
def _np_ones_like(a, **kw):
    return Vec(a.length, lambda j: "1.0")


@_implements(np.asarray)
def _np_asarray(a, *args, **kw):
    return a


@_implements(np.copy)
def _np_copy(a, *args, **kw):
    return a.copy()


@_implements(np.gradient)
def _np_gradient(a, *args, **kw):
    raise TraceError("np.gradient is not traceable; write the stencil with slices or np.roll")



# --------------------------------------------------------------------------------------
# N-d views of a traced vector: method-of-lines right-hand sides on 2-D / 3-D grids
# --------------------------------------------------------------------------------------
# ``u = y.reshape(N, N)``, ``np.roll(u, 1, axis=0)``, ``u[1:-1, 2:]``, ``out[1:-1, 1:-1] = ...``, ``out.ravel()``: a
# `Grid` is a shape on top of a flat `Vec` (C order).  Element-wise arithmetic works on the flat vectors; rolls, slices
# and slice assignment turn the flat index j into the multi-index of the shape with integer division / modulo by
# compile-time constants (which the compiler strength-reduces) and back.


def _unravel(j: str, shape):
    """C expressions of the multi-index of flat index ``j`` in ``shape`` (C order)."""
    out, stride = [], 1
    strides = []
    for dim in reversed(shape):
        strides.append(stride)
        stride *= dim
    strides.reverse()
    for k, (dim, st) in enumerate(zip(shape, strides)):
        e = f"({j})" if st == 1 else f"(({j}) / {st})"
        out.append(e if k == 0 else f"({e} % {dim})")
    return out, strides


def _prod(shape) -> int:
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


class Grid:
    """A traced N-d array: ``flat`` (a `Vec` of prod(shape) elements, C order) seen with ``shape``."""

    __array_priority__ = 1002.0

    def __init__(self, flat, shape):
        shape = tuple(int(d) for d in shape)
        if not isinstance(flat, Vec) or flat.length != _prod(shape):
            raise TraceError(f"cannot view {getattr(flat, 'length', '?')} elements as shape {shape}")
        self.flat, self.shape = flat, shape

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: self.flat.length)
    dtype = np.dtype(np.float64)
    T = property(lambda self: self.transpose())

    def __len__(self):
        return self.shape[0]

    def _like(self, flat):
        return Grid(flat, self.shape)

    # ---- shape
    def reshape(self, *shape, **kw):
        return _reshape(self.flat, shape)

    def ravel(self, *a, **kw):
        return self.flat

    flatten = ravel

    def copy(self):
        return Grid(self.flat.copy(), self.shape)

    def transpose(self, *axes):
        axes = tuple(axes[0]) if len(axes) == 1 and isinstance(axes[0], (tuple, list)) else (axes or tuple(reversed(range(self.ndim))))
        new_shape = tuple(self.shape[a] for a in axes)
        _, src_strides = _unravel("0", self.shape)

        def code(j, base=self.flat.code):
            idx, _ = _unravel(j, new_shape)
            return base(" + ".join(f"{e} * {src_strides[a]}" for e, a in zip(idx, axes)))

        return Grid(Vec(self.size, code), new_shape)

    # ---- element-wise arithmetic: on the flat vectors
    def _operand(self, o):
        if isinstance(o, Grid):
            if o.shape != self.shape:
                raise TraceError(f"shape mismatch in a traced right-hand side: {o.shape} vs {self.shape}")
            return o.flat
        if isinstance(o, Vec):
            if self.ndim >= 1 and o.length == self.shape[-1]:  # a row vector against every row
                return Vec(self.size, lambda j, c=o.code, m=o.length: c(f"(({j}) % {m})"))
            raise TraceError(f"shape mismatch in a traced right-hand side: ({o.length},) vs {self.shape}")
        if isinstance(o, np.ndarray) and o.ndim > 0:
            try:
                return np.ascontiguousarray(np.broadcast_to(o, self.shape), dtype=np.float64).ravel()
            except ValueError as exc:
                raise TraceError(f"shape mismatch in a traced right-hand side: {o.shape} vs {self.shape}") from exc
        return o

    def _bin(self, o, op, swap=False, fn=None):
        a, b = (self._operand(o), self.flat) if swap else (self.flat, self._operand(o))
        return self._like(_vec_binary(a, b, op, fn))

    def __add__(self, o): return self._bin(o, "+")
    def __radd__(self, o): return self._bin(o, "+", True)
    def __sub__(self, o): return self._bin(o, "-")
    def __rsub__(self, o): return self._bin(o, "-", True)
    def __mul__(self, o): return self._bin(o, "*")
    def __rmul__(self, o): return self._bin(o, "*", True)
    def __truediv__(self, o): return self._bin(o, "/")
    def __rtruediv__(self, o): return self._bin(o, "/", True)
    def __lt__(self, o): return self._bin(o, "<")
    def __le__(self, o): return self._bin(o, "<=")
    def __gt__(self, o): return self._bin(o, ">")
    def __ge__(self, o): return self._bin(o, ">=")
    def __eq__(self, o): return self._bin(o, "==")
    def __ne__(self, o): return self._bin(o, "!=")
    __hash__ = None
    def __neg__(self): return self._like(-self.flat)
    def __pos__(self): return self
    def __abs__(self): return self._like(abs(self.flat))
    def __pow__(self, e): return self._like(_vec_pow(self.flat, self._operand(e)))
    def __rpow__(self, b): return self._like(_vec_pow(self._operand(b), self.flat))
    def sum(self, axis=None):
        if axis is not None:
            raise TraceError("sums along one axis of a traced grid cannot be traced; sum the whole grid or write CUDA C")
        return _vec_sum(self.flat)
    def mean(self, axis=None): return self.sum(axis) / float(self.size)

    def __bool__(self):
        raise TraceError("data-dependent control flow on a whole grid cannot be traced; write the right-hand side as CUDA C")

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            return NotImplemented
        ref = next(x for x in inputs if isinstance(x, Grid))
        out = Vec.__array_ufunc__(None, ufunc, method, *[ref._operand(x) if x is not ref else ref.flat for x in inputs], **kwargs)
        return ref._like(out) if isinstance(out, Vec) and out.length == ref.size else out

    def __array_function__(self, func, types, args, kwargs):
        if func in _GRID_HANDLED:
            return _GRID_HANDLED[func](*args, **kwargs)
        raise TraceError(f"numpy.{func.__name__} cannot be traced on an N-d grid; write the right-hand side as CUDA C "
                         "defining nbk_rhs_i(t, y, p, i, n)")

    # ---- indexing
    def _region(self, key):
        """-> per dimension (start, step, count or None for an integer index)"""
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            pos = next(i for i, k in enumerate(key) if k is Ellipsis)
            key = key[:pos] + (slice(None),) * (self.ndim - len(key) + 1) + key[pos + 1:]
        if len(key) > self.ndim:
            raise IndexError("too many indices for a traced grid")
        key = key + (slice(None),) * (self.ndim - len(key))
        region = []
        for k, dim in zip(key, self.shape):
            if isinstance(k, (numbers.Integral, np.integer)):
                i = int(k) + (dim if k < 0 else 0)
                if not 0 <= i < dim:
                    raise IndexError(k)
                region.append((i, 1, None))
            elif isinstance(k, slice):
                a, b, st = k.indices(dim)
                region.append((a, st, len(range(a, b, st))))
            else:
                raise TraceError("only integers and slices can index a traced grid")
        return region

    def __getitem__(self, key):
        region = self._region(key)
        out_shape = tuple(c for _, _, c in region if c is not None)
        _, strides = _unravel("0", self.shape)
        if not out_shape:
            return Sc(self.flat.code(str(sum(a * st for (a, _, _), st in zip(region, strides)))))

        def code(j, base=self.flat.code):
            idx, _ = _unravel(j, out_shape)
            terms, it = [], iter(idx)
            for (a, step, count), st in zip(region, strides):
                if count is None:
                    terms.append(str(a * st))
                else:
                    terms.append(f"({a} + ({step}) * {next(it)}) * {st}")
            return base(" + ".join(terms))

        flat = Vec(_prod(out_shape), code)
        return flat if len(out_shape) == 1 else Grid(flat, out_shape)

    def __setitem__(self, key, value):
        region = self._region(key)
        if any(step != 1 for _, step, c in region if c is not None):
            raise TraceError("only unit-stride slice assignments can be traced")
        sub_shape = tuple(c for _, _, c in region if c is not None)
        n_sub = _prod(sub_shape)
        if isinstance(value, Grid):
            if value.shape != sub_shape:
                raise TraceError(f"shape mismatch in a traced right-hand side: {value.shape} vs {sub_shape}")
            val = value.flat
        elif isinstance(value, np.ndarray) and value.ndim > 0:
            val = _as_vec(np.ascontiguousarray(np.broadcast_to(value, sub_shape), dtype=np.float64).ravel(), n_sub)
        else:
            val = _as_vec(value, n_sub)
        old = self.flat.code

        def code(j, old=old, v=val.code):
            idx, _ = _unravel(j, self.shape)
            conds, sub_idx = [], []
            for e, (a, _, count) in zip(idx, region):
                if count is None:
                    conds.append(f"{e} == {a}")
                else:
                    conds.append(f"{e} >= {a} && {e} < {a + count}")
                    sub_idx.append(f"({e} - {a})")
            flat_sub, stride = [], 1
            for e, dim in zip(reversed(sub_idx), reversed(sub_shape)):
                flat_sub.append(e if stride == 1 else f"{e} * {stride}")
                stride *= dim
            inner = " + ".join(reversed(flat_sub)) if flat_sub else "0"
            return f"(({' && '.join(conds)}) ? {v(inner)} : {old(j)})"

        self.flat = Vec(self.size, code)


def _reshape(flat, shape):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
        shape = tuple(shape[0])
    shape = [int(d) for d in shape]
    if shape.count(-1) == 1:
        known = _prod([d for d in shape if d != -1])
        shape[shape.index(-1)] = flat.length // known if known else 0
    if _prod(shape) != flat.length:
        raise TraceError(f"cannot reshape {flat.length} elements into shape {tuple(shape)}")
    return flat if len(shape) == 1 else Grid(flat, shape)


Vec.reshape = lambda self, *shape, **kw: _reshape(self, shape)
Vec.ravel = lambda self, *a, **kw: self
Vec.flatten = Vec.ravel

_GRID_HANDLED = {}
_NP_ZEROS, _NP_ONES = np.zeros, np.ones  # the originals (`_tracing` replaces the module attributes while a callable runs)


def _grid_implements(np_func):
    def deco(f):
        _GRID_HANDLED[np_func] = f
        return f
    return deco


@_grid_implements(np.roll)
def _grid_roll(a, shift, axis=None):
    if axis is None:
        return Grid(_np_roll(a.flat, shift), a.shape)
    if isinstance(axis, (tuple, list)):
        shifts = shift if isinstance(shift, (tuple, list)) else [shift] * len(axis)
        for sh, ax in zip(shifts, axis):
            a = _grid_roll(a, sh, ax)
        return a
    ax = int(axis) + (a.ndim if axis < 0 else 0)
    dim = a.shape[ax]
    k = int(shift) % dim
    if k == 0:
        return a
    _, strides = _unravel("0", a.shape)

    def code(j, base=a.flat.code):
        idx, _ = _unravel(j, a.shape)
        idx[ax] = f"(({idx[ax]} + {dim - k}) % {dim})"
        return base(" + ".join(f"{e} * {st}" for e, st in zip(idx, strides)))

    return a._like(Vec(a.size, code))


@_grid_implements(np.where)
def _grid_where(cond, a=None, b=None):
    ref = next(x for x in (cond, a, b) if isinstance(x, Grid))
    return ref._like(_np_where(ref._operand(cond) if cond is not ref else ref.flat,
                               ref._operand(a) if a is not ref else ref.flat, ref._operand(b) if b is not ref else ref.flat))


@_grid_implements(np.zeros_like)
def _grid_zeros_like(a, **kw):
    return Grid(_ConstVec(_NP_ZEROS(a.size)), a.shape)


@_grid_implements(np.empty_like)
def _grid_empty_like(a, **kw):
    return Grid(_ConstVec(_NP_ZEROS(a.size)), a.shape)


@_grid_implements(np.ones_like)
def _grid_ones_like(a, **kw):
    return Grid(_ConstVec(_NP_ONES(a.size)), a.shape)


@_grid_implements(np.reshape)
def _grid_reshape(a, shape, **kw):
    return _reshape(a.flat, shape if isinstance(shape, (tuple, list)) else (shape,))


@_grid_implements(np.ravel)
def _grid_ravel(a, **kw):
    return a.flat


@_grid_implements(np.transpose)
def _grid_transpose(a, axes=None):
    return a.transpose(*(axes,) if axes is not None else ())


@_grid_implements(np.sum)
def _grid_sum(a, axis=None, **kw):
    return a.sum(axis)


@_grid_implements(np.mean)
def _grid_mean(a, axis=None, **kw):
    return a.mean(axis)


@_grid_implements(np.clip)
def _grid_clip(a, a_min=None, a_max=None, **kw):
    return a._like(_np_clip(a.flat, a._operand(a_min) if a_min is not None else None, a._operand(a_max) if a_max is not None else None))


@_grid_implements(np.copy)
def _grid_copy(a, *args, **kw):
    return a.copy()


@_grid_implements(np.asarray)
def _grid_asarray(a, *args, **kw):
    return a


@_implements(np.reshape)
def _np_reshape(a, shape, **kw):
    return _reshape(a, shape if isinstance(shape, (tuple, list)) else (shape,))


@_implements(np.ravel)
def _np_ravel(a, **kw):
    return a


def trace_cta(func, n: int, params_shape, has_params: bool) -> CudaRHS:
    """Trace a NumPy-style whole-vector right-hand side into the component-wise ``nbk_rhs_i`` of the CTA regime.
    Control flow on scalars (``if t < t_on:``, ``if p[0] > 0:``) is explored path by path like in :func:`trace`; the
    paths become one nested select around the component expressions."""
    return _retry_if_converted(_trace_cta_paths, _python_function(func), n, params_shape, has_params)


def _trace_cta_paths(func, n: int, params_shape, has_params: bool) -> CudaRHS:
    global _TABLES, _SC_PATHS
    _TABLES = _Tables()

    def run(paths):
        global _SC_PATHS
        _SC_PATHS = paths
        y = Vec(n, lambda j: f"y[{j}]")
        t = Sc("t")
        try:
            if has_params:
                size = int(np.prod(params_shape, dtype=int))
                p = np.empty(size, dtype=object)
                for k in range(size):
                    p[k] = Sc(f"p[{k}]")
                out = func(t, y, p.reshape(params_shape))
            else:
                out = func(t, y)
        except TraceError:
            raise
        except Exception as exc:  # noqa: BLE001
            raise TraceError(
                f"could not trace {getattr(func, '__name__', func)!r} into a component-wise CUDA right-hand side "
                f"({type(exc).__name__}: {exc}); pass a CTA built-in ('rd1d', 'linear') or CUDA C source defining "
                "nbk_rhs_i(t, y, p, i, n)") from exc
        if isinstance(out, Grid):
            out = out.flat
        if isinstance(out, (list, tuple)) or (isinstance(out, np.ndarray) and out.dtype == object):
            out = _np_concatenate([[e] if not isinstance(e, Vec) else e for e in np.asarray(out, dtype=object).ravel()])
        return (_as_vec(out, n).code("i"),)

    def select(node):
        if node.is_leaf:
            return node.payload[0]
        return f"({node.cond} ? {select(node.true)} : {select(node.false)})"

    try:
        with _tracing([func], vector=True):
            body = select(_explore(run))
        src = _TABLES.source()
    finally:
        _TABLES = None
        _SC_PATHS = None
    src += ("\nNBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n) {\n"
            f"    return {body};\n}}\n")
    return CudaRHS(src, name=getattr(func, "__name__", "traced"))


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
        ev = _python_function(ev)
        try:
            nargs = len(inspect.signature(ev).parameters)
        except (TypeError, ValueError):
            nargs = 2

        def explore(ev, nargs=nargs):
            def run(paths):
                tr = _Trace(paths)
                try:
                    out = ev(*_symbolic_args(tr, n, params_shape, nargs >= 3 and has_params))
                except TraceError:
                    raise
                except Exception as exc:  # noqa: BLE001
                    raise TraceError(f"could not trace event {getattr(ev, '__name__', ev)!r} into CUDA C "
                                     f"({type(exc).__name__}: {exc}); pass nbkode_b200.CudaEvents instead") from exc
                out = np.asarray(out, dtype=object).ravel()
                if out.size != 1:
                    raise TraceError("an event function must return one float")
                return tr, Expr._c(out[0])

            with _tracing([ev]):
                return _Explored(_explore(run, "event function"))

        tree = _retry_if_converted(explore, ev).tree
        lines.append(f"  if (k == {k}) {{")
        _emit_tree(tree, lambda payload: [f"return {payload[1]};"], lines)
        lines.append("  }")
    lines.append("  return 0.0;")
    lines.append("}")
    return CudaEvents("\n".join(lines) + "\n", len(events), terminal, direction)


# Above this state size a traced right-hand side goes to the CTA regime when the solver has it: the per-thread
# kernels keep 7 stage vectors of n doubles per thread and leave the register file around n = 12 (measured on B200,
# RungeKutta45, 1 M decay trajectories: n = 16: 7.9 vs 11.4 ms, n = 24: 28 vs 11.8 ms, n = 32: 73 vs 11.8 ms,
# n = 64: 202 vs 22.5 ms, per-thread vs CTA).
CTA_AUTO_N = 16
SELECT_OVER_PATHS = 32  # a per-thread trace with more control-flow paths than this tries the select form first
GROUP_AUTO_N = 10
# largest n a traced callable still runs per thread, by method id (tools/bench_group.py, tools/bench_team_decay.py: the
# component-wise form then runs as sub-warp groups / teams, csrc/nbk_group.cuh, nbk_cta.cuh)
AUTO_N_BY_METHOD = {23: 10, 45: 10, 853: 5, 1: 17, 2: 17, 3: 17, 4: 17, 5: 17, 202: 20, 203: 20, 213: 20, 204: 20, 238: 20}


def resolve(rhs, n: int, params_shape, has_params: bool, regime: str = "auto", cta_ok: bool = True, group_ok: bool = False,
            auto_n=None):
    """-> (builtin_name or None, cuda_source or None).  ``regime``: "auto" | "thread" | "cta" (Python callables only:
    built-ins and CUDA C source fix their regime themselves)."""
    if isinstance(rhs, CudaRHS):
        return None, rhs.source
    if isinstance(rhs, str):
        if rhs.lower() in BUILTINS:
            # "decay" exists per thread and component-wise: mid-size systems under RungeKutta23 / RungeKutta45 take the
            # component-wise form (sub-warp regime) unless regime="thread" asks for the per-thread kernels
            return ("decay:thread" if rhs.lower() == "decay" and regime == "thread" else rhs.lower()), None
        if rhs.lower() == "decay:thread":
            return "decay:thread", None
        if "nbk_rhs" in rhs:
            return None, rhs
        raise ValueError(f"unknown built-in right-hand side {rhs!r}; known: {sorted(BUILTINS)}")
    if callable(rhs):
        if regime not in ("auto", "thread", "cta"):
            raise ValueError("regime must be 'auto', 'thread' or 'cta'")
        # RungeKutta23 / RungeKutta45 have the sub-warp kernels (csrc/nbk_group.cuh) for component-wise right-hand sides:
        # they overtake the per-thread kernels from n = 11 on (decay, n = 12 / 16: 1.4x / 2.2x)
        # (`auto_n`: the solver class's own measured crossover, see AUTO_N_BY_METHOD)
        auto_n = auto_n if auto_n is not None else (GROUP_AUTO_N if group_ok else CTA_AUTO_N)
        if regime == "thread" or (regime == "auto" and (n <= auto_n or not cta_ok)):
            if n > MAX_TRACED_N:
                raise ValueError(f"n = {n} needs the CTA regime, which this solver class does not have "
                                 "(RungeKutta23/45, AdamsBashforth*, ForwardEuler and the fixed-step Runge-Kutta classes do)")
            try:
                traced = trace(rhs, n, params_shape, has_params)
            except TraceError:
                # e.g. element-wise branches beyond MAX_TRACED_PATHS: the whole-vector tracer expresses them as selects
                if regime != "auto" or not cta_ok:
                    raise
                return None, trace_cta(rhs, n, params_shape, has_params).source
            if traced.paths > SELECT_OVER_PATHS and regime == "auto" and cta_ok:
                # np.where / np.maximum on the object array of a small system branch element by element (2 ** n paths
                # per thread); as a whole-vector expression the same callable is one select per component
                try:
                    return None, trace_cta(rhs, n, params_shape, has_params).source
                except TraceError:
                    pass
            return None, traced.source
        # large (and mid-size) systems run one CTA per trajectory: the right-hand side is traced as a whole-vector
        # expression; a function the vector tracer cannot handle falls back to the per-thread form while it can
        try:
            return None, trace_cta(rhs, n, params_shape, has_params).source
        except TraceError:
            if regime == "cta" or n > MAX_TRACED_N:
                raise
            return None, trace(rhs, n, params_shape, has_params).source
    raise TypeError("rhs must be a built-in name, CUDA C source, a CudaRHS or a traceable Python callable")
