"""If-conversion of a Python right-hand side for the tracer (second line of defence of ``rhs.trace`` / ``rhs.trace_cta``).

Path exploration (``rhs._explore``) re-runs the callable once per control-flow path, which is exact and wastes no
device work -- but n independent branches (``for i in range(n): if y[i] > 0: ...``, the numba spelling of an
element-wise right-hand side) are 2 ** n paths.  When the path budget is exceeded the tracer retries with the function
rewritten by this module: an ``if`` statement whose test is a traced condition runs BOTH arms and merges the values they
assigned into selects (``x = c ? x_then : x_else``), so independent branches cost one select each.  On the device both
arms are evaluated and the select picks one -- the same result as the branch numba compiles for the reference
(nbkode/core.py:125-132), since the arms of a right-hand side have no side effects beyond their assignments.

What is rewritten (``convert``): ``if`` / ``elif`` / ``else`` statements whose arms contain only assignments, ``pass``,
``for`` loops and nested convertible ``if`` statements (no calls made for their side effect); conditional expressions (``a if c else b``);
``and`` / ``or`` / ``not`` and chained comparisons inside the tests (Python would ask the traced condition for its truth
value); the builtins ``min`` / ``max``.  Arms with ``return`` / ``break`` / ``continue`` / ``while`` / ``try`` ... are left
alone and keep being explored path by path.  If the arms disagree on something that is not a number (an index, a flag, a
shape), ``NoConvert`` is raised at trace time and the caller reports the failure of plain path exploration.
"""

from __future__ import annotations

import ast
import inspect
import numbers
import textwrap

import numpy as np


class NoConvert(Exception):
    """The rewritten function met something a select cannot express."""


UNSET = type("Unset", (), {"__repr__": lambda self: "<unbound>"})()


# ----------------------------------------------------------------------------------------------------------------
# run-time helpers the rewritten function calls
# ----------------------------------------------------------------------------------------------------------------
def _types():
    from . import rhs

    return rhs


def is_traced_condition(c) -> bool:
    rhs = _types()
    return isinstance(c, rhs.Cond) or (isinstance(c, rhs.Sc) and c.boolean)


def _is_traced(x) -> bool:
    rhs = _types()
    return isinstance(x, (rhs.Expr, rhs.Cond, rhs.Sc, rhs.Vec))


def snapshot(value):
    """Copy of a container an arm may assign into (``dy[i] = ...`` mutates in place)."""
    rhs = _types()
    if value is UNSET:
        return value
    if isinstance(value, np.ndarray):
        return value.copy()
    if isinstance(value, rhs.Vec):
        return value.copy()
    if isinstance(value, list):
        return list(value)
    return value


def _select(c, a, b):
    rhs = _types()
    if isinstance(c, rhs.Cond):
        if isinstance(a, rhs.Vec) or isinstance(b, rhs.Vec):
            raise NoConvert("a whole vector inside a per-thread trace")
        return c.tr.new(f"({c.code} ? {rhs.Expr._c(a)} : {rhs.Expr._c(b)})")
    # whole-vector tracer: the condition is a boolean scalar
    if isinstance(a, rhs.Vec) or isinstance(b, rhs.Vec) or isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.where(c, a, b)  # dispatches to the tracer's np.where (a select per component)
    return rhs.Sc(f"({c.c} ? {rhs._sc(a)} : {rhs._sc(b)})")


def merge(c, a, b):
    """Value of a name after ``if c: (... a ...) else: (... b ...)``."""
    rhs = _types()
    if a is b:
        return a
    if a is UNSET or b is UNSET:
        # bound in one arm only and unbound before the `if` (a temporary of that arm): a correct program reads it again
        # only where that arm was taken, so the arm's value is the value
        return b if a is UNSET else a
    if isinstance(a, (bool, np.bool_)) and isinstance(b, (bool, np.bool_)):
        if bool(a) == bool(b):
            return a
        raise NoConvert("the arms of a traced `if` disagree on a flag")
    if isinstance(a, (numbers.Integral, np.integer)) and isinstance(b, (numbers.Integral, np.integer)):
        if int(a) == int(b):
            return a
        raise NoConvert("the arms of a traced `if` disagree on an integer (an index or a counter?)")
    scalar = (numbers.Real, np.floating, np.integer, rhs.Expr, rhs.Cond, rhs.Sc)
    if isinstance(a, scalar) and isinstance(b, scalar):
        if not _is_traced(a) and not _is_traced(b) and float(a) == float(b):
            return a
        return _select(c, a, b)
    if isinstance(a, rhs._ConstVec) and isinstance(b, rhs._ConstVec) and a.assembling and b.assembling and a.length == b.length:
        # two vectors still being assembled element by element: merged element by element (one select where they differ)
        out = rhs._ConstVec(a.values)
        for k in range(a.length):
            va, vb = a.element(k), b.element(k)
            if va is vb or (not isinstance(va, rhs.Sc) and not isinstance(vb, rhs.Sc) and va == vb):
                if isinstance(va, rhs.Sc):
                    out.traced[k] = va
                continue
            out.traced[k] = _select(c, va, vb)
        return out
    if isinstance(a, rhs.Vec) or isinstance(b, rhs.Vec):
        return _select(c, a, b)
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        if a.shape != b.shape:
            raise NoConvert("the arms of a traced `if` build arrays of different shapes")
        if a.dtype != object and b.dtype != object:
            if np.array_equal(a, b):
                return a
            if isinstance(c, rhs.Sc):
                return _select(c, a, b)
        out = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(a.shape):
            out[idx] = merge(c, a[idx], b[idx])
        return out
    if isinstance(a, (list, tuple)) and type(a) is type(b) and len(a) == len(b):
        return type(a)(merge(c, x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray) and isinstance(b, scalar) or isinstance(b, np.ndarray) and isinstance(a, scalar):
        aa, bb = np.broadcast_arrays(np.asarray(a, dtype=object), np.asarray(b, dtype=object))
        return merge(c, np.array(aa), np.array(bb))
    try:
        if a == b:
            return a
    except Exception:  # noqa: BLE001
        pass
    raise NoConvert(f"cannot merge {type(a).__name__} and {type(b).__name__} of a traced `if`")


def ifexp(c, then, orelse):
    """``then() if c else orelse()``"""
    if is_traced_condition(c):
        return merge(c, then(), orelse())
    return then() if c else orelse()


def _as_condition(x):
    rhs = _types()
    if is_traced_condition(x):
        return x
    if isinstance(x, rhs.Expr):
        return x != 0.0
    if isinstance(x, rhs.Sc):
        return x != 0.0
    return None


def logical_and(*thunks):
    """``a and b and ...`` without asking a traced condition for its truth value (Python's result for untraced operands)."""
    acc = None
    for k, th in enumerate(thunks):
        v = th()
        cv = _as_condition(v)
        if cv is None:
            if not v:
                if acc is None:
                    return v           # short circuit on a plain falsy value, like Python
                return acc & False
            if k == len(thunks) - 1 and acc is None:
                return v
            continue                    # a plain truthy operand does not change a conjunction
        acc = cv if acc is None else (acc & cv)
    return acc if acc is not None else v


def logical_or(*thunks):
    acc = None
    for k, th in enumerate(thunks):
        v = th()
        cv = _as_condition(v)
        if cv is None:
            if v:
                if acc is None:
                    return v
                return acc | True
            if k == len(thunks) - 1 and acc is None:
                return v
            continue
        acc = cv if acc is None else (acc | cv)
    return acc if acc is not None else v


def logical_not(x):
    cx = _as_condition(x)
    return (not x) if cx is None else ~cx


def builtin_min(*args, **kw):
    if kw or len(args) != 2 or not (_is_traced(args[0]) or _is_traced(args[1])):
        return min(*args, **kw)
    a, b = args
    return ifexp(b < a, lambda: b, lambda: a)  # Python: the first of equal minima


def builtin_max(*args, **kw):
    if kw or len(args) != 2 or not (_is_traced(args[0]) or _is_traced(args[1])):
        return max(*args, **kw)
    a, b = args
    return ifexp(b > a, lambda: b, lambda: a)


HELPERS = {
    "__nbk_cond": is_traced_condition, "__nbk_snap": snapshot, "__nbk_merge": merge, "__nbk_ifexp": ifexp,
    "__nbk_and": logical_and, "__nbk_or": logical_or, "__nbk_not": logical_not, "__nbk_min": builtin_min,
    "__nbk_max": builtin_max, "__NBK_UNSET": UNSET,
}


# ----------------------------------------------------------------------------------------------------------------
# the rewrite
# ----------------------------------------------------------------------------------------------------------------
_SIMPLE = (ast.Assign, ast.AugAssign, ast.AnnAssign, ast.Expr, ast.Pass, ast.For, ast.If)
_FORBIDDEN = (ast.Return, ast.Break, ast.Continue, ast.While, ast.Try, ast.With, ast.Raise, ast.Delete, ast.Global, ast.Nonlocal,
              ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef, ast.Yield, ast.YieldFrom, ast.Await, ast.Lambda, ast.Import,
              ast.ImportFrom, ast.NamedExpr, ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp, ast.Assert, ast.Match)


def _convertible(stmts) -> bool:
    for st in stmts:
        if not isinstance(st, _SIMPLE):
            return False
        if isinstance(st, ast.Expr) and not isinstance(st.value, ast.Constant):
            return False  # a call for its side effect (`terms.append(...)`): both arms would run it
        for node in ast.walk(st):
            if isinstance(node, _FORBIDDEN):
                return False
    return True


def _stored_names(stmts):
    """Names an arm may rebind or mutate in place: plain stores, and the bases of subscript / attribute stores."""
    names = []

    def base(node):
        while isinstance(node, (ast.Subscript, ast.Attribute)):
            node = node.value
        return node.id if isinstance(node, ast.Name) else None

    for st in stmts:
        for node in ast.walk(st):
            targets = []
            if isinstance(node, ast.Assign):
                targets = node.targets
            elif isinstance(node, (ast.AugAssign, ast.AnnAssign)):
                targets = [node.target]
            elif isinstance(node, ast.For):
                targets = [node.target]
            for t in targets:
                for leaf in ast.walk(t):
                    if isinstance(leaf, ast.Name) and isinstance(leaf.ctx, ast.Store) and leaf.id not in names:
                        names.append(leaf.id)
                    elif isinstance(leaf, (ast.Subscript, ast.Attribute)) and isinstance(leaf.ctx, ast.Store):
                        b = base(leaf)
                        if b is not None and b not in names:
                            names.append(b)
    return names


def _stmts(code: str):
    return ast.parse(textwrap.dedent(code)).body


class _Rewriter(ast.NodeTransformer):
    def __init__(self):
        self.counter = 0
        self.converted = 0

    # ---- expressions
    def visit_BoolOp(self, node):
        self.generic_visit(node)
        fn = "__nbk_and" if isinstance(node.op, ast.And) else "__nbk_or"
        thunks = [ast.Lambda(args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]), body=v)
                  for v in node.values]
        return ast.Call(func=ast.Name(id=fn, ctx=ast.Load()), args=thunks, keywords=[])

    def visit_UnaryOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Not):
            return ast.Call(func=ast.Name(id="__nbk_not", ctx=ast.Load()), args=[node.operand], keywords=[])
        return node

    def visit_Compare(self, node):
        self.generic_visit(node)
        if len(node.ops) == 1:
            return node
        # a < b < c  ->  (a < b) and (b < c); the middle operands of a right-hand side are side-effect free
        parts, left = [], node.left
        for op, right in zip(node.ops, node.comparators):
            parts.append(ast.Compare(left=left, ops=[op], comparators=[right]))
            left = right
        return self.visit_BoolOp(ast.BoolOp(op=ast.And(), values=parts))

    def visit_IfExp(self, node):
        self.generic_visit(node)
        lam = lambda body: ast.Lambda(args=ast.arguments(posonlyargs=[], args=[], kwonlyargs=[], kw_defaults=[], defaults=[]), body=body)  # noqa: E731
        return ast.Call(func=ast.Name(id="__nbk_ifexp", ctx=ast.Load()), args=[node.test, lam(node.body), lam(node.orelse)], keywords=[])

    def visit_Call(self, node):
        self.generic_visit(node)
        if isinstance(node.func, ast.Name) and node.func.id in ("min", "max"):
            node.func = ast.Name(id="__nbk_" + node.func.id, ctx=ast.Load())
        return node

    # ---- statements
    def visit_If(self, node):
        # decided on the ORIGINAL arms (the rewrite of an inner `if` adds try / del statements of its own)
        ok = _convertible(node.body) and _convertible(node.orelse)
        names = _stored_names(node.body + node.orelse) if ok else []
        self.generic_visit(node)  # inner statements and the test
        if not ok:
            return node
        k = self.counter
        self.counter += 1
        self.converted += 1
        c = f"__nbk_c{k}"
        out = [ast.Assign(targets=[ast.Name(id=c, ctx=ast.Store())], value=node.test)]
        pre, then_grab, restore, merge_back = [], [], [], []
        for j, name in enumerate(names):
            p, a = f"__nbk_p{k}_{j}", f"__nbk_a{k}_{j}"
            pre += _stmts(f"""
                try:
                    {p} = __nbk_snap({name})
                except NameError:
                    {p} = __NBK_UNSET
            """)
            then_grab += _stmts(f"""
                try:
                    {a} = {name}
                except NameError:
                    {a} = __NBK_UNSET
            """)
            restore += _stmts(f"""
                if {p} is __NBK_UNSET:
                    try:
                        del {name}
                    except NameError:
                        pass
                else:
                    {name} = __nbk_snap({p})
            """)
            merge_back += _stmts(f"""
                try:
                    {p} = {name}
                except NameError:
                    {p} = __NBK_UNSET
                {p} = __nbk_merge({c}, {a}, {p})
                if {p} is __NBK_UNSET:
                    try:
                        del {name}
                    except NameError:
                        pass
                else:
                    {name} = {p}
            """)
        both = pre + node.body + then_grab + restore + (node.orelse or [ast.Pass()]) + merge_back
        plain = ast.If(test=ast.Name(id=c, ctx=ast.Load()), body=node.body, orelse=node.orelse)
        out.append(ast.If(test=ast.Call(func=ast.Name(id="__nbk_cond", ctx=ast.Load()), args=[ast.Name(id=c, ctx=ast.Load())], keywords=[]),
                          body=both, orelse=[plain]))
        return out


def convert(func):
    """The if-converted twin of ``func`` (same globals, closure values frozen), or None when there is nothing to convert
    or the source is not available (lambdas, built-ins, interactive definitions)."""
    try:
        source = textwrap.dedent(inspect.getsource(func))
        tree = ast.parse(source)
    except (OSError, TypeError, SyntaxError, IndentationError):
        return None
    if len(tree.body) != 1 or not isinstance(tree.body[0], ast.FunctionDef):
        return None
    fdef = tree.body[0]
    fdef.decorator_list = []
    rw = _Rewriter()
    try:
        rw.generic_visit(fdef)
    except RecursionError:
        return None
    if rw.converted == 0 and not any(isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id.startswith("__nbk_")
                                     for n in ast.walk(fdef)):
        return None
    ast.fix_missing_locations(tree)
    namespace = dict(getattr(func, "__globals__", {}))
    code = getattr(func, "__code__", None)
    if code is not None and func.__closure__:
        for name, cell in zip(code.co_freevars, func.__closure__):
            try:
                namespace[name] = cell.cell_contents
            except ValueError:
                return None
    namespace.update(HELPERS)
    try:
        exec(compile(tree, filename=f"<if-converted {getattr(func, '__name__', 'rhs')}>", mode="exec"), namespace)
    except Exception:  # noqa: BLE001
        return None
    new = namespace.get(fdef.name)
    if not callable(new):
        return None
    new.__nbk_original__ = func
    return new
