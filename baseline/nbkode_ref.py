"""The LIVE reference (hgrecco/numbakit-ode, numba mode) as a CPU arm -- reference arm of bench.py.

Nothing here is part of the product: nbkode_b200 never imports this module.  It drives the
UNMODIFIED reference package installed under the git-ignored ``baseline/_ref`` (``install()``
below: ``pip install --no-index --no-deps --target baseline/_ref`` from a /tmp copy of
/root/reference, the recipe of the task statement) through the reference's own public API:
one ``nbkode.RungeKutta45`` object per trajectory, ``run([T])`` (call chain
nbkode/core.py:106-141 ctor, :696-705 ``select_initial_step``, :598-619 ``_run_eval``,
runge_kutta/explicit.py:60-99 FSAL stage loop + retry loop).

Two adapters, neither of which edits the reference:
 * scipy >= 1.11 changed the private signature of ``select_initial_step`` (two extra positional
   arguments); ``nbkode.core.select_initial_step`` is rebound to a 4-line adapter that passes
   infinite bounds, which reproduces the old formula.
 * ``params=`` recompiles a closure per Solver object (~5 s each, nbkode/core.py:132), so the
   ensemble loop compiles ONE ``rhs(t, y)`` that reads the current trajectory's parameters
   from a fixed host buffer (numba ``carray`` over a constant address; SURVEY.md App. C.5 --
   bit-identical to the ``params=`` path, checked by tests/test_reference_live.py).

Workers: fork()ed after the parent has JIT-compiled every entry point, so children inherit
the machine code and the timed region holds integration only.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference"

_state = {}


def installed() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "nbkode", "core.py"))


def install(force: bool = False) -> str:
    """pip-install the unmodified reference into baseline/_ref (build container only)."""
    if installed() and not force:
        return "present"
    if not os.path.isdir(REF_SRC):
        return "unavailable: /root/reference absent and baseline/_ref not populated"
    tmp = "/tmp/nbkode_ref_src"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(REF_SRC, tmp)  # the build writes egg-info into the source tree; /root/reference is read-only
    shutil.rmtree(REF_DIR, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", REF_DIR, tmp]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0 or not installed():
        return "unavailable: pip install failed: " + (p.stdout + p.stderr).strip().splitlines()[-1][:200]
    return "installed"


def load():
    """Import the reference from baseline/_ref with the scipy adapter; returns the nbkode module."""
    if "nbkode" in _state:
        return _state["nbkode"]
    if not installed():
        raise ImportError("baseline/_ref/nbkode is missing: run baseline/nbkode_ref.py --install in the build container")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbkode_ref_numba_cache")
    os.environ.pop("NBKODE_NONUMBA", None)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    warnings.simplefilter("ignore")
    import scipy.integrate._ivp.common as sc

    import nbkode.core as core

    if sc.select_initial_step.__code__.co_argcount == 10:  # (fun, t0, y0, t_bound, max_step, f0, direction, order, rtol, atol)
        new = sc.select_initial_step
        core.select_initial_step = lambda fun, t0, y0, f0, direction, order, rtol, atol: new(
            fun, t0, y0, np.inf, np.inf, f0, direction, order, rtol, atol)
    import nbkode

    assert os.path.realpath(nbkode.__file__).startswith(os.path.realpath(REF_DIR)), nbkode.__file__
    _state["nbkode"] = nbkode
    return nbkode


def _param_rhs(kind: str, n_params: int):
    """One njit-compiled rhs(t, y) whose parameters live in a fixed host buffer."""
    import numba
    from numba import types
    from numba.extending import intrinsic

    buf = np.zeros(n_params)
    addr = int(buf.ctypes.data)

    @intrinsic
    def as_ptr(typingctx, a):
        sig = types.CPointer(types.float64)(types.uintp)

        def codegen(context, builder, signature, args):
            return builder.inttoptr(args[0], context.get_value_type(signature.return_type))

        return sig, codegen

    if kind == "lorenz":
        @numba.njit
        def rhs(t, y):
            p = numba.carray(as_ptr(numba.uintp(addr)), (3,), np.float64)
            out = np.empty(3)
            out[0] = p[0] * (y[1] - y[0])
            out[1] = y[0] * (p[1] - y[2]) - y[1]
            out[2] = y[0] * y[1] - p[2] * y[2]
            return out
    elif kind == "vdp":
        @numba.njit
        def rhs(t, y):
            p = numba.carray(as_ptr(numba.uintp(addr)), (1,), np.float64)
            out = np.empty(2)
            out[0] = y[1]
            out[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0]
            return out
    else:
        raise ValueError(kind)
    return rhs, buf


def prepare(kind="lorenz", cls="RungeKutta45", **kw):
    """Compile everything the workers will call (parent process, before fork)."""
    nbkode = load()
    key = (kind, cls, tuple(sorted(kw.items())))
    if _state.get("key") == key:
        return
    t0 = time.perf_counter()
    n_params = {"lorenz": 3, "vdp": 1}[kind]
    rhs, buf = _param_rhs(kind, n_params)
    y_warm = {"lorenz": np.array([1.0, 1.0, 20.0]), "vdp": np.array([1.0, 0.5])}[kind]
    buf[:] = {"lorenz": [10.0, 28.0, 8.0 / 3.0], "vdp": [1.0]}[kind]
    C = getattr(nbkode, cls)
    s = C(rhs, 0.0, y_warm, **kw)
    s.run(np.array([0.1]))
    s.step(n=4)
    _state.update(key=key, rhs=rhs, buf=buf, cls=C, kw=kw, jit_s=time.perf_counter() - t0)


def _integrate(args):
    """Worker: run([T]) per trajectory of a slice (timed), then -- optionally -- the accepted-step counts with
    a second, untimed pass (the reference has no counter; steps are never clipped to the output times, so the
    step sequence of step(n=...) equals the one inside run([T]))."""
    y0, p, ts, count = args
    rhs, buf, C, kw = _state["rhs"], _state["buf"], _state["cls"], _state["kw"]
    B, m = len(y0), len(ts)
    T = float(ts[-1])
    ys = np.empty((B, m, y0.shape[1]))
    t_end = np.empty(B)
    y_end = np.empty_like(y0)
    t0 = time.perf_counter()
    for i in range(B):
        buf[:] = p[i]
        s = C(rhs, 0.0, y0[i], **kw)
        _, y = s.run(ts)
        ys[i] = y
        t_end[i] = s.t
        y_end[i] = s.y
    wall = time.perf_counter() - t0
    nacc = np.zeros(B, dtype=np.int64)
    if count:
        for i in range(B):
            buf[:] = p[i]
            s = C(rhs, 0.0, y0[i], **kw)
            n = 0
            while s.t < T:
                tt, _ = s.step(n=64)
                k = int(np.searchsorted(tt, T, side="left"))  # first step whose end time is >= T closes run([T])
                n += min(k + 1, len(tt))
                if k < len(tt):
                    break
            nacc[i] = n
    return ys, t_end, y_end, nacc, wall


class Pool:
    """fork()ed workers, one per host core, contiguous slices."""

    def __init__(self, workers=None, kind="lorenz", cls="RungeKutta45", **kw):
        import multiprocessing as mp

        prepare(kind, cls, **kw)
        self.workers = workers or os.cpu_count() or 1
        self.jit_s = _state["jit_s"]
        self.pool = mp.get_context("fork").Pool(self.workers)

    def run(self, y0, p, ts, count=False):
        """Integrate the ensemble; wall = slowest worker's integration time (the job's critical path)."""
        y0 = np.ascontiguousarray(y0, dtype=float)
        p = np.ascontiguousarray(p, dtype=float)
        ts = np.atleast_1d(np.asarray(ts, dtype=float))
        bounds = np.linspace(0, len(y0), self.workers + 1).astype(int)
        jobs = [(y0[a:b], p[a:b], ts, count) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        t0 = time.perf_counter()
        parts = self.pool.map(_integrate, jobs, chunksize=1)
        span = time.perf_counter() - t0
        return {
            "ys": np.concatenate([q[0] for q in parts]), "t_end": np.concatenate([q[1] for q in parts]),
            "y_end": np.concatenate([q[2] for q in parts]), "n_accepted": np.concatenate([q[3] for q in parts]),
            "wall": max(q[4] for q in parts), "span": span,
        }

    def close(self):
        self.pool.close()
        self.pool.join()


if __name__ == "__main__":
    if "--install" in sys.argv:
        print(install(force="--force" in sys.argv))
