"""Solver objects: the drop-in boundary of the package.

Same constructor, methods, attributes and exceptions as ``nbkode.core.Solver``
(nbkode/core.py:37-481) for the explicit solvers, extended so that ``y0`` (and ``params``)
may carry a leading batch axis.  All numerical work happens in libnbkode_b200.so on the
GPU; this module only moves arguments across the C-ABI (include/nbkode_b200.h) and shapes
the results.

Shapes.  ``y0`` of shape ``()`` or ``(n,)`` gives an *unbatched* solver that returns exactly
what the reference returns (``t`` float, ``y (n,)``, ``run -> (m,), (m, n)``, ``step(n=N) ->
(N,), (N, n)``).  ``y0`` of shape ``(B, n)`` gives a *batched* solver: ``t (B,)``, ``y (B, n)``,
``run -> (m,), (B, m, n)``, ``step(n=N) -> (B, N), (B, N, n)``.  NumPy in -> NumPy out;
``torch.Tensor`` in -> ``torch.Tensor`` out on the input's device (CUDA tensors never leave
the GPU).
"""

from __future__ import annotations

import ctypes as C
import math
import os
from abc import ABC
from typing import Optional

import numpy as np

from . import _lib
from . import rhs as _rhs
from .util import CaseInsensitiveDict

_INT_MAX = 2**31 - 1
# method ids the CTA-per-trajectory kernels implement (include/nbkode_b200.h)
_CTA_METHODS = frozenset({23, 45, 853, 1, 2, 3, 4, 5, 202, 203, 213, 204, 238})

# device copies of recently used output grids: uploading a grid from pageable host memory makes the
# driver synchronise the stream, which would serialise back-to-back passes on the same grid
_GRID_CACHE: dict = {}
_GRID_CACHE_MAX = 32


def _device_grid(grid_host: np.ndarray, device):
    torch = _torch()
    key = (str(device), grid_host.tobytes())
    hit = _GRID_CACHE.get(key)
    if hit is not None:
        return hit
    pinned = torch.from_numpy(np.ascontiguousarray(grid_host)).pin_memory()
    dev = pinned.to(device, non_blocking=True)
    if len(_GRID_CACHE) >= _GRID_CACHE_MAX:
        _GRID_CACHE.pop(next(iter(_GRID_CACHE)))
    if grid_host.nbytes <= 1 << 20:
        _GRID_CACHE[key] = dev
    # a cached grid may be used from other streams later: make sure the upload has landed
    torch.cuda.current_stream(device).synchronize()
    return dev


def _torch():
    import torch

    return torch


def default_device() -> int:
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.NbkError("no CUDA device available: nbkode_b200 has no CPU fallback")
    if "LOCAL_RANK" in os.environ and "NBK_DEVICE" not in os.environ:
        return int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
    return int(os.environ.get("NBK_DEVICE", torch.cuda.current_device()))


class _Cache:
    """Read-only view of the history window with the reference's attribute names
    (nbkode/buffer.py: ``ts``, ``ys``, ``fs``, ``t``, ``y``, ``f``)."""

    def __init__(self, solver):
        self._s = solver

    @property
    def ts(self):  # (L,) or (L, B)
        s = self._s
        full = s._ts if s._regime == 0 else s._ts.t()
        return s._out(full[:, 0] if not s._batched else full)

    def _window(self, arr):  # (L, n) or (B, L, n)
        s = self._s
        full = arr.permute(2, 0, 1) if s._regime == 0 else arr
        return s._out(full[0] if not s._batched else full)

    @property
    def ys(self):
        return self._window(self._s._ys)

    @property
    def fs(self):
        return self._window(self._s._fs)

    t = property(lambda self: self._s.t)
    y = property(lambda self: self._s.y)
    f = property(lambda self: self._s.f)


class Solver(ABC):
    """Base class of all solvers (counterpart of nbkode/core.py:37-481)."""

    SOLVERS = CaseInsensitiveDict()
    SOLVERS_BY_GROUP = CaseInsensitiveDict()

    GROUP: str = None
    IMPLICIT: bool = False
    FIXED_STEP: bool = True
    ALIASES: tuple = ()
    LEN_HISTORY: int = 2
    METHOD_ID: int = 0
    STAGE_ROWS: int = 0
    OUT_LAYOUT: str = "bmn"  # device layout of run()'s dense output: "bmn" = (B, m, n) as returned

    def __init_subclass__(cls, abstract=False, **kwargs):
        super().__init_subclass__(**kwargs)
        if abstract:
            return
        if not isinstance(cls.LEN_HISTORY, int):
            raise ValueError(f"{cls.__name__}.LEN_HISTORY must be an integer.")
        if cls.LEN_HISTORY < 2:
            raise ValueError(f"While defining {cls.__name__}, LEN_HISTORY cannot be smaller than 1")
        for name_or_alias in (cls.__name__,) + tuple(cls.ALIASES):
            if name_or_alias in cls.SOLVERS:
                raise Exception(
                    f"Duplicate name/alias {cls.__name__} in {cls} collides with {cls.SOLVERS[name_or_alias]}"
                )
            cls.SOLVERS[name_or_alias] = cls
        cls.SOLVERS_BY_GROUP.setdefault(cls.GROUP, [])
        cls.SOLVERS_BY_GROUP[cls.GROUP].append(cls)

    # ------------------------------------------------------------------ construction
    def __new__(cls, *args, devices=None, **kwargs):
        # devices=[0, 1, ...]: one solver per device behind one object (one device / stream per slice,
        # launch all, one gather) -- see sharded.py
        if devices is not None and len(devices) > 1:
            from .sharded import ShardedSolver

            return ShardedSolver(cls, *args, devices=list(devices), **kwargs)
        return super().__new__(cls)

    def __init__(self, rhs, t0, y0, params=None, *, h=None, t_bound=np.inf, device=None, devices=None,
                 params_batched: Optional[bool] = None, regime: str = "auto", _options=None, _first_step=None,
                 _first_stepper=0):
        torch = _torch()
        self.t_bound = t_bound
        self.user_rhs = rhs
        self._handle = None
        if devices is not None:
            if len(devices) != 1 or device is not None:
                raise ValueError("pass either device= or devices=[...]")
            device = devices[0]
        self._dev_index = default_device() if device is None else (
            int(device) if isinstance(device, (int, np.integer)) else int(torch.device(device).index or 0))
        self._device = torch.device("cuda", self._dev_index)

        # ---- y0: () / (n,) unbatched, (B, n) batched
        self._torch_out = isinstance(y0, torch.Tensor)
        self._out_device = y0.device if self._torch_out else None
        y0_t = self._to_device(y0)
        if y0_t.ndim > 2:
            raise ValueError("y0 must have shape (), (n,) or (B, n)")
        self._batched = y0_t.ndim == 2
        self._y0_shape = tuple(y0_t.shape[-1:]) if y0_t.ndim >= 1 else (1,)
        y0_t = y0_t.reshape(1, -1) if not self._batched else y0_t
        self.B, self.n = int(y0_t.shape[0]), int(y0_t.shape[1])

        # ---- params: shared vector or one row per trajectory
        self.params = None
        p_t, shared, p_shape = None, 1, ()
        if params is not None:
            p_t = self._to_device(params)
            if not isinstance(params, torch.Tensor):
                self.params = np.ascontiguousarray(params)
            else:
                self.params = params
            per_traj = self._batched and p_t.ndim >= 2 and p_t.shape[0] == self.B and params_batched is not False
            if per_traj and params_batched is None and self.B > 1 and tuple(p_t.shape) == (self.n, self.n):
                # an n x n array with B == n: one shared matrix or one parameter row per trajectory?
                if isinstance(rhs, str) and rhs == "linear":
                    per_traj = False  # the built-in y' = A y takes exactly one shared n x n matrix
                elif not isinstance(rhs, str):
                    import warnings

                    warnings.warn(
                        f"params of shape {tuple(p_t.shape)} is ambiguous for a batch of {self.B} trajectories with n = {self.n}; "
                        "taken as one parameter row per trajectory -- pass params_batched=False for one shared array "
                        "(or params_batched=True to silence this warning)", stacklevel=3)
            if params_batched and not per_traj:
                raise ValueError("params_batched=True needs params of shape (B, ...)")
            if per_traj:
                p_shape = tuple(p_t.shape[1:])
                p_t = p_t.reshape(self.B, -1)
                shared = 0
            else:
                p_shape = tuple(p_t.shape)
                p_t = p_t.reshape(1, -1)
                shared = 1
        self.n_params = 0 if p_t is None else int(p_t.shape[1])
        self._params_shared = shared
        self._p_shape = p_shape if p_shape else (self.n_params,)
        self._events_key = None

        # ---- right-hand side -> built-in name or CUDA C
        # Python callables are traced for the regime that suits the state size (``regime="thread"`` / ``"cta"`` force
        # one, e.g. "thread" to use run_events or DOP853 on a mid-size system)
        cta_ok = self.METHOD_ID in _CTA_METHODS and not (2 <= self.METHOD_ID <= 5 and 1 <= _first_stepper <= 5)
        name, src = _rhs.resolve(rhs, self.n, p_shape if p_shape else (self.n_params,), params is not None, regime, cta_ok,
                                 group_ok=self.METHOD_ID in (23, 45), auto_n=_rhs.AUTO_N_BY_METHOD.get(self.METHOD_ID))
        if name in ("decay", "decay:thread") and self.n_params not in (1, self.n):
            raise ValueError("built-in 'decay' needs 1 or n parameters")
        self.rhs = name if name is not None else _rhs.CudaRHS(src)

        opts = _options or {}
        cfg = _lib.NbkConfig(
            self.METHOD_ID, self.n, self.n_params, shared, self.B,
            name.encode() if name else None, src.encode() if src else None,
            float(opts.get("rtol", 1e-3)), float(opts.get("atol", 1e-6)), float(opts.get("min_factor", 0.2)),
            float(opts.get("max_factor", 10.0)), float(opts.get("safety_factor", 0.9)),
            int(_first_stepper), self._dev_index,
        )
        handle = C.c_void_p()
        _lib.check(_lib.lib().nbk_create(C.byref(cfg), C.byref(handle)))
        self._handle = handle
        L, SR = C.c_int(0), C.c_int(0)
        _lib.check(_lib.lib().nbk_layout(handle, C.byref(L), C.byref(SR)))
        assert L.value == self.LEN_HISTORY and SR.value == self.STAGE_ROWS
        # 0: one trajectory per thread (trajectory index fastest); 1: one CTA per trajectory (a trajectory's
        # n values contiguous) -- see nbk_regime() in include/nbkode_b200.h
        self._regime = int(_lib.lib().nbk_regime(handle))
        # regime 1 only: lanes of a warp per trajectory in the advance kernels (sub-warp regime), 0 = the whole CTA
        self._group = int(_lib.lib().nbk_group_lanes(handle))

        # ---- ensemble state in HBM
        B, n, dev = self.B, self.n, self._device
        f64 = torch.float64
        if self._regime == 0:
            self._ts = torch.empty((L.value, B), dtype=f64, device=dev)
            self._ys = torch.empty((L.value, n, B), dtype=f64, device=dev)
            self._fs = torch.empty((L.value, n, B), dtype=f64, device=dev)
            self._K = torch.empty((SR.value, n, B), dtype=f64, device=dev) if SR.value else None
            p_shape_dev = (self.n_params,) if shared else (self.n_params, B)
        else:
            self._ts = torch.empty((B, L.value), dtype=f64, device=dev)
            self._ys = torch.empty((B, L.value, n), dtype=f64, device=dev)
            self._fs = torch.empty((B, L.value, n), dtype=f64, device=dev)
            self._K = torch.empty((B, SR.value, n), dtype=f64, device=dev) if SR.value else None
            p_shape_dev = (self.n_params,) if shared else (B, self.n_params)
        self._h = torch.empty((B,), dtype=f64, device=dev)
        self._p = torch.empty(p_shape_dev, dtype=f64, device=dev) if self.n_params else None
        self._nacc = torch.empty((B,), dtype=torch.int64, device=dev)
        self._nrej = torch.empty((B,), dtype=torch.int64, device=dev)
        self._status = torch.empty((B,), dtype=torch.int32, device=dev)
        self._nout = torch.zeros((B,), dtype=torch.int32, device=dev)
        self._queue = torch.zeros((8 + (3 * 4096 if os.environ.get('NBK_DEBUG_COMPACT') else 0),), dtype=torch.int64, device=dev)  # [queue head, failures of the last call, 6 diagnostic words]
        ptr = lambda x: x.data_ptr() if x is not None else None  # noqa: E731
        self._state = _lib.NbkState(
            ptr(self._ts), ptr(self._ys), ptr(self._fs), ptr(self._K), ptr(self._h), ptr(self._p),
            ptr(self._nacc), ptr(self._nrej), ptr(self._status), ptr(self._nout), ptr(self._queue),
        )

        # ---- f0, initial step, multistep start-up: one kernel
        h_scalar, h_vec = math.nan, None
        hh = _first_step if not self.FIXED_STEP else h
        if hh is not None:
            if np.ndim(hh) == 0:
                h_scalar = float(hh)
            else:
                h_vec = self._to_device(hh).reshape(-1).contiguous()
                if h_vec.numel() != B:
                    raise ValueError("per-trajectory step sizes must have shape (B,)")
        elif self.FIXED_STEP:
            h_scalar = 1.0  # the reference's default (nbkode/core.py:134-137)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nbk_init(
                handle, C.byref(self._state), float(t0), y0_t.data_ptr(), y0_t.stride(0), y0_t.stride(1),
                ptr(p_t), 0 if p_t is None else p_t.stride(0), 1 if p_t is None else p_t.stride(1),
                h_scalar, ptr(h_vec), self._stream(),
            ))
        self._keepalive = (y0_t, p_t, h_vec)
        if self.FIXED_STEP:
            # like the reference, a fixed-step solver keeps the user's value (or the integer 1)
            self.h = np.array(h, dtype=float) if h is not None and np.ndim(h) == 0 else (1 if h is None else h)
        # the constructor kernel's status is checked at the first synchronising call (run / step / a read of
        # t, y, f, status, cache, ...), so constructing a solver does not stall the stream
        self._unchecked = "constructor"
        self._oldest_host = float(t0)  # oldest history time, known on the host until the first advance

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _lib.lib().nbk_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        self._last_stream = _torch().cuda.current_stream(self._device)
        return self._last_stream.cuda_stream

    def _to_device(self, x):
        torch = _torch()
        if isinstance(x, torch.Tensor):
            return x.to(device=self._device, dtype=torch.float64, non_blocking=True)
        arr = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        return torch.from_numpy(arr).to(self._device, non_blocking=True)

    def _out(self, x, wait=True):
        """Device tensor -> the caller's array type (NumPy, or a torch tensor on the device the
        inputs came from; host tensors are returned in pinned memory).  With ``wait=False`` the copy
        to pinned memory is only queued: call :meth:`synchronize` before reading it."""
        if self._torch_out:
            if self._out_device.type == "cuda":
                return x
            torch = _torch()
            host = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            host.copy_(x, non_blocking=True)
            if wait:
                torch.cuda.current_stream(self._device).synchronize()
            return host
        return x.detach().cpu().numpy()

    def synchronize(self):
        """Wait for everything queued by this solver on the current stream and raise if a trajectory
        failed.  Needed after ``run(..., wait=False)`` / ``counts(wait=False)`` before the returned
        pinned host tensors are read."""
        stream = getattr(self, "_last_stream", None) or _torch().cuda.current_stream(self._device)
        stream.synchronize()
        with _torch().cuda.stream(stream):
            self._raise_on_failure(self._unchecked or "a previous call")

    def counts(self, wait=True):
        """(n_accepted, n_rejected) per trajectory in the caller's array type."""
        return self._out(self._nacc, wait), self._out(self._nrej, wait)

    def _host(self, x):
        return x.detach().cpu().numpy()

    def raise_for_status(self):
        """Synchronise and raise if a kernel queued earlier reported failed trajectories (the check
        is deferred while results stay on the GPU, like CUDA's own asynchronous error model)."""
        self._raise_on_failure(self._unchecked or "a previous call")

    def _raise_on_failure(self, where):
        self._unchecked = None
        # one 8-byte copy (no kernel: a reduction could not start while a persistent kernel of another stream
        # fills the GPU, which would serialise pipelined ensembles)
        bad = int(self._queue[1].item())
        if bad:
            raise RuntimeError(
                f"{bad} of {self.B} trajectories failed in {where}: non-finite error norm, collapsed step size or "
                "attempt budget exhausted (status codes in solver.status; the reference would not terminate)"
            )

    def _check_time(self, t):
        if t > self.t_bound:
            raise ValueError(f"Time {t} is larger than solver bound time t_bound={self.t_bound}")

    # ------------------------------------------------------------------ state
    def _checked(self):
        """Deferred failure report: a state read is a synchronising call (it copies to the host or hands out
        device data the caller will trust), so a failure of an earlier, unchecked kernel surfaces here instead
        of as silent NaN rows."""
        if self._unchecked:
            self._raise_on_failure(self._unchecked)

    @property
    def cache(self):
        # built on demand: a stored view would form a reference cycle and keep the ensemble's HBM
        # alive until the cyclic garbage collector runs
        self._checked()
        return _Cache(self)

    def _slot_t(self, slot):  # (B,) history times of one slot (0 oldest, -1 newest)
        return self._ts[slot] if self._regime == 0 else self._ts[:, slot]

    def _slot(self, arr, slot):  # (B, n)
        return arr[slot].t() if self._regime == 0 else arr[:, slot]

    @property
    def t(self):
        self._checked()
        cur = self._slot_t(-1)
        return float(cur[0].item()) if not self._batched else self._out(cur)

    @property
    def y(self):
        self._checked()
        cur = self._slot(self._ys, -1)
        return self._out(cur[0].reshape(self._y0_shape)) if not self._batched else self._out(cur)

    @property
    def f(self):
        self._checked()
        cur = self._slot(self._fs, -1)
        return self._out(cur[0].reshape(self._y0_shape)) if not self._batched else self._out(cur)

    @property
    def n_accepted(self):
        return int(self._nacc[0].item()) if not self._batched else self._out(self._nacc)

    @property
    def n_rejected(self):
        return int(self._nrej[0].item()) if not self._batched else self._out(self._nrej)

    @property
    def status(self):
        return int(self._status[0].item()) if not self._batched else self._out(self._status)

    @property
    def is_jit(self):
        return bool(_lib.lib().nbk_is_jit(self._handle))

    def last_launch(self):
        g, b, r = C.c_int(0), C.c_int(0), C.c_int(0)
        _lib.check(_lib.lib().nbk_last_launch(self._handle, C.byref(g), C.byref(b), C.byref(r)))
        out = {"grid": g.value, "block": b.value, "regs": r.value}
        lanes = int(_lib.lib().nbk_group_lanes(self._handle))
        if lanes:
            out["lanes_per_trajectory"] = lanes
        return out

    # ------------------------------------------------------------------ step / skip
    def _advance(self, n_steps, bound, emit):
        torch = _torch()
        self._oldest_host = None
        B, n, dev = self.B, self.n, self._device
        t_out = y_out = None
        if emit:
            t_out = torch.full((B, n_steps), math.nan, dtype=torch.float64, device=dev)
            y_out = torch.full((B, n_steps, n), math.nan, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nbk_step(
                self._handle, C.byref(self._state), int(n_steps), float(bound), int(bool(emit)),
                t_out.data_ptr() if emit else None, n_steps,
                y_out.data_ptr() if emit else None, n_steps * n, n, 1, 0, self._stream(),
            ))
        self._raise_on_failure("step")
        return t_out, y_out, self._nout

    def _shape_steps(self, t_out, y_out, counts):
        if not self._batched:
            k = int(counts[0].item())
            return self._out(t_out[0, :k]), self._out(y_out[0, :k].reshape((k,) + self._y0_shape))
        return self._out(t_out), self._out(y_out)

    def step(self, *, n: int = None, upto_t: float = None):
        """Advance ``n`` steps or until the next step would go beyond ``upto_t``, returning
        every intermediate ``(t, y)`` (nbkode/core.py:214-267).  Batched solvers return padded
        arrays ``(B, N)``, ``(B, N, n)`` with NaN where a trajectory took fewer steps."""
        torch = _torch()
        if upto_t is not None and not self._batched and upto_t < self.t:
            return np.asarray([]), np.asarray([])
        if n is None and upto_t is None:
            t_out, y_out, counts = self._advance(1, self.t_bound, True)
            if not self._batched:
                if int(counts[0].item()) == 0:
                    return None
                return np.atleast_1d(self.t), self.y
            return self._shape_steps(t_out, y_out, counts)
        if upto_t is None:
            t_out, y_out, counts = self._advance(int(n), self.t_bound, True)
            if int(counts.min().item()) < int(n):
                raise RuntimeError("Integrator reached t_bound.")
            return self._shape_steps(t_out, y_out, counts)
        self._check_time(upto_t)
        if n is not None:
            return self._shape_steps(*self._advance(int(n), upto_t, True))
        # only upto_t: the number of steps is not known in advance -> fixed-size chunks
        chunk, ts_parts, ys_parts = 256, [], []
        total = torch.zeros((self.B,), dtype=torch.int64, device=self._device)
        while True:
            t_out, y_out, counts = self._advance(chunk, upto_t, True)
            ts_parts.append(t_out)
            ys_parts.append(y_out)
            total += counts
            if int(counts.max().item()) < chunk:
                break
        t_all, y_all = torch.cat(ts_parts, 1), torch.cat(ys_parts, 1)
        if not self._batched:
            k = int(total[0].item())
            return self._out(t_all[0, :k]), self._out(y_all[0, :k].reshape((k,) + self._y0_shape))
        # chunks of a trajectory are contiguous prefixes only while it keeps running, and a
        # trajectory that stopped stays stopped: the concatenation is already left-packed
        kmax = int(total.max().item())
        return self._out(t_all[:, :kmax]), self._out(y_all[:, :kmax])

    def skip(self, *, n: int = None, upto_t: float = None) -> None:
        """Like :meth:`step` without output (nbkode/core.py:269-313)."""
        if upto_t is not None and not self._batched and upto_t < self.t:
            return
        if n is None and upto_t is None:
            self._advance(1, self.t_bound, False)
        elif upto_t is None:
            _, _, counts = self._advance(int(n), self.t_bound, False)
            if int(counts.min().item()) < int(n):
                raise RuntimeError("Integrator reached t_bound.")
        elif n is None:
            self._check_time(upto_t)
            self._advance(_INT_MAX - 1, upto_t, False)
        else:
            self._check_time(upto_t)
            self._advance(int(n), upto_t, False)

    # ------------------------------------------------------------------ run / interpolate
    def _run_grid(self, grid_host: np.ndarray, bound: float, wait: bool = True):
        torch = _torch()
        B, n, dev, m = self.B, self.n, self._device, int(grid_host.size)
        grid = _device_grid(grid_host, dev)
        fill = torch.empty if math.isinf(bound) else (lambda *a, **k: torch.full(*a, math.nan, **k))
        layout = os.environ.get("NBK_OUT_LAYOUT") or self.OUT_LAYOUT
        if layout == "mnb" and self._regime == 0 and B > 1:
            # device layout [m][n][B]: a warp's 32 trajectories write 32 consecutive doubles per output
            # point and component (fixed-step kernels advance in lock-step); returned as a (B, m, n) view
            store = fill((m, n, B), dtype=torch.float64, device=dev)
            y_out, strides = store.permute(2, 0, 1), (1, n * B, B)
        else:
            y_out = fill((B, m, n), dtype=torch.float64, device=dev)
            strides = (m * n, n, 1)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nbk_run(
                self._handle, C.byref(self._state), grid.data_ptr(), m, float(bound),
                y_out.data_ptr(), strides[0], strides[1], strides[2], 0, self._stream(),
            ))
        self._oldest_host = None
        if (self._torch_out and self._out_device.type == "cuda" and self._batched) or not wait:
            self._unchecked = "run"  # results stay on the device (or the caller asked not to wait): failed
            # trajectories (sticky status, NaN outputs) are reported by the next synchronising call
        else:
            self._raise_on_failure("run")
        return y_out

    def run(self, t, wait: bool = True):
        """Integrate, interpolating at each time of ``t`` (nbkode/core.py:315-332, 385-447).
        Steps are never clipped to the output times.  ``wait=False`` (batched solvers with host
        torch tensors) only queues the work and the copy of the result to pinned memory, so that
        several ensembles can be pipelined on different CUDA streams; see :meth:`synchronize`."""
        t = np.atleast_1d(np.asarray(t.detach().cpu() if hasattr(t, "detach") else t)).astype(np.float64)
        is_sorted = t.size == 1 or bool(np.all(t[:-1] <= t[1:]))
        if not is_sorted:
            ndx = np.argsort(t)
            t = t[ndx]
        # oldest history time: known on the host for a solver that has not advanced yet, otherwise
        # one small device reduction (failed trajectories keep their status inside the kernels, so
        # no failure pre-check is needed here)
        oldest = self._oldest_host if self._oldest_host is not None else float(self._slot_t(0).max().item())
        if t[0] < oldest:
            raise ValueError(
                f"Cannot interpolate at t={t[0]} as it is smaller than the current smallest value in history ({oldest})"
            )
        self._check_time(np.max(t))
        y_out = self._run_grid(t, self.t_bound, wait or not self._batched)
        if not self._batched:
            k = int(self._nout[0].item())  # < m only if the t_bound guard fired (prefix, core.py:611-617)
            ts, ys = t[:k], self._out(y_out[0, :k].reshape((k,) + self._y0_shape))
            if not is_sorted and k == t.size:
                ondx = np.argsort(ndx)
                return ts[ondx], ys[ondx]
            return ts, ys
        ys = y_out
        if not is_sorted:
            ondx = np.argsort(ndx)
            ys = ys[:, _torch().from_numpy(ondx).to(self._device)]
            t = t[ondx]
        return t, self._out(ys, wait)

    def run_events(self, t, events, max_events: int = 64):
        """:meth:`run` with event detection (nbkode/core.py:334-447, nbkode/event_handler.py).

        ``events``: a callable ``event(t, y)`` (or ``event(t, y, p)``), an iterable of them -- traced to CUDA C
        like a Python right-hand side, with the optional ``terminal`` / ``direction`` attributes of the SciPy
        convention -- or a :class:`nbkode_b200.CudaEvents`.  After every accepted step a sign change of an
        event function is located by bisection on the step's dense output (tolerance 1.48e-8 on the event
        value, at most 50 halvings) and recorded; a terminal event stops its trajectory and its ``(t, y)``
        replaces the output row the trajectory was integrating towards.

        Unbatched solvers return the reference's ``(t, y, t_events, y_events)`` (lists per event).  Batched
        solvers return ``t`` of shape ``(m,)`` -- or ``(B, m)`` when a terminal event stopped any trajectory,
        NaN after the stop --, ``y (B, m, n)`` (NaN after the stop), and per event the NaN-padded arrays
        ``t_events[k] (B, K_k)``, ``y_events[k] (B, K_k, n)``; ``solver.event_counts`` is ``(B, n_events)``.
        ``max_events`` bounds the events stored per trajectory and event function."""
        if not events:
            return self.run(t) + ([], [])
        torch = _torch()
        ev = _rhs.trace_events(events, self.n, self._p_shape, self.params is not None)
        if not (1 <= ev.n_events <= _lib.MAX_EVENTS):
            raise ValueError(f"between 1 and {_lib.MAX_EVENTS} event functions are supported")
        if self._events_key != (ev.n_events, ev.source):
            _lib.check(_lib.lib().nbk_set_events(self._handle, ev.n_events, ev.source.encode()))
            self._events_key = (ev.n_events, ev.source)
        t = np.atleast_1d(np.asarray(t.detach().cpu() if hasattr(t, "detach") else t)).astype(np.float64)
        is_sorted = t.size == 1 or bool(np.all(t[:-1] <= t[1:]))
        if not is_sorted:
            ndx = np.argsort(t)
            t = t[ndx]
        oldest = self._oldest_host if self._oldest_host is not None else float(self._slot_t(0).max().item())
        if t[0] < oldest:
            raise ValueError(
                f"Cannot interpolate at t={t[0]} as it is smaller than the current smallest value in history ({oldest})"
            )
        self._check_time(np.max(t))
        if t[0] <= float(self._slot_t(-1).max().item()):
            import warnings

            # the reference calls the non-existent warnings.warning here (core.py:441)
            warnings.warn("Events for past events are not implemented yet.", stacklevel=2)
        B, n, dev, m, NE, cap = self.B, self.n, self._device, int(t.size), ev.n_events, int(max_events)
        grid = _device_grid(t, dev)
        y_out = torch.full((B, m, n), math.nan, dtype=torch.float64, device=dev)
        ev_t = torch.full((B, NE, cap), math.nan, dtype=torch.float64, device=dev)
        ev_y = torch.full((B, NE, cap, n), math.nan, dtype=torch.float64, device=dev)
        ev_c = torch.zeros((B, NE), dtype=torch.int32, device=dev)
        t_term = torch.full((B,), math.nan, dtype=torch.float64, device=dev)
        direction = (C.c_int * NE)(*ev.direction)
        terminal = (C.c_int * NE)(*[int(x) for x in ev.terminal])
        block = _lib.NbkEvents(NE, cap, direction, terminal, ev_t.data_ptr(), ev_y.data_ptr(), ev_c.data_ptr(),
                               t_term.data_ptr())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nbk_run_events(
                self._handle, C.byref(self._state), grid.data_ptr(), m, float(self.t_bound),
                y_out.data_ptr(), m * n, n, 1, C.byref(block), 0, self._stream(),
            ))
        self._oldest_host = None
        status = self._status
        if bool((status == _lib.TRAJ_EVBRACKET).any().item()):
            self._unchecked = None
            raise ValueError("In bisect, 'fun(left, *args)' must have and opposite sign to 'fun(right, *args)'.")
        self._raise_on_failure("run_events")
        counts = ev_c.cpu().numpy()
        if counts.size and int(counts.max()) > cap:
            raise RuntimeError(f"a trajectory produced {int(counts.max())} events of one kind but max_events={cap}; "
                               "call run_events with a larger max_events")
        self.event_counts = counts[0] if not self._batched else counts
        stopped = status == _lib.TRAJ_EVENT
        if not self._batched:
            k = int(self._nout[0].item())
            ts, ys = t[:k].copy(), self._out(y_out[0, :k].reshape((k,) + self._y0_shape))
            if bool(stopped[0].item()):
                ts[k - 1] = float(t_term[0].item())
            et, ey = ev_t[0].cpu().numpy(), ev_y[0].cpu().numpy()
            t_events = [[float(v) for v in et[j, :counts[0, j]]] for j in range(NE)]
            y_events = [[ey[j, i].reshape(self._y0_shape).copy() for i in range(counts[0, j])] for j in range(NE)]
            if not is_sorted and k == t.size:
                ondx = np.argsort(ndx)
                return ts[ondx], ys[ondx], t_events, y_events
            return ts, ys, t_events, y_events
        ys = y_out
        if bool(stopped.any().item()):
            nout = self._nout.to(torch.int64)
            cols = torch.arange(m, device=dev)
            tt = torch.from_numpy(t).to(dev).repeat(B, 1)
            tt = torch.where(cols[None, :] < nout[:, None], tt, torch.full_like(tt, math.nan))
            rows = torch.nonzero(stopped).ravel()
            tt[rows, nout[rows] - 1] = t_term[rows]
            t_ret = self._out(tt)
        else:
            t_ret = t
        if not is_sorted and not bool(stopped.any().item()):
            ondx = np.argsort(ndx)
            ys = ys[:, torch.from_numpy(ondx).to(dev)]
            t_ret = t[ondx]
        t_events, y_events = [], []
        for j in range(NE):
            kj = int(counts[:, j].max()) if counts.size else 0
            t_events.append(self._out(ev_t[:, j, :kj]))
            y_events.append(self._out(ev_y[:, j, :kj]))
        return t_ret, self._out(ys), t_events, y_events

    def interpolate(self, t: float):
        """Dense output inside the recorded history (nbkode/core.py:449-472)."""
        if self._unchecked:
            self._raise_on_failure(self._unchecked)
        lo, hi = float(self._slot_t(0).max().item()), float(self._slot_t(-1).min().item())
        if not (lo <= t <= hi):
            raise ValueError(f"Time {t} to interpolate outside range ([{lo}, {hi}])")
        y_out = self._run_grid(np.array([float(t)]), math.inf)
        return self._out(y_out[0, 0].reshape(self._y0_shape)) if not self._batched else self._out(y_out[:, 0])


# ----------------------------------------------------------------------------------------
# registry (nbkode/core.py:709-800)
# ----------------------------------------------------------------------------------------

def check(solver, implicit=None, fixed_step=None, runge_kutta=None, multistep=None):
    if implicit is not None and solver.IMPLICIT is not implicit:
        return False
    if fixed_step is not None and solver.FIXED_STEP is not fixed_step:
        return False
    if runge_kutta is not None:
        from .runge_kutta import RungeKutta

        if issubclass(solver, RungeKutta) is not runge_kutta:
            return False
    if multistep is not None:
        from .multistep import Multistep

        if issubclass(solver, Multistep) is not multistep:
            return False
    return True


def get_solvers(*groups, implicit=None, fixed_step=None, runge_kutta=None, multistep=None):
    """Available solver classes, optionally filtered (nbkode/core.py:729-763)."""
    if not groups:
        groups = Solver.SOLVERS_BY_GROUP.keys()
    out = []
    for group in groups:
        try:
            members = Solver.SOLVERS_BY_GROUP[group]
        except KeyError:
            raise KeyError(f"Group {group} not found. Valid values: {tuple(Solver.SOLVERS_BY_GROUP.keys())}")
        out.extend(s for s in members if check(s, implicit, fixed_step, runge_kutta, multistep))
    return tuple(out)


def get_groups():
    return tuple(sorted(Solver.SOLVERS_BY_GROUP.keys()))


def list_solvers(fmt_string="{cls.__name__}", alias_fmt_string="{name} (alias of {cls.__name__})", include_alias=True):
    out = []
    for k, v in Solver.SOLVERS.items():
        if k == v.__name__:
            out.append(fmt_string.format(cls=v, name=k))
        elif include_alias:
            out.append(alias_fmt_string.format(cls=v, name=k))
    return out


def get_solver(name_or_alias):
    try:
        return Solver.SOLVERS[name_or_alias]
    except KeyError:
        pass
    valid = "- " + "\n- ".join(sorted(list_solvers()))
    raise ValueError(f"No solver named {name_or_alias}, valid options are:\n{valid}")
