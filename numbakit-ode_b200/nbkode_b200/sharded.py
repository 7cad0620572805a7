"""One solver object over several GPUs of one process.

``nbkode_b200.RungeKutta45(rhs, t0, y0, params, devices=[0, 1, ..., 7], ...)`` splits the batch axis
into contiguous, near-equal slices, builds one ordinary solver per device and drives them from one
host thread each: every call (``run``, ``step``, ``skip``, ...) is launched on all devices at once --
the C-ABI allows different handles to be driven from different host threads, and both ctypes and
torch release the GIL while they wait -- and the per-device results are gathered once into the
caller's array type (NumPy, or a torch tensor on the device ``y0`` lives on).  There is no collective:
trajectories are independent, the output grid, the tableaux and shared parameters are replicated.

Results are bit-identical to the single-GPU solver (the kernels are position-invariant: a trajectory's
arithmetic does not depend on its index or on its neighbours), which tests/test_gpu_multi.py checks.
"""

from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import distributed as _dist
from . import rhs as _rhs


def _torch():
    import torch

    return torch


class ShardedSolver:
    """Batched solver whose ensemble lives on several devices.  Same methods and attributes as a batched
    :class:`nbkode_b200.core.Solver`; ``shards`` lists the per-device solvers."""

    def __init__(self, cls, rhs, t0, y0, params=None, *, devices, params_batched=None, **kwargs):
        torch = _torch()
        devices = [int(torch.device(d).index or 0) if not isinstance(d, int) else d for d in devices]
        if len(set(devices)) != len(devices) or not devices:
            raise ValueError("devices must be a non-empty list of distinct CUDA device indices")
        if any(d < 0 or d >= torch.cuda.device_count() for d in devices):
            raise ValueError(f"devices {devices} out of range: {torch.cuda.device_count()} CUDA device(s) visible")
        self._torch_out = isinstance(y0, torch.Tensor)
        self._out_device = y0.device if self._torch_out else None
        if (y0.dim() if self._torch_out else np.ndim(y0)) != 2:
            raise ValueError("devices=[...] needs a batched y0 of shape (B, n)")
        self.B, self.n = int(y0.shape[0]), int(y0.shape[1])
        if self.B < len(devices):
            devices = devices[: self.B]
        self.devices = devices
        self._cls = cls
        self.t_bound = kwargs.get("t_bound", np.inf)
        self.user_rhs = rhs
        self.params = params
        G = len(devices)
        self._bounds = [_dist.shard_bounds(self.B, g, G) for g in range(G)]
        # per-trajectory parameters are sliced with y0, shared ones replicated
        p_is_t = isinstance(params, torch.Tensor)
        p_ndim = 0 if params is None else (params.dim() if p_is_t else np.ndim(params))
        per_traj = params is not None and p_ndim >= 2 and int(params.shape[0]) == self.B and params_batched is not False
        h = kwargs.pop("h", None)
        first_step = kwargs.pop("first_step", None)

        def slice_b(x, lo, hi):  # per-trajectory option arrays (h, first_step) are sliced too
            if x is None or np.ndim(x) == 0:
                return x
            return x[lo:hi]

        rhs_g = [rhs]

        def build(g):
            lo, hi = self._bounds[g]
            kw = dict(kwargs)
            if h is not None:
                kw["h"] = slice_b(h, lo, hi)
            if first_step is not None:
                kw["first_step"] = slice_b(first_step, lo, hi)
            p_g = params[lo:hi] if per_traj else params
            return cls(rhs_g[0], t0, y0[lo:hi], p_g, device=devices[g],
                       params_batched=(True if per_traj else (False if params is not None and p_ndim >= 2 else None)), **kw)

        self._pool = ThreadPoolExecutor(max_workers=G, thread_name_prefix="nbk-dev")
        # the first shard resolves the right-hand side (a Python callable is traced to CUDA C once; the tracer is not
        # re-entrant), the others reuse the resolved name / source and are built concurrently
        s0 = build(0)
        rhs_g[0] = s0.rhs
        self.shards = [s0] + list(self._pool.map(build, range(1, G)))
        self._batched = True
        self.rhs = s0.rhs
        self.n_params = s0.n_params
        for name in ("options", "FIXED_STEP", "IMPLICIT", "GROUP", "LEN_HISTORY", "METHOD_ID", "STAGE_ROWS"):
            if hasattr(s0, name):
                setattr(self, name, getattr(s0, name))

    # ------------------------------------------------------------------ plumbing
    def _map(self, fn):
        """fn(shard) on every device at once; results in device order (exceptions propagate)."""
        return list(self._pool.map(fn, self.shards))

    def _gather(self, parts, axis=0):
        torch = _torch()
        if isinstance(parts[0], torch.Tensor):
            dev = self._out_device if self._out_device is not None else parts[0].device
            return torch.cat([p.to(dev, non_blocking=True) for p in parts], dim=axis)
        return np.concatenate(parts, axis=axis)

    def __del__(self):
        pool = getattr(self, "_pool", None)
        if pool is not None:
            pool.shutdown(wait=False)

    def __repr__(self):
        return f"<{self._cls.__name__} over devices {self.devices}: B={self.B}, n={self.n}>"

    # ------------------------------------------------------------------ state
    t = property(lambda self: self._gather(self._map(lambda s: s.t)))
    y = property(lambda self: self._gather(self._map(lambda s: s.y)))
    f = property(lambda self: self._gather(self._map(lambda s: s.f)))
    n_accepted = property(lambda self: self._gather(self._map(lambda s: s.n_accepted)))
    n_rejected = property(lambda self: self._gather(self._map(lambda s: s.n_rejected)))
    status = property(lambda self: self._gather(self._map(lambda s: s.status)))

    @property
    def h(self):
        hs = self._map(lambda s: s.h)
        return hs[0] if np.ndim(hs[0]) == 0 else self._gather(hs)

    @property
    def K(self):
        return self._gather(self._map(lambda s: s.K))

    @property
    def cache(self):
        return _ShardedCache(self)

    def counts(self, wait=True):
        parts = self._map(lambda s: s.counts(wait))
        return self._gather([a for a, _ in parts]), self._gather([b for _, b in parts])

    def synchronize(self):
        self._map(lambda s: s.synchronize())

    def raise_for_status(self):
        self._map(lambda s: s.raise_for_status())

    def last_launch(self):
        return self._map(lambda s: s.last_launch())

    # ------------------------------------------------------------------ drivers
    def run(self, t, wait: bool = True):
        parts = self._map(lambda s: s.run(t, wait))
        return parts[0][0], self._gather([y for _, y in parts])

    def step(self, *, n: int = None, upto_t: float = None):
        parts = self._map(lambda s: s.step(n=n, upto_t=upto_t))
        ts, ys = [p[0] for p in parts], [p[1] for p in parts]
        width = max(int(x.shape[1]) for x in ts)
        if any(int(x.shape[1]) != width for x in ts):  # step(upto_t=...): shards may have taken different numbers of steps
            ts = [_pad_nan(x, width) for x in ts]
            ys = [_pad_nan(x, width) for x in ys]
        return self._gather(ts), self._gather(ys)

    def skip(self, *, n: int = None, upto_t: float = None) -> None:
        self._map(lambda s: s.skip(n=n, upto_t=upto_t))

    def interpolate(self, t: float):
        return self._gather(self._map(lambda s: s.interpolate(t)))

    def run_events(self, t, events, max_events: int = 64):
        if events and not isinstance(events, _rhs.CudaEvents):
            # traced once, here (the tracer is not re-entrant); the shards get the CUDA C
            s0 = self.shards[0]
            events = _rhs.trace_events(events, s0.n, s0._p_shape, s0.params is not None)
        parts = self._map(lambda s: s.run_events(t, events, max_events))
        if not events:
            return parts[0][0], self._gather([p[1] for p in parts]), [], []
        torch = _torch()
        m = max((int(np.shape(p[0])[-1]) for p in parts), default=0)
        if any(np.ndim(p[0]) == 2 for p in parts):  # a terminal event stopped a trajectory somewhere: per-trajectory times
            tt = []
            for p, s in zip(parts, self.shards):
                ti = p[0]
                if np.ndim(ti) == 1:
                    ti = (ti.unsqueeze(0).expand(s.B, -1) if isinstance(ti, torch.Tensor) else np.broadcast_to(ti, (s.B, m)).copy())
                tt.append(ti)
            t_ret = self._gather(tt)
        else:
            t_ret = parts[0][0]
        ne = len(parts[0][2])
        t_ev, y_ev = [], []
        for j in range(ne):
            width = max(int(p[2][j].shape[1]) for p in parts)
            t_ev.append(self._gather([_pad_nan(p[2][j], width) for p in parts]))
            y_ev.append(self._gather([_pad_nan(p[3][j], width) for p in parts]))
        self.event_counts = np.concatenate([s.event_counts for s in self.shards], axis=0)
        return t_ret, self._gather([p[1] for p in parts]), t_ev, y_ev


def _pad_nan(x, width):
    """Pad axis 1 of a (B, k, ...) array with NaN up to `width`."""
    k = int(x.shape[1])
    if k == width:
        return x
    torch = _torch()
    shape = (int(x.shape[0]), width - k) + tuple(int(v) for v in x.shape[2:])
    if isinstance(x, torch.Tensor):
        return torch.cat([x, torch.full(shape, float("nan"), dtype=x.dtype, device=x.device)], dim=1)
    return np.concatenate([x, np.full(shape, np.nan, dtype=x.dtype)], axis=1)


class _ShardedCache:
    """History window of all shards with the reference's attribute names (nbkode/buffer.py)."""

    def __init__(self, solver):
        self._s = solver

    ts = property(lambda self: self._s._gather(self._s._map(lambda s: s.cache.ts), axis=1))  # (L, B)
    ys = property(lambda self: self._s._gather(self._s._map(lambda s: s.cache.ys)))           # (B, L, n)
    fs = property(lambda self: self._s._gather(self._s._map(lambda s: s.cache.fs)))
    t = property(lambda self: self._s.t)
    y = property(lambda self: self._s.y)
    f = property(lambda self: self._s.f)
