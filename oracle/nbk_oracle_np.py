"""NumPy restatement of nbkode's explicit hot path (TEST INFRASTRUCTURE, not product code).

This module is an *oracle*: a single-trajectory CPU restatement of the algorithms the
reference (hgrecco/numbakit-ode, ``/root/reference``) runs for ForwardEuler,
AdamsBashforth1..5, RungeKutta23, RungeKutta45, DOP853 and the fixed-step explicit Runge-Kutta classes.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it; the
product package ``nbkode_b200`` never does.

Parity status: PINNED.  ``tests/golden/*.json`` were produced by running the live
reference in this container (``tests/golden/make_golden.py``); ``tests/test_oracle_np.py``
checks this restatement against them bit-for-bit for vector-valued systems
(it uses the same NumPy primitives in the same association order as the reference).

Each function cites the reference lines it restates (paths relative to /root/reference).
"""

from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------------------
# Tableaux (nbkode/runge_kutta/explicit.py:146-151 RK23, :185-239 RK45).  The values are
# the published Bogacki-Shampine 3(2) and Dormand-Prince 5(4) coefficients.
# --------------------------------------------------------------------------------------


class BS23:
    name = "RungeKutta23"
    q = 2  # error_estimator_order
    C = np.array([0, 1 / 2, 3 / 4], dtype=float)
    A = np.array([[0, 0, 0], [1 / 2, 0, 0], [0, 3 / 4, 0]], dtype=float)
    B = np.array([2 / 9, 1 / 3, 4 / 9], dtype=float)
    B2 = np.array([7 / 24, 1 / 4, 1 / 3, 1 / 8], dtype=float)
    # FSAL.E (explicit.py:54-58): E = -B2; E[:-1] += B
    E = -B2.copy()
    E[:-1] += B
    P = np.array([[1, -4 / 3, 5 / 9], [0, 1, -2 / 3], [0, 4 / 3, -8 / 9], [0, -1, 1]])


class DP45:
    name = "RungeKutta45"
    q = 4
    C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
    A = np.array(
        [
            [0, 0, 0, 0, 0, 0],
            [1 / 5, 0, 0, 0, 0, 0],
            [3 / 40, 9 / 40, 0, 0, 0, 0],
            [44 / 45, -56 / 15, 32 / 9, 0, 0, 0],
            [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0],
            [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0],
        ]
    )
    B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
    E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
    P = np.array(
        [
            [1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
            [0, 0, 0, 0],
            [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
            [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
            [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
            [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
            [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423],
        ]
    )


class DOP853Tab:
    """DOP853 (nbkode/runge_kutta/explicit.py:282-303).  The reference's
    nbkode/runge_kutta/dop853_coefficients.py is a copy of SciPy's module of the same name (Hairer's
    published 8(5,3) tableau); the oracle reads the numbers from the installed SciPy, and
    tests/golden/make_golden.py asserts they equal the reference's arrays."""

    name = "DOP853"
    q = 7

    def __init__(self):
        from scipy.integrate._ivp import dop853_coefficients as dc

        self.n_stages = dc.N_STAGES  # 12
        self.A = dc.A[:12, :12].copy()
        self.B = dc.B
        self.C = dc.C[:12]
        self.E3, self.E5, self.D = dc.E3, dc.E5, dc.D
        self.A_FULL, self.C_FULL = dc.A, dc.C
        self.A_EXTRA, self.C_EXTRA = dc.A[13:], dc.C[13:]


class _LazyMethods(dict):
    def __missing__(self, key):
        if key == "DOP853":
            self[key] = DOP853Tab()
            return self[key]
        raise KeyError(key)


RK_METHODS = _LazyMethods({"RungeKutta23": BS23, "RungeKutta45": DP45})

# Adams-Bashforth weights exactly as listed by the reference classes
# (nbkode/multistep/adams_bashforth.py:20,31,44,57,61).  NB the reference applies B[0]
# to the OLDEST derivative in the window (multistep/core.py:101,109 + buffer.py:48-56).
AB_B = {
    1: np.array([1.0]),
    2: np.array([3, -1]) / 2,
    3: np.array([23, -16, 5]) / 12,
    4: np.array([55, -59, 37, -9]) / 24,
    5: np.array([1901, -2774, 2616, -1274, 251]) / 720,
}


def _rms(x):
    # scipy.integrate._ivp.common.norm: ||x||_2 / sqrt(size)
    return np.linalg.norm(x) / x.size**0.5


def select_initial_step(rhs, t0, y0, f0, order, rtol, atol):
    """scipy.integrate._ivp.common.select_initial_step with direction=+1 and unbounded
    interval / max_step, as called from nbkode/core.py:695-705."""
    if y0.size == 0:
        return np.inf
    scale = atol + np.abs(y0) * rtol
    d0 = _rms(y0 / scale)
    d1 = _rms(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = 1e-6
    else:
        h0 = 0.01 * d0 / d1
    y1 = y0 + h0 * 1 * f0
    f1 = rhs(t0 + h0 * 1, y1)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1 / (order + 1))
    return min(100 * h0, h1)


class History:
    """Sliding window oldest->newest of (t, y, f); nbkode/buffer.py:12-56."""

    def __init__(self, length, t, y, f):
        n = y.size
        self.ts = np.empty(length)
        self.ys = np.empty((length, n))
        self.fs = np.empty((length, n))
        for _ in range(length):
            self.push(t, y, f)

    def push(self, t, y, f):
        self.ts[:-1] = self.ts[1:]
        self.ts[-1] = t
        self.ys[:-1] = self.ys[1:]
        self.ys[-1] = y
        self.fs[:-1] = self.fs[1:]
        self.fs[-1] = f

    @property
    def t(self):
        return self.ts[-1]

    @property
    def y(self):
        return self.ys[-1]

    @property
    def f(self):
        return self.fs[-1]


def _bind(rhs, params):
    if params is None:
        return rhs
    params = np.ascontiguousarray(params)
    return lambda t, y: rhs(t, y, params)


class AdaptiveRK:
    """RungeKutta23 / RungeKutta45 of the reference (FSAL pairs).

    ctor: nbkode/core.py:106-141, :687-706; runge_kutta/core.py:25-27.
    attempt: runge_kutta/explicit.py:60-79; error/norm/controller: runge_kutta/core.py:73-99;
    accept loop: explicit.py:81-99; t_bound guard: core.py:176-184;
    dense output: explicit.py:354-403.
    """

    def __init__(
        self,
        method,
        rhs,
        t0,
        y0,
        params=None,
        *,
        rtol=1e-3,
        atol=1e-6,
        first_step=None,
        t_bound=np.inf,
        min_factor=0.2,
        max_factor=10.0,
        safety_factor=0.9,
    ):
        self.M = RK_METHODS[method] if isinstance(method, str) else method
        self.rhs = _bind(rhs, params)
        self.rtol, self.atol = rtol, atol
        self.min_factor, self.max_factor, self.safety = min_factor, max_factor, safety_factor
        self.t_bound = t_bound
        t0 = float(t0)
        y0 = np.array(y0, dtype=float, ndmin=1)
        self.cache = History(2, t0, y0, self.rhs(t0, y0))
        self.S = len(self.M.A)  # stages without the FSAL one
        self.K = np.zeros((self.S + 1, y0.size))
        if first_step is None:
            first_step = select_initial_step(
                self.rhs, self.cache.t, self.cache.y, self.cache.f, self.M.q, rtol, atol
            )
        self.h = np.array(first_step, dtype=float)
        self.n_accepted = 0
        self.n_rejected = 0

    t = property(lambda self: self.cache.t)
    y = property(lambda self: self.cache.y)
    f = property(lambda self: self.cache.f)

    def _attempt(self):
        M, K, c, h = self.M, self.K, self.cache, self.h
        t, y = c.t, c.y
        K[0] = c.f
        for s in range(1, self.S):
            dy = M.A[s, :s] @ K[:s]
            K[s] = self.rhs(t + h * M.C[s], y + h * dy)
        t_new = t + h
        y_new = y + h * M.B @ K[:-1]
        K[-1] = self.rhs(t_new, y_new)
        return t_new, y_new

    def step(self, bound=None):
        """One accepted step. Returns False (without stepping) if t + h > bound."""
        bound = self.t_bound if bound is None else bound
        c = self.cache
        if c.t + self.h > bound:
            return False
        while True:
            t_new, y_new = self._attempt()
            scale = self.atol + self.rtol * np.maximum(np.abs(y_new), np.abs(c.y))
            if self.M.name == "DOP853":
                # explicit.py:325-336: two estimates combined; 0/0 -> nan when the step is exact
                err5 = np.sqrt(np.mean(((self.h * (self.M.E5 @ self.K)) / scale) ** 2))
                err3 = np.sqrt(np.mean(((self.h * (self.M.E3 @ self.K)) / scale) ** 2))
                with np.errstate(invalid="ignore", divide="ignore"):
                    norm = err5 / np.sqrt(1 + ((err3 / (10 * err5)) ** 2))
            else:
                err = self.h * (self.M.E @ self.K)
                norm = np.sqrt(np.mean((err / scale) ** 2))
            if norm == 0:
                self.h *= self.max_factor
                ok = True
            else:
                factor = self.safety / norm ** (1 / (self.M.q + 1))
                self.h *= min(self.max_factor, max(self.min_factor, factor))
                ok = norm < 1
            if ok:
                c.push(t_new, y_new, self.K[-1])
                self.n_accepted += 1
                return True
            self.n_rejected += 1

    def interpolate(self, t_eval):
        c = self.cache
        if t_eval == c.ts[-1]:
            return c.ys[-1].copy()
        if t_eval == c.ts[-2]:
            return c.ys[-2].copy()
        if self.M.name == "DOP853":
            return self._interpolate_dop853(t_eval)
        Q = self.K.T.dot(self.M.P)
        t_old, y_old = c.ts[-2], c.ys[-2]
        H = c.ts[-1] - t_old
        x = (t_eval - t_old) / H
        p = np.cumprod(np.repeat(x, Q.shape[1]))
        return H * np.dot(Q, p) + y_old


def _interpolate_dop853(self, t_eval):
    """DOP853Interpolator.update + evaluate (nbkode/runge_kutta/explicit.py:406-481): three extra
    stages on the last accepted step, F[0..6], alternating x / (1 - x) Horner scheme."""
    M, c = self.M, self.cache
    K = np.empty((16, c.y.size))
    K[:13] = self.K
    t_old, y_old, t, y = c.ts[-2], c.ys[-2], c.ts[-1], c.ys[-1]
    h = t - t_old
    f = c.f
    for s, (a, cc) in enumerate(zip(M.A_EXTRA, M.C_EXTRA), 13):
        dy = np.dot(K[:s].T, a[:s]) * h
        K[s] = self.rhs(t_old + cc * h, y_old + dy)
    F = np.empty((7, c.y.size))
    f_old = K[0]
    delta_y = y - y_old
    F[0] = delta_y
    F[1] = h * f_old - delta_y
    F[2] = 2 * delta_y - h * (f + f_old)
    F[3:] = h * np.dot(M.D, K)
    x = (t_eval - t_old) / h
    out = np.zeros_like(y_old)
    for i, fr in enumerate(F[::-1]):
        out += fr
        if i % 2 == 0:
            out *= x
        else:
            out *= 1 - x
    out += y_old
    return out


AdaptiveRK._interpolate_dop853 = _interpolate_dop853


class AdamsBashforth:
    """AdamsBashforth-k (k=1..5; ForwardEuler == AB1) with the reference's semantics.

    ctor/start-up: nbkode/multistep/core.py:37-65; step: :71-80, :90-112;
    coefficients: multistep/adams_bashforth.py; window Hermite: nbkode/core.py:577-596.
    ``first_stepper`` is None, "RungeKutta45"/"auto" (adaptive start-up, default
    tolerances) or an int k' (AdamsBashforth-k' fixed-step start-up).
    """

    def __init__(self, order, rhs, t0, y0, params=None, *, h=None, t_bound=np.inf, first_stepper="auto"):
        self.order = order
        self.B = AB_B[order]
        self.L = max(order, 2)
        self.rhs = _bind(rhs, params)
        self.raw_rhs, self.params = rhs, params
        self.t_bound = t_bound
        self.h = np.array(h, dtype=float) if h is not None else 1
        t0 = float(t0)
        y0 = np.array(y0, dtype=float, ndmin=1)
        self.cache = History(self.L, t0, y0, self.rhs(t0, y0))
        c = self.cache
        if first_stepper is None or order == 1:
            return
        if first_stepper == "auto":
            first_stepper = "RungeKutta45"
        if isinstance(first_stepper, int):
            sub = AdamsBashforth(first_stepper, self.rhs, c.t, c.y, h=self.h)
            for _ in range(order - 1):
                sub.step()
                c.push(sub.t, sub.y, sub.f)
        else:
            sub = AdaptiveRK(first_stepper, self.rhs, c.t, c.y)
            for ti in np.arange(1, order) * self.h:
                while sub.t < ti:
                    sub.step()
                yi = sub.interpolate(ti)
                c.push(ti, yi, self.rhs(ti, yi))

    t = property(lambda self: self.cache.t)
    y = property(lambda self: self.cache.y)
    f = property(lambda self: self.cache.f)

    def step(self, bound=None):
        bound = self.t_bound if bound is None else bound
        c, k = self.cache, self.order
        if c.t + self.h > bound:
            return False
        A = np.zeros_like(self.B)
        A[-1] = -1
        t_new = c.t + self.h
        y_new = self.h * self.B @ c.fs[self.L - k :] - A @ c.ys[self.L - k :]
        c.push(t_new, y_new, self.rhs(t_new, y_new))
        return True

    def interpolate(self, t_eval):
        c = self.cache
        t0, y0 = c.ts[0], c.ys[0]
        if t_eval == t0:
            return y0.copy()
        dt, dy = c.t - t0, c.y - y0
        f0, f1 = c.fs[0], c.f
        T = (t_eval - t0) / dt
        return y0 + T * dy + T * (T - 1) * ((1 - 2 * T) * dy + dt * ((T - 1) * f0 + T * f1))


# Implicit linear multistep classes (nbkode/multistep/adams_moulton.py:31-61, bdf.py:16-44, euler.py:30-36):
# name -> (A, B, Bn).  Adams-Moulton: A = [0, ..., 0, -1] of B's size; BDF: B = zeros of A's size.
def _am(B, Bn):
    B = np.array(B, dtype=float)
    A = np.zeros_like(B)
    A[-1] = -1
    return A, B, Bn


def _bdf(A, Bn):
    A = np.array(A, dtype=float)
    return A, np.zeros_like(A), Bn


IMPLICIT_TABLEAUX = {
    "AdamsMoulton1": _am([0.0], 1.0),
    "BackwardEuler": _am([0.0], 1.0),
    "AdamsMoulton2": _am([1 / 2], 1 / 2),
    "AdamsMoulton3": _am([-1 / 12, 2 / 3], 5 / 12),
    "AdamsMoulton4": _am(np.array([19, -5, 1]) / 24, 9 / 24),
    "AdamsMoulton5": _am(np.array([646, -264, 106, -19]) / 720, 251 / 720),
    "BDF1": _bdf([-1.0], 1.0),
    "BDF2": _bdf(np.array([1, -4]) / 3, 2 / 3),
    "BDF3": _bdf(np.array([-2, 9, -18]) / 11, 6 / 11),
    "BDF4": _bdf(np.array([3, -16, 36, -48]) / 25, 12 / 25),
    "BDF5": _bdf(np.array([-12, 75, -200, 300, -300]) / 137, 60 / 137),
    "BDF6": _bdf(np.array([10, -72, 225, -400, 450, -360]) / 147, 60 / 147),
}


def _secant(func, x0, tol=1.48e-8, maxiter=50):
    """scipy's `newton` without derivative as vendored in nbkode/nbcompat/zeros.py:350-392 (j_newton: any
    flag other than OK raises RuntimeError, :432-434)."""
    p0 = 1.0 * x0
    eps = 1e-4
    p1 = x0 * (1 + eps)
    p1 += eps if p1 >= 0 else -eps
    q0 = func(p0)
    q1 = func(p1)
    if np.abs(q1) < np.abs(q0):
        p0, p1, q0, q1 = p1, p0, q1, q0
    for _ in range(maxiter):
        if q1 == q0:
            if p1 != p0:
                raise RuntimeError
            return (p1 + p0) / 2.0
        if np.abs(q1) > np.abs(q0):
            p = (-q0 / q1 * p1 + p0) / (1 - q0 / q1)
        else:
            p = (-q1 / q0 * p0 + p1) / (1 - q1 / q0)
        if np.abs(p - p1) <= tol:
            return p
        p0, q0 = p1, q1
        p1 = p
        q1 = func(p1)
    raise RuntimeError


def _newton_hd(func, y, atol=1.48e-8, maxiter=50):
    """nbkode/nbcompat/zeros.py:440-505: Newton with a central finite-difference Jacobian (eps = 1e-10)."""
    def jacobian(x):
        eps = 1e-10
        J = np.zeros((len(x), len(x)), dtype=np.float64)
        for i in range(len(x)):
            x1 = x.copy()
            x2 = x.copy()
            x1[i] += eps
            x2[i] -= eps
            J[:, i] = (func(x1) - func(x2)) / (2 * eps)
        return J

    value = func(y)
    distance = np.linalg.norm(value, ord=2)
    it = 0
    while not np.abs(distance - 0) <= atol:
        delta = np.linalg.solve(jacobian(y), -value)
        y = y + delta
        value = func(y)
        distance = np.linalg.norm(value, ord=2)
        it += 1
        if it > maxiter:
            raise RuntimeError
    return y


class ImplicitMultistep:
    """AdamsMoulton1..5 / BackwardEuler / BDF1..6: ctor and start-up nbkode/multistep/core.py:37-65 (the same as the
    Adams-Bashforth classes), implicit step :115-170, dense output = window Hermite (nbkode/core.py:577-596)."""

    def __init__(self, name, rhs, t0, y0, params=None, *, h=None, t_bound=np.inf, first_stepper="auto"):
        self.A, self.B, self.Bn = IMPLICIT_TABLEAUX[name]
        self.order = self.A.size
        self.L = max(self.order, 2)
        self.rhs = _bind(rhs, params)
        self.t_bound = t_bound
        self.h = np.array(h, dtype=float) if h is not None else 1
        t0 = float(t0)
        y0 = np.array(y0, dtype=float, ndmin=1)
        self.cache = History(self.L, t0, y0, self.rhs(t0, y0))
        c = self.cache
        if first_stepper is None or self.order == 1:
            return
        if first_stepper == "auto":
            first_stepper = "RungeKutta45"
        if isinstance(first_stepper, int):
            sub = AdamsBashforth(first_stepper, self.rhs, c.t, c.y, h=self.h)
            for _ in range(self.order - 1):
                sub.step()
                c.push(sub.t, sub.y, sub.f)
        else:
            sub = AdaptiveRK(first_stepper, self.rhs, c.t, c.y)
            for ti in np.arange(1, self.order) * self.h:
                while sub.t < ti:
                    sub.step()
                yi = sub.interpolate(ti)
                c.push(ti, yi, self.rhs(ti, yi))

    t = property(lambda self: self.cache.t)
    y = property(lambda self: self.cache.y)
    f = property(lambda self: self.cache.f)

    def step(self, bound=None):
        bound = self.t_bound if bound is None else bound
        c, h = self.cache, self.h
        if c.t + h > bound:
            return False
        dA, dB = self.L - self.A.size, self.L - self.B.size
        t_new = c.t + h
        K = h * self.B @ c.fs[dB:] - self.A @ c.ys[dA:]
        h_Bn = h * self.Bn
        root = lambda y: y - h_Bn * self.rhs(t_new, y) - K  # noqa: E731
        if c.y.size == 1:
            y_new = _secant(root, c.y.copy())
        else:
            y_new = _newton_hd(root, c.y.copy())
        c.push(t_new, y_new, self.rhs(t_new, y_new))
        return True

    interpolate = AdamsBashforth.interpolate


# Fixed-step explicit Runge-Kutta tableaux (nbkode/runge_kutta/explicit.py:102-129)
ERK_TABLEAUX = {
    "Runge2": (np.array([[0, 0], [1 / 2, 0]]), np.array([0, 1.0]), np.array([0, 1 / 2])),
    "Runge3": (np.array([[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]), np.array([1, 4, 0, 1]) / 6,
               np.array([0, 1 / 2, 1, 1])),
    "Heun3": (np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0]]) / 3, np.array([1, 0, 3]) / 4, np.array([0, 1, 2]) / 3),
    "RungeKutta4": (np.array([[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 1 / 2, 0, 0], [0, 0, 1, 0]]),
                    np.array([1, 2, 2, 1]) / 6, np.array([0, 1, 1, 2]) / 2),
    "RungeKutta3_8": (np.array([[0, 0, 0, 0], [1 / 3, 0, 0, 0], [-1 / 3, 1, 0, 0], [1, -1, 1, 0]]),
                      np.array([1, 3, 3, 1]) / 8, np.array([0, 1, 2, 3]) / 3),
}


class FixedRK:
    """Fixed-step explicit Runge-Kutta classes of the reference (Runge2, Runge3, Heun3, RungeKutta4,
    RungeKutta3_8): stage loop nbkode/runge_kutta/explicit.py:28-46, step runge_kutta/core.py:33-42,
    dense output = window Hermite over the last step (nbkode/core.py:577-596)."""

    def __init__(self, name, rhs, t0, y0, params=None, *, h=None, t_bound=np.inf):
        self.A, self.B, self.C = ERK_TABLEAUX[name]
        self.rhs = _bind(rhs, params)
        self.t_bound = t_bound
        self.h = np.array(h, dtype=float) if h is not None else 1
        t0 = float(t0)
        y0 = np.array(y0, dtype=float, ndmin=1)
        self.cache = History(2, t0, y0, self.rhs(t0, y0))
        self.K = np.zeros((len(self.A), y0.size))

    t = property(lambda self: self.cache.t)
    y = property(lambda self: self.cache.y)
    f = property(lambda self: self.cache.f)

    def step(self, bound=None):
        bound = self.t_bound if bound is None else bound
        c, K, h = self.cache, self.K, self.h
        if c.t + h > bound:
            return False
        t, y = c.t, c.y
        K[0] = c.f
        for s in range(1, len(self.A)):
            dy = self.A[s, :s] @ K[:s]
            K[s] = self.rhs(t + h * self.C[s], y + h * dy)
        t = t + h
        y = y + h * self.B @ K
        c.push(t, y, self.rhs(t, y))
        return True

    interpolate = AdamsBashforth.interpolate


def run(solver, ts):
    """nbkode/core.py:598-619 (_run_eval) for points ahead of solver.t, preceded by the
    past-point branch of run_events (core.py:401-408).  ``ts`` must be sorted."""
    ts = np.atleast_1d(ts).astype(np.float64)
    out = []
    for ti in ts:
        if ti > solver.t:
            while solver.t < ti:
                if not solver.step():
                    return ts[: len(out)], np.array(out)
        out.append(solver.interpolate(ti))
    return ts, np.array(out).reshape(len(ts), -1)


def _bisect(fun, left, right, tol=1.48e-8, maxiter=50):
    """nbkode/nbcompat/zeros.py:508-533 (rtol = 0)."""
    if not left < right:
        raise ValueError("In bisect, 'left' must be smaller than 'rigth'.")
    yl = fun(left)
    if yl == 0:
        return left
    yr = fun(right)
    if yr == 0:
        return right
    if not yl * yr < 0:
        raise ValueError("In bisect, 'fun(left, *args)' must have and opposite sign to 'fun(right, *args)'.")
    for _ in range(maxiter):
        mid = (left + right) / 2
        ym = fun(mid)
        if np.abs(ym - 0) <= tol:
            break
        if np.sign(ym) == np.sign(yl):
            left, yl = mid, ym
        else:
            right, yr = mid, ym
    return mid


class _Event:
    """nbkode/event_handler.py:96-177."""

    def __init__(self, func, is_terminal, direction, init_t, init_y):
        self.func, self.is_terminal, self.direction = func, is_terminal, direction
        self.last_t = init_t
        self.last_value = func(init_t, init_y)
        self.t, self.y = [], []

    def evaluate(self, solver):
        t, y = solver.t, solver.y
        value = self.func(t, y)
        up = (self.last_value <= 0.0) & (value >= 0.0)
        down = (self.last_value >= 0.0) & (value <= 0.0)
        either = up | down
        trigger = up & (self.direction > 0) | down & (self.direction < 0) | either & (self.direction == 0)
        if trigger:
            if value == 0:
                root = t
            else:
                root = _bisect(lambda tt: self.func(tt, solver.interpolate(tt)), self.last_t, t)
            if not (self.t and self.t[-1] == root):
                self.t.append(root)
                self.y.append(solver.interpolate(root))
        self.last_t = t
        self.last_value = value
        return trigger and self.is_terminal


def run_events(solver, ts, events):
    """nbkode/core.py:385-447 + :621-647 (_run_eval_events) + event_handler.build_handler/EventHandler.evaluate.
    Returns (t, y, t_events, y_events); ``ts`` sorted and ahead of solver.t."""
    ts = np.atleast_1d(ts).astype(np.float64).copy()
    if callable(events):
        events = (events,)
    evs = [_Event(e, bool(getattr(e, "terminal", False)), int(np.sign(getattr(e, "direction", 0))), solver.t, solver.y.copy())
           for e in events]
    out = np.empty((ts.size, solver.y.size))
    n_rows = ts.size
    done = False
    for ndx, ti in enumerate(ts):
        while solver.t < ti:
            if not solver.step():
                n_rows, done = ndx, True
                break
            terminate, last_event = False, None
            for e in evs:
                if e.evaluate(solver):
                    terminate = True
                    last_event = (e.t[-1], e.y[-1]) if e.t else (np.nan, np.empty(0) * np.nan)
            if terminate:
                ts[ndx], out[ndx] = last_event
                n_rows, done = ndx + 1, True
                break
        if done:
            break
        out[ndx] = solver.interpolate(ti)
    return ts[:n_rows], out[:n_rows], [list(e.t) for e in evs], [list(e.y) for e in evs]


def steps(solver, n=None, upto_t=None):
    """nbkode/core.py:214-267 step(n=, upto_t=) -> (t[k], y[k, n])."""
    tt, yy = [], []
    bound = solver.t_bound if upto_t is None else upto_t
    while n is None or len(tt) < n:
        if not solver.step(bound):
            break
        tt.append(solver.t)
        yy.append(solver.y.copy())
    return np.array(tt), np.array(yy).reshape(len(tt), -1)


# --------------------------------------------------------------------------------------
# RHS catalogue used by the configs in BASELINE.json.
# --------------------------------------------------------------------------------------


def rhs_decay(t, y, p):
    return -p * y


def rhs_lorenz(t, y, p):
    return np.array([p[0] * (y[1] - y[0]), y[0] * (p[1] - y[2]) - y[1], y[0] * y[1] - p[2] * y[2]])


def rhs_vdp(t, y, p):
    return np.array([y[1], p[0] * (1 - y[0] * y[0]) * y[1] - y[0]])


def rhs_linear(t, y, A):
    return A @ y


def rhs_rd1d(t, y, p):
    """Periodic 1-D reaction diffusion: d*(y[i+1]-2y[i]+y[i-1]) + r*y[i]*(1-y[i])."""
    d, r = p[0], p[1]
    return d * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + r * y * (1 - y)


def rhs_pwl(t, y, p):
    """Oscillator with run-time branches on the state and the time (the reference compiles them with numba): a
    spring that is four times stiffer beyond a stop at |x| = p[1], a saturated damper written with Python's
    min / max, a forcing ramp that ends at t = 6.  Every piece joins continuously, so that a branch decided
    differently in the last bit next to a kink moves the result by rounding error only."""
    x, v = y[0], y[1]
    if x > p[1]:
        f = -p[0] * p[1] - 4.0 * p[0] * (x - p[1])
    elif x < -p[1]:
        f = p[0] * p[1] - 4.0 * p[0] * (x + p[1])
    else:
        f = -p[0] * x
    u = min(max(-p[2] * v, -0.5), 0.5)
    g = 0.3 * (6.0 - t) if t < 6.0 else 0.0
    return np.array([v, f + u + g])


def ev_switch(t, y):
    """An event function with a branch: the level it watches moves at t = 0.9."""
    if t < 0.9:
        return y[0] - 0.5
    return y[0] + 0.25


# event functions shared by the golden-vector generator and the tests (SciPy convention: optional
# ``terminal`` / ``direction`` attributes, nbkode/core.py:343-366)
def ev_y0(t, y):
    return y[0]


def ev_y1_down(t, y):
    return y[1]


ev_y1_down.direction = -1


def ev_z25_up(t, y):
    return y[2] - 25.0


ev_z25_up.direction = 1


def ev_x_m8_terminal(t, y):
    return y[0] + 8.0


ev_x_m8_terminal.terminal = True
ev_x_m8_terminal.direction = -1


def ev_time(t, y):
    return t - 2.5 + 0.0 * y[0]


def ev_y0_up_terminal(t, y):
    return y[0]


ev_y0_up_terminal.terminal = True
ev_y0_up_terminal.direction = 1

EVENTS = {"switch": ev_switch, "y0_up_terminal": ev_y0_up_terminal, "y0": ev_y0, "y1_down": ev_y1_down, "z25_up": ev_z25_up, "x_m8_terminal": ev_x_m8_terminal, "time": ev_time}

RHS = {"decay": rhs_decay, "lorenz": rhs_lorenz, "vdp": rhs_vdp, "linear": rhs_linear, "rd1d": rhs_rd1d, "pwl": rhs_pwl}


# input generators of the BASELINE configurations: defined in the package (so that the product harness never
# needs oracle/), re-exported here for the tests and the CPU baseline
import os as _os
import sys as _sys

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "numbakit-ode_b200")
if _pkg not in _sys.path:
    _sys.path.insert(0, _pkg)
from nbkode_b200.ensembles import lorenz_ensemble, linear_ensemble, rd1d_ensemble, vdp_ensemble  # noqa: E402,F401
