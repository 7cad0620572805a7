"""Explicit Runge-Kutta solvers: fixed-step Runge2..RungeKutta3_8, adaptive RungeKutta23, RungeKutta45, DOP853.

Class attributes mirror nbkode/runge_kutta/explicit.py:132-253 (tableaux ``A, B, C, E, P``,
``error_estimator_order``, ``stages``); the arithmetic runs in the CUDA kernels of
csrc/nbk_device.cuh, which carry their own compile-time copy of these tableaux
(tests/test_host_api.py checks that the two agree).
"""

from __future__ import annotations

import numpy as np

from .core import Solver
from .util import classproperty

VARIABLE_STEP_OPTIONS = ("atol", "rtol", "min_step", "max_step", "min_factor", "max_factor", "safety_factor")


class VariableStepOptions:
    """nbkode/core.py:661-679.  ``min_step``/``max_step`` are stored but, exactly like the
    reference, never enforced."""

    def __init__(self, atol=1e-6, rtol=1e-3, min_step=1e-15, max_step=np.inf, min_factor=0.2, max_factor=10.0,
                 safety_factor=0.9):
        self.atol, self.rtol = float(atol), float(rtol)
        self.min_step, self.max_step = float(min_step), float(max_step)
        self.min_factor, self.max_factor, self.safety_factor = float(min_factor), float(max_factor), float(safety_factor)

    def as_dict(self):
        return {k: getattr(self, k) for k in VARIABLE_STEP_OPTIONS}


class RungeKutta(Solver, abstract=True):
    GROUP = "Runge-Kutta"
    FIXED_STEP = True
    LEN_HISTORY = 2
    IMPLICIT = False

    A: np.ndarray
    B: np.ndarray
    C: np.ndarray

    @property
    def K(self):
        """Stage derivatives of the last accepted step, (S+1, n) or (B, S+1, n)."""
        full = self._K.permute(2, 0, 1) if self._regime == 0 else self._K
        return self._out(full[0] if not self._batched else full)


class ERK(RungeKutta, abstract=True):
    """Fixed-step explicit Runge-Kutta (nbkode/runge_kutta/explicit.py:9-46, 102-129): default h = 1,
    dense output by the Hermite cubic over the last step."""

    # all trajectories advance in lock-step: dense output written as [m][n][B] (coalesced), like the multistep kernels
    OUT_LAYOUT = "mnb"

    @classproperty
    def stages(cls):
        return len(cls.A)

    @classproperty
    def STAGE_ROWS(cls):
        return len(cls.A)


class Runge2(ERK):
    METHOD_ID = 202
    A = np.array([[0, 0], [1 / 2, 0]])
    B = np.array([0, 1.0])
    C = np.array([0, 1 / 2])


class Runge3(ERK):
    METHOD_ID = 203
    A = np.array([[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]])
    B = np.array([1, 4, 0, 1]) / 6
    C = np.array([0, 1 / 2, 1, 1])


class Heun3(ERK):
    METHOD_ID = 213
    A = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0]]) / 3
    B = np.array([1, 0, 3]) / 4
    C = np.array([0, 1, 2]) / 3


class RungeKutta4(ERK):
    METHOD_ID = 204
    A = np.array([[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 1 / 2, 0, 0], [0, 0, 1, 0]])
    B = np.array([1, 2, 2, 1]) / 6
    C = np.array([0, 1, 1, 2]) / 2


class RungeKutta3_8(ERK):
    METHOD_ID = 238
    A = np.array([[0, 0, 0, 0], [1 / 3, 0, 0, 0], [-1 / 3, 1, 0, 0], [1, -1, 1, 0]])
    B = np.array([1, 3, 3, 1]) / 8
    C = np.array([0, 1, 2, 3]) / 3


class AdaptiveRungeKutta(RungeKutta, abstract=True):
    """FSAL embedded pairs with the reference's I-controller
    (nbkode/runge_kutta/core.py:56-128, explicit.py:49-99)."""

    FIXED_STEP = False
    error_estimator_order: int

    def __init__(self, rhs, t0, y0, params=None, **kwargs):
        opts = {k: kwargs.pop(k) for k in VARIABLE_STEP_OPTIONS if k in kwargs}
        self.options = VariableStepOptions(**opts)
        if self.options.max_step <= 0:
            raise ValueError("`max_step` must be positive.")
        if self.options.rtol < 0 or self.options.atol < 0:
            raise ValueError("`rtol` and `atol` must be non-negative.")
        first_step = kwargs.pop("first_step", None)
        kwargs.pop("h", None)  # the reference overwrites h with select_initial_step (core.py:695-706)
        super().__init__(rhs, t0, y0, params, _options=self.options.as_dict(), _first_step=first_step, **kwargs)

    @property
    def h(self):
        """Next step size: 0-d array (unbatched, like the reference) or (B,)."""
        return np.array(float(self._h[0].item())) if not self._batched else self._out(self._h)

    @classproperty
    def stages(cls):
        return len(cls.A) + 1

    @classproperty
    def error_exponent(cls):
        return 1 / (cls.error_estimator_order + 1)


class RungeKutta23(AdaptiveRungeKutta):
    """Bogacki-Shampine 3(2) pair with cubic Hermite-type dense output."""

    METHOD_ID = 23
    STAGE_ROWS = 4
    C = np.array([0, 1 / 2, 3 / 4], dtype=float)
    A = np.array([[0, 0, 0], [1 / 2, 0, 0], [0, 3 / 4, 0]], dtype=float)
    B = np.array([2 / 9, 1 / 3, 4 / 9], dtype=float)
    B2 = np.array([7 / 24, 1 / 4, 1 / 3, 1 / 8], dtype=float)
    E = np.array([2 / 9 - 7 / 24, 1 / 3 - 1 / 4, 4 / 9 - 1 / 3, -1 / 8])
    P = np.array([[1, -4 / 3, 5 / 9], [0, 1, -2 / 3], [0, 4 / 3, -8 / 9], [0, -1, 1]])
    error_estimator_order = 2


class RungeKutta45(AdaptiveRungeKutta):
    """Dormand-Prince 5(4) pair with Shampine's quartic dense output."""

    METHOD_ID = 45
    STAGE_ROWS = 7
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
    C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
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
    error_estimator_order = 4


class DOP853(AdaptiveRungeKutta):
    """Hairer's explicit 8(5,3) pair (nbkode/runge_kutta/explicit.py:282-351): 12 stages + FSAL,
    error norm ``err5 / sqrt(1 + (err3 / (10 err5))**2)``, 7th-order dense output that costs three
    extra right-hand-side evaluations per interpolated step (``DOP853Interpolator``, :406-481)."""

    from . import dop853_tableau as _tab

    METHOD_ID = 853
    STAGE_ROWS = 13
    n_stages = _tab.N_STAGES
    order = 8
    error_estimator_order = 7
    A = _tab.A[:12, :12].copy()
    B = _tab.B
    C = _tab.C[:12]
    E3 = _tab.E3
    E5 = _tab.E5
    D = _tab.D
    A_EXTRA = _tab.A[13:]
    C_EXTRA = _tab.C[13:]
