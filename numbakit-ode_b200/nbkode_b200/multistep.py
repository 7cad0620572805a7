"""Fixed-step multistep solvers: AdamsBashforth1..5 / ForwardEuler (explicit) and AdamsMoulton1..5 /
BackwardEuler / BDF1..6 (implicit).

Mirrors nbkode/multistep/core.py:12-112, nbkode/multistep/adams_bashforth.py and
nbkode/euler.py:19-27, including the behaviours SURVEY.md section 0 lists:
``B[0]`` multiplies the OLDEST derivative of the window, the default start-up is an
adaptive RungeKutta45 with default tolerances evaluated at the absolute times ``h, 2h, ...``,
and dense output is a cubic Hermite across the whole history window.

Deliberate deviation: ``IMPLICIT`` is ``False`` for these explicit methods.  The reference
reports ``True`` because ``Multistep.IMPLICIT`` tests ``Bn == 0`` (multistep/core.py:29-31),
which hides them from ``get_solvers(implicit=False)``.
"""

from __future__ import annotations

import numpy as np

from .core import Solver
from .util import classproperty

_FIRST_STEPPER_IDS = {
    "rungekutta45": 45, "rungekutta23": 23, "forwardeuler": 1, "euler": 1,
    "adamsbashforth1": 1, "adamsbashforth2": 2, "adamsbashforth3": 3, "adamsbashforth4": 4, "adamsbashforth5": 5,
}


class Multistep(Solver, abstract=True):
    GROUP = "Multistep"
    FIXED_STEP = True
    IMPLICIT = False
    FIRST_STEPPER_CLS = "RungeKutta45"
    # fixed-step kernels advance all trajectories in lock-step: write dense output as [m][n][B]
    # (coalesced, measured 2.3x faster than (B, m, n) on the 1M x 1000-point Van der Pol config)
    OUT_LAYOUT = "mnb"

    A: np.ndarray
    B: np.ndarray

    @classproperty
    def ORDER(cls):
        return cls.A.size

    def __init__(self, rhs, t0, y0, params=None, *, first_stepper_cls="auto", **kwargs):
        fs_id = 0
        if first_stepper_cls is not None and self.ORDER > 1:
            if first_stepper_cls == "auto":
                first_stepper_cls = self.FIRST_STEPPER_CLS
            name = first_stepper_cls if isinstance(first_stepper_cls, str) else first_stepper_cls.__name__
            try:
                fs_id = _FIRST_STEPPER_IDS[name.lower()]
            except KeyError:
                raise ValueError(f"first_stepper_cls {name!r} is not an explicit solver of nbkode_b200") from None
            h = kwargs.get("h")
            h_min = 1.0 if h is None else float(np.min(h.detach().cpu().numpy() if hasattr(h, "detach") else h))
            if fs_id in (23, 45) and float(t0) > h_min:
                # the reference evaluates the start-up at the ABSOLUTE times h, 2h, ... and its run()
                # raises for points before t0 (multistep/core.py:63, core.py:393-397)
                raise ValueError(
                    f"Cannot interpolate at t={h_min} as it is smaller than the current smallest value in history ({float(t0)})"
                )
        super().__init__(rhs, t0, y0, params, _first_stepper=fs_id, **kwargs)


class ExplicitMultistep(Multistep, abstract=True):
    Bn = 0


class AdamsBashforth(ExplicitMultistep, abstract=True):
    GROUP = "Adams-Bashforth"

    @classproperty
    def A(cls):
        A = np.zeros_like(cls.B)
        A[-1] = -1
        return A

    @classproperty
    def LEN_HISTORY(cls):
        return max(cls.B.size, 2)

    @classproperty
    def METHOD_ID(cls):
        return cls.B.size


def _ab(order, weights, doc):
    return type(
        f"AdamsBashforth{order}", (AdamsBashforth,),
        {"B": np.asarray(weights, dtype=float), "LEN_HISTORY": max(order, 2), "METHOD_ID": order, "__doc__": doc,
         "__module__": __name__},
    )


AdamsBashforth1 = _ab(1, [1.0], "Adams-Bashforth with one step (explicit Euler).")
AdamsBashforth2 = _ab(2, np.array([3, -1]) / 2, "Adams-Bashforth, two steps (reference weight order).")
AdamsBashforth3 = _ab(3, np.array([23, -16, 5]) / 12, "Adams-Bashforth, three steps (reference weight order).")
AdamsBashforth4 = _ab(4, np.array([55, -59, 37, -9]) / 24, "Adams-Bashforth, four steps (reference weight order).")
AdamsBashforth5 = _ab(5, np.array([1901, -2774, 2616, -1274, 251]) / 720, "Adams-Bashforth, five steps.")


class ForwardEuler(AdamsBashforth1):
    """y[n+1] = y[n] + h f(t[n], y[n])  (nbkode/euler.py:19-27)."""

    ALIASES = ("Euler",)
    GROUP = "Euler"


class ImplicitMultistep(Multistep, abstract=True):
    """Implicit linear multistep methods (nbkode/multistep/core.py:115-170): every step solves
    ``y - h Bn f(t_new, y) - (h B . fs - A . ys) = 0`` on the device the way the reference does -- Newton with a
    finite-difference Jacobian for n > 1, the secant method for n == 1 (nbkode/nbcompat/zeros.py).  Per-thread
    regime, n <= 8; the kernels are compiled by NVRTC on first use.  ``IMPLICIT`` is ``True`` (the reference's
    flag is inverted, see the module docstring)."""

    IMPLICIT = True

    @classproperty
    def LEN_HISTORY(cls):
        return max(cls.A.size, 2)


class AdamsMoulton(ImplicitMultistep, abstract=True):
    GROUP = "Adams-Moulton"

    @classproperty
    def A(cls):
        A = np.zeros_like(cls.B)
        A[-1] = -1
        return A


class BDF(ImplicitMultistep, abstract=True):
    GROUP = "BDF"

    @classproperty
    def B(cls):
        return np.zeros_like(cls.A)


def _implicit(base, name, method_id, doc, **attrs):
    arr = {k: (np.asarray(v, dtype=float) if k in ("A", "B") else v) for k, v in attrs.items()}
    size = (arr.get("A") if "A" in arr else arr["B"]).size
    return type(name, (base,), {**arr, "LEN_HISTORY": max(size, 2), "METHOD_ID": method_id, "__doc__": doc,
                                "__module__": __name__})


AdamsMoulton1 = _implicit(AdamsMoulton, "AdamsMoulton1", 11, "Adams-Moulton, one step (implicit Euler).", B=[0.0], Bn=1.0)
AdamsMoulton2 = _implicit(AdamsMoulton, "AdamsMoulton2", 12, "Adams-Moulton (trapezoidal rule).", B=[1 / 2], Bn=1 / 2)
AdamsMoulton3 = _implicit(AdamsMoulton, "AdamsMoulton3", 13, "Adams-Moulton, two steps.", B=[-1 / 12, 2 / 3], Bn=5 / 12)
AdamsMoulton4 = _implicit(AdamsMoulton, "AdamsMoulton4", 14, "Adams-Moulton, three steps.", B=np.array([19, -5, 1]) / 24,
                          Bn=9 / 24)
AdamsMoulton5 = _implicit(AdamsMoulton, "AdamsMoulton5", 15, "Adams-Moulton, four steps.",
                          B=np.array([646, -264, 106, -19]) / 720, Bn=251 / 720)


class BackwardEuler(AdamsMoulton1):
    """y[n+1] = y[n] + h f(t[n+1], y[n+1])  (nbkode/euler.py:30-36)."""

    GROUP = "Euler"


BDF1 = _implicit(BDF, "BDF1", 31, "Backward differentiation formula, order 1.", A=[-1.0], Bn=1.0)
BDF2 = _implicit(BDF, "BDF2", 32, "Backward differentiation formula, order 2.", A=np.array([1, -4]) / 3, Bn=2 / 3)
BDF3 = _implicit(BDF, "BDF3", 33, "Backward differentiation formula, order 3.", A=np.array([-2, 9, -18]) / 11,
                 Bn=6 / 11)
BDF4 = _implicit(BDF, "BDF4", 34, "Backward differentiation formula, order 4.", A=np.array([3, -16, 36, -48]) / 25,
                 Bn=12 / 25)
BDF5 = _implicit(BDF, "BDF5", 35, "Backward differentiation formula, order 5.",
                 A=np.array([-12, 75, -200, 300, -300]) / 137, Bn=60 / 137)
BDF6 = _implicit(BDF, "BDF6", 36, "Backward differentiation formula, order 6.",
                 A=np.array([10, -72, 225, -400, 450, -360]) / 147, Bn=60 / 147)
