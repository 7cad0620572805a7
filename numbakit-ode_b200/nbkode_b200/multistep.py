"""Fixed-step explicit multistep solvers: AdamsBashforth1..5 and ForwardEuler.

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
        return cls.B.size

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
