"""nbkode_b200 -- B200-native batched ODE-ensemble engine behind the nbkode solver API.

Drop-in for the explicit hot path of hgrecco/numbakit-ode: ``ForwardEuler``,
``AdamsBashforth1..5``, ``RungeKutta23``, ``RungeKutta45``, ``get_solvers()`` & co.
(nbkode/__init__.py:15-75), with ``y0``/``params`` optionally carrying a leading batch axis.
All arithmetic runs in hand-written sm_100a CUDA kernels (libnbkode_b200.so) reached through
a ctypes C-ABI; there is no CPU fallback.
"""

from . import multistep, runge_kutta  # noqa: F401  (importing registers the solver classes)
from .core import Solver, get_groups, get_solver, get_solvers, list_solvers
from .rhs import CudaEvents, CudaRHS
from ._lib import CompileError, NbkError, fp64_peak

__version__ = "0.1.0"

_SOLVERS_NAMES = None


def test():  # pragma: no cover
    """Run the repository test-suite (counterpart of nbkode.test())."""
    import os

    import pytest

    return pytest.main([os.path.join(os.path.dirname(__file__), "..", "..", "tests"), "-q"])


def __dir__():
    global _SOLVERS_NAMES
    if not _SOLVERS_NAMES:
        names = ["get_groups", "get_solvers", "get_solver", "test"]
        names += list_solvers(include_alias=False)
        _SOLVERS_NAMES = sorted(names)
    return _SOLVERS_NAMES


def __getattr__(name):
    try:
        solver = get_solver(name)
    except ValueError:
        raise AttributeError(f"module {__name__} has no attribute {name}")
    if solver.__name__ != name:
        from warnings import warn

        warn(
            f"The name '{name}' is an alias and its use is discouraged. Use '{solver.__name__}' instead.",
            stacklevel=2,
        )
    return solver
