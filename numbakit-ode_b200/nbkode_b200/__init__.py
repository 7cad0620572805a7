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
from .sharded import ShardedSolver
from ._lib import CompileError, NbkError, fp64_peak

Solver.register(ShardedSolver)  # isinstance(solver, Solver) holds for the multi-device object, too

__version__ = "0.1.0"

_FUNCTIONS = ("get_groups", "get_solver", "get_solvers", "test")


def __dir__():
    # the four helpers + every registered solver class under its own name (aliases are not listed):
    # what nbkode/__init__.py:51-61 exposes, recomputed on demand (the registry is small)
    return sorted(_FUNCTIONS + tuple(list_solvers(include_alias=False)))


def __getattr__(name):
    # solver classes are reachable as module attributes by name or alias (nbkode/__init__.py:64-75); an alias works
    # but is reported, anything else is an ordinary AttributeError
    import warnings

    try:
        cls = get_solver(name)
    except ValueError:
        raise AttributeError(f"module {__name__} has no attribute {name}") from None
    if cls.__name__ != name:
        warnings.warn(f"The name '{name}' is an alias and its use is discouraged. Use '{cls.__name__}' instead.",
                      stacklevel=2)
    return cls
