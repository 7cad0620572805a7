"""ctypes binding of libnbkode_b200.so (include/nbkode_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present,
every compute entry point raises.  Importing the package (registry, class attributes,
RHS tracing) works without a GPU.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBK_LIB_PATH") or os.path.join(HERE, "libnbkode_b200.so")

NBK_OK, NBK_EINVAL, NBK_ECUDA, NBK_ECOMPILE, NBK_ENOTFOUND = 0, -1, -2, -3, -4
TRAJ_OK, TRAJ_BOUND, TRAJ_NONFINITE, TRAJ_MAXATTEMPTS, TRAJ_EVBRACKET, TRAJ_EVENT = 0, 1, 2, 3, 4, -1
MAX_EVENTS = 8


class NbkConfig(C.Structure):
    _fields_ = [
        ("method", C.c_int), ("n", C.c_int), ("n_params", C.c_int), ("params_shared", C.c_int),
        ("batch", C.c_longlong), ("rhs", C.c_char_p), ("rhs_src", C.c_char_p),
        ("rtol", C.c_double), ("atol", C.c_double), ("min_factor", C.c_double),
        ("max_factor", C.c_double), ("safety_factor", C.c_double),
        ("first_stepper", C.c_int), ("device", C.c_int),
    ]


class NbkState(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("ts", "ys", "fs", "K", "h", "params", "n_accepted", "n_rejected", "status", "n_out", "work_counter")]


class NbkEvents(C.Structure):
    _fields_ = [("n_events", C.c_int), ("capacity", C.c_int), ("direction", C.POINTER(C.c_int)),
                ("terminal", C.POINTER(C.c_int)), ("t_events", C.c_void_p), ("y_events", C.c_void_p),
                ("counts", C.c_void_p), ("t_terminal", C.c_void_p)]


class NbkError(RuntimeError):
    pass


class CompileError(NbkError):
    """NVRTC rejected the right-hand side; the message carries the compiler log."""


_lib = None

EXPORTS = (
    "nbk_version", "nbk_last_error", "nbk_create", "nbk_destroy", "nbk_layout", "nbk_is_jit", "nbk_regime", "nbk_group_lanes", "nbk_init",
    "nbk_run", "nbk_step", "nbk_last_launch", "nbk_jit_compile_only", "nbk_fp64_peak", "nbk_fp64_mma_peak",
    "nbk_set_events", "nbk_run_events", "nbk_jit_compile_events_only",
)


def lib():
    """Load the native library (once). Raises ImportError with a build hint if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python numbakit-ode_b200/build.py` "
            "(nbkode_b200 has no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, ll, dbl, i = C.c_void_p, C.c_longlong, C.c_double, C.c_int
    L.nbk_version.restype = i
    L.nbk_last_error.restype = C.c_char_p
    L.nbk_create.argtypes = [C.POINTER(NbkConfig), C.POINTER(vp)]
    L.nbk_destroy.argtypes = [vp]
    L.nbk_destroy.restype = None
    L.nbk_layout.argtypes = [vp, C.POINTER(i), C.POINTER(i)]
    L.nbk_is_jit.argtypes = [vp]
    L.nbk_regime.argtypes = [vp]
    L.nbk_group_lanes.argtypes = [vp]
    L.nbk_init.argtypes = [vp, C.POINTER(NbkState), dbl, vp, ll, ll, vp, ll, ll, dbl, vp, vp]
    L.nbk_run.argtypes = [vp, C.POINTER(NbkState), vp, i, dbl, vp, ll, ll, ll, ll, vp]
    L.nbk_step.argtypes = [vp, C.POINTER(NbkState), ll, dbl, i, vp, ll, vp, ll, ll, ll, ll, vp]
    L.nbk_last_launch.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.nbk_jit_compile_only.argtypes = [C.POINTER(NbkConfig), C.POINTER(ll)]
    L.nbk_fp64_peak.argtypes = [i, i, C.POINTER(dbl), C.POINTER(dbl)]
    L.nbk_fp64_mma_peak.argtypes = [i, i, C.POINTER(dbl), C.POINTER(dbl)]
    L.nbk_set_events.argtypes = [vp, i, C.c_char_p]
    L.nbk_run_events.argtypes = [vp, C.POINTER(NbkState), vp, i, dbl, vp, ll, ll, ll, C.POINTER(NbkEvents), ll, vp]
    L.nbk_jit_compile_events_only.argtypes = [C.POINTER(NbkConfig), i, C.c_char_p, C.POINTER(ll)]
    _lib = L
    return L


def check(rc):
    if rc == NBK_OK:
        return
    msg = lib().nbk_last_error().decode(errors="replace")
    if rc == NBK_EINVAL:
        raise ValueError(msg)
    if rc == NBK_ECOMPILE:
        raise CompileError(msg)
    if rc == NBK_ENOTFOUND:
        raise KeyError(msg)
    raise NbkError(msg)


def fp64_peak(device=0, iters=20000):
    """Measured FP64 DFMA throughput of the device: (TFLOP/s, milliseconds)."""
    tf, ms = C.c_double(0), C.c_double(0)
    check(lib().nbk_fp64_peak(device, iters, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def fp64_mma_peak(device=0, iters=4000):
    """Measured FP64 tensor-pipe (DMMA m8n8k4) throughput: (TFLOP/s, milliseconds)."""
    tf, ms = C.c_double(0), C.c_double(0)
    check(lib().nbk_fp64_mma_peak(device, iters, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value
