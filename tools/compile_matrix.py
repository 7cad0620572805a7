"""CPU-only diagnostic: NVRTC-compile (no GPU needed) the matrix of explicit solver classes x regimes (per thread n = 3,
sub-warp groups / teams n = 24 / 100, one CTA per trajectory n = 300 / 2048) x parameter blocks (2, none, 20 parameters)
x {run, run_events} from traced Python right-hand sides and events.  Include-order and macro mistakes in the JIT source
assembly (csrc/nbk_jit.cpp) show up here without a GPU.  About six minutes with a cold NVRTC cache.

    python tools/compile_matrix.py
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]

import numpy as np  # noqa: E402

from nbkode_b200 import _lib, rhs  # noqa: E402


def rd(t, y, p):
    return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)


def rd0(t, y):
    return 0.1 * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + y * (1 - y)


def rdbig(t, y, p):
    return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[19] * y * (1 - y)


METHODS = (45, 23, 853, 1, 2, 3, 4, 5, 202, 203, 213, 204, 238)  # ids of include/nbkode_b200.h


def main():
    lib = _lib.lib()
    sz = C.c_longlong(0)
    fails, count, t0 = [], 0, time.time()
    for method in METHODS:
        for n in (3, 24, 100, 300, 2048):
            for f, n_par in ((rd, 2), (rd0, 0), (rdbig, 20)):
                tracer = rhs.trace if n == 3 else rhs.trace_cta
                src = tracer(f, n, (n_par,), n_par > 0).source.encode()
                cfg = _lib.NbkConfig(method, n, n_par, 0, 16, None, src, 1e-6, 1e-9, 0.2, 10.0, 0.9, 0, 0)
                for n_events in (0, 2):
                    try:
                        if n_events:
                            ev = rhs.trace_events([lambda t, y: y[0] - 0.5, lambda t, y: (y[1] - y[2] if t < 1 else y[1])],
                                                  n, (n_par,), n_par > 0)
                            _lib.check(lib.nbk_jit_compile_events_only(C.byref(cfg), 2, ev.source.encode(), C.byref(sz)))
                        else:
                            _lib.check(lib.nbk_jit_compile_only(C.byref(cfg), C.byref(sz)))
                        count += 1
                    except Exception as exc:  # noqa: BLE001
                        fails.append((method, n, n_par, n_events, str(exc)[:300]))
                        print("FAIL", *fails[-1], flush=True)
    print(f"{count} programs compiled, {len(fails)} failed, {time.time() - t0:.0f} s")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
