"""Diagnostic (GPU box): kernel time of one Lorenz ensemble pass (run([10.0])) for the library in NBK_LIB_PATH.
env: NBK_TUNE_METHOD (RungeKutta45 | RungeKutta23 | DOP853), NBK_TUNE_B, NBK_TUNE_REPS, NBK_TUNE_RTOL/ATOL.
FLOP per attempt (SURVEY.md 8(d) convention, FMA = 2): RK45 6R+66n+18 = 264, RK23 3R+27n+18 = 123 (R = 8, n = 3),
DOP853 12R + n(2 nnz(A) + 11) + 2n nnz(B) + n(2 nnz(E5) + 2 nnz(E3) + 12) + 26 = 96 + 3*111 + 48 + 3*44 + 26 = 635."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles as onp  # input generators

method = os.environ.get("NBK_TUNE_METHOD", "RungeKutta45")
FLOP = {"RungeKutta45": 264, "RungeKutta23": 123, "DOP853": 635}[method]
B = int(os.environ.get("NBK_TUNE_B", 1_000_000)); reps = int(os.environ.get("NBK_TUNE_REPS", 6))
rtol, atol = float(os.environ.get("NBK_TUNE_RTOL", 1e-6)), float(os.environ.get("NBK_TUNE_ATOL", 1e-9))
T_END = float(os.environ.get("NBK_TUNE_T", 10.0))  # a short horizon = short trajectories (few attempts each)
y0, p = onp.lorenz_ensemble(B, 0)


def rhs_lorenz(t, y, p):
    return np.array([p[0] * (y[1] - y[0]), y[0] * (p[1] - y[2]) - y[1], y[0] * y[1] - p[2] * y[2]])


y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
times, att = [], 0
for it in range(reps + 3):
    # NBK_TUNE_JIT=1: the same right-hand side as a Python callable (traced -> CUDA C -> NVRTC) instead of the catalogue kernel
    s = getattr(nbk, method)(rhs_lorenz if os.environ.get("NBK_TUNE_JIT") else "lorenz", 0.0, y0d, pd, rtol=rtol, atol=atol)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.run([T_END]); e1.record(); torch.cuda.synchronize()
    if it >= 3: times.append(e0.elapsed_time(e1))
    att = int((s._nacc.sum() + s._nrej.sum()).item()); acc = int(s._nacc.sum().item())
ms = float(np.median(times))
print(json.dumps({"lib": os.path.basename(os.environ.get("NBK_LIB_PATH", "default")), "method": method, "rtol": rtol, "t_end": T_END,
                  "ms_median": ms, "ms_min": min(times), "attempts": att, "accepted": acc,
                  "attempts_per_s": att / ms * 1e3, "accepted_per_s": acc / ms * 1e3, "tflops": att * FLOP / ms * 1e3 / 1e12,
                  "launch": s.last_launch()}))
