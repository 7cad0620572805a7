"""Secondary benchmark (GPU box): BASELINE.json configs[4] -- 1-D reaction-diffusion method of lines,
n = 2048 per trajectory, RungeKutta45 rtol=1e-6 atol=1e-9, t in [0, 1], one CTA per trajectory (K3) --
and configs[3] -- linear y' = A y, n = 128, shared dense A (generic K3 path; the DMMA kernel K4 is not built yet).
Reports attempted steps/s and the FP64 roofline fraction (6R + 66n + 18 FLOP per attempt)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles as onp  # input generators

which = sys.argv[1] if len(sys.argv) > 1 else "rd1d"
cls = sys.argv[4] if len(sys.argv) > 4 else "RungeKutta45"  # a fixed-step class runs rd1d with h = 5e-4 to T = 0.1
fixed = getattr(nbk, cls).FIXED_STEP
B = int(sys.argv[2]) if len(sys.argv) > 2 else (16384 if which == "rd1d" else 65536)
peak, _ = nbk.fp64_peak(0, 20000)
traced = which == "rd1d_py"  # the same right-hand side as a NumPy-style Python callable (vector tracer -> nbk_rhs_i -> NVRTC)
if traced:
    which = "rd1d"


def rd1d_py(t, y, p):
    return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)


if which == "rd1d":
    n, T = (int(sys.argv[3]) if len(sys.argv) > 3 else 2048), 1.0
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    p[:, 0] *= (n / 2048.0) ** 2  # same h*d (stability-limited step count) on coarser grids
    flop = 6 * 8 * n + 66 * n + 18
    mk = lambda: nbk.RungeKutta45(rd1d_py if traced else "rd1d", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    if cls != "RungeKutta45":
        if fixed:
            T = 0.1
            stages = {"RungeKutta4": 4, "RungeKutta3_8": 4, "Runge3": 4, "Heun3": 3, "Runge2": 2}.get(cls, 1)
            k = int(cls[-1]) if cls.startswith("AdamsBashforth") else 1
            # R + 2 k n + k + 1 (Adams-Bashforth-k) / S R + (stage arguments + y_new) (Runge-Kutta), SURVEY.md 8(d)
            flop = 8 * n + 2 * k * n + k + 1 if stages == 1 else stages * 8 * n + (stages * (stages - 1) + 2 * stages) * n
            mk = lambda: getattr(nbk, cls)("rd1d", 0.0, y0d, pd, h=5e-4)
        else:
            S = {"RungeKutta23": 3, "DOP853": 12}[cls]
            flop = S * 8 * n + (27 if S == 3 else 171) * n + 18
            mk = lambda: getattr(nbk, cls)("rd1d", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
else:
    n, T = 128, 10.0
    y0, p = onp.linear_ensemble(B, n=n, seed=2)
    flop = 6 * 2 * n * n + 66 * n + 18
    mk = lambda: nbk.RungeKutta45("linear", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
times = []
for it in range(4):
    s = mk()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.run([T]); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
s.raise_for_status()
att = int((s._nacc.sum() + s._nrej.sum()).item()); acc = int(s._nacc.sum().item())
ms = min(times[1:])
print(json.dumps({"config": which, "B": B, "n": n, "ms": ms, "times": times, "accepted_steps": acc, "attempts": att,
                  "accepted_steps_per_s": acc / ms * 1e3, "flop_per_attempt": flop, "tflops": att * flop / ms * 1e3 / 1e12,
                  "peak_tflops": peak, "frac": att * flop / ms * 1e3 / 1e12 / peak, "launch": s.last_launch()}))
