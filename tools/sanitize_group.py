"""Small cases of the sub-warp groups / teams and of run_events for compute-sanitizer (memcheck / racecheck), GPU box:
    compute-sanitizer --tool racecheck python tools/sanitize_group.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

rng = np.random.default_rng(0)


def crossing(t, y):
    return y[0] - 0.5


def gap(t, y):
    return y[2] - y[4]


gap.terminal = True
for cls, n, kw in (("RungeKutta45", 13, dict(rtol=1e-6, atol=1e-9)), ("RungeKutta23", 32, dict(rtol=1e-5, atol=1e-8)),
                   ("RungeKutta45", 100, dict(rtol=1e-6, atol=1e-9)), ("DOP853", 20, dict(rtol=1e-6, atol=1e-9)),
                   ("AdamsBashforth4", 24, dict(h=0.002)), ("RungeKutta4", 50, dict(h=0.004)), ("RungeKutta45", 300, dict(rtol=1e-6, atol=1e-9))):
    B = 200
    y0, p = ensembles.rd1d_ensemble(B, n=n, seed=3)
    p[:, 0] *= 0.3 * (100.0 / n) ** 2
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
    t, y = s.run([0.1, 0.3])
    s.step(n=2)
    s.run([float(np.max(s.t)) + 0.2])
    msg = f"{cls} n={n} lanes={s._group} ok"
    if cls in ("RungeKutta45", "RungeKutta23", "DOP853"):
        e = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
        out = e.run_events([0.5, 1.5], [crossing, gap])
        msg += f", events {e.event_counts.sum()}"
    print(msg, flush=True)
yd, kd = rng.uniform(0.5, 2, (300, 16)), rng.uniform(0.1, 1.0, (300, 16))
s = nbk.RungeKutta45("decay", 0.0, yd, kd, rtol=1e-6, atol=1e-9)
s.run([1.0, 5.0])
print("decay n=16 lanes", s._group, "ok")
