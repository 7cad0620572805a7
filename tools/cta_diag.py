import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_c as oc, nbk_oracle_np as onp  # diagnostic tool: compares against the oracle on purpose
for cls, n, B in (("RungeKutta45", 2048, 48), ("RungeKutta45", 100, 40), ("RungeKutta23", 2048, 8)):
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    ts = [0.25, 0.5]
    ref = oc.run_ensemble(cls, "rd1d", y0, p, ts, rtol=1e-6, atol=1e-9, nthreads=8)
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t, y = s.run(ts)
    tol = 1e-6 * np.abs(ref["ys"]) + 1e-9
    ratio = (np.abs(y - ref["ys"]) / tol).max(axis=(1, 2))
    dn = s.n_accepted - ref["n_accepted"]
    print(cls, n, "n_acc ref", ref["n_accepted"][:8], "dn", dn[:16], "rej", s.n_rejected[:8], ref["n_rejected"][:8])
    print("   max |dy|/tol per traj:", np.round(ratio[:16], 3), "overall", ratio.max())
    # first-step check: one step only
    s2 = getattr(nbk, cls)("rd1d", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    t1, y1 = s2.step(n=3)
    o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[0], p[0], rtol=1e-6, atol=1e-9)
    tr, yr = onp.steps(o, n=3)
    print("   3 steps traj0: t", t1[0], tr, "max|dy|", np.abs(y1[0] - yr).max(), "h", s2.h[0], float(o.h))
