"""Diagnostic (GPU box): the classes without a dedicated sub-warp kernel (DOP853, Adams-Bashforth, fixed-step Runge-Kutta)
on mid-size systems -- sub-warp teams (NBK_TEAM_G) against one CTA per trajectory (NBK_NO_TEAM=1); rd1d stencil.
usage: bench_team.py [B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
for cls, kw, T in (("AdamsBashforth4", dict(h=0.002), 1.0), ("RungeKutta4", dict(h=0.004), 1.0), ("DOP853", dict(rtol=1e-8, atol=1e-11), 2.0)):
    for n in (16, 32, 64, 100):
        y0, p = ensembles.rd1d_ensemble(B, n=n, seed=3)
        p[:, 0] *= 0.05 * (100.0 / n) ** 2
        y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
        row = {"cls": cls, "n": n, "B": B}
        for label, env in (("team", None), ("cta", "1")):
            if env:
                os.environ["NBK_NO_TEAM"] = env
            else:
                os.environ.pop("NBK_NO_TEAM", None)
            best = None
            for _ in range(3):
                s = getattr(nbk, cls)("rd1d", 0.0, y0d, pd, **kw)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.run([T]); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            row[label] = {"ms": round(best, 3), "launch": s.last_launch()}
        os.environ.pop("NBK_NO_TEAM", None)
        row["speedup"] = round(row["cta"]["ms"] / row["team"]["ms"], 2)
        print(json.dumps(row), flush=True)
