"""Diagnostic (GPU box): shape scan of the sub-warp regime -- lanes per trajectory (NBK_GROUP_G), resident blocks the
kernels are compiled for (NBK_GRP_MINB), attempts per queue check (NBK_GRP_INNER).  usage: tune_group.py [rhs] [B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

rhs = sys.argv[1] if len(sys.argv) > 1 else "decay"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
peak, _ = nbk.fp64_peak(0, 20000)
rng = np.random.default_rng(0)
for n, shapes in ((16, (4, 2, 8)), (24, (8, 4)), (32, (8, 4, 16)), (64, (16, 8, 32))):
    if rhs == "decay":
        y0, p, flop = rng.uniform(0.5, 2, (B, n)), rng.uniform(0.1, 1.0, (B, n)), 72 * n + 18
    else:
        y0, p = ensembles.rd1d_ensemble(B, n=n, seed=3)
        flop = 6 * 8 * n + 66 * n + 18
    y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
    for G in shapes:
        for minb in (2, 3, 4, 5):
            for inner in (4, 8):
                E = -(-n // G)
                if (E >= 8 and minb > 3) or (E <= 2 and minb < 4) or (inner == 8 and minb not in (3, 4)):
                    continue
                os.environ.update(NBK_GROUP_G=str(G), NBK_GRP_MINB=str(minb), NBK_GRP_INNER=str(inner))
                try:
                    best = None
                    for _ in range(3):
                        s = nbk.RungeKutta45(rhs, 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); s.run([10.0]); e1.record(); torch.cuda.synchronize()
                        ms = e0.elapsed_time(e1)
                        best = ms if best is None else min(best, ms)
                    att = int((s._nacc.sum() + s._nrej.sum()).item())
                    tf = att * flop / best * 1e3 / 1e12
                    print(json.dumps({"rhs": rhs, "n": n, "G": G, "E": E, "minb": minb, "inner": inner, "ms": round(best, 3),
                                      "frac": round(tf / peak, 3), "launch": s.last_launch()}), flush=True)
                except Exception as exc:  # noqa: BLE001
                    print(json.dumps({"n": n, "G": G, "minb": minb, "error": str(exc)[:200]}), flush=True)
