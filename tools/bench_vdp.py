"""Secondary benchmark (GPU box): BASELINE.json configs[2] -- Van der Pol ensemble, mu in [0.1, 5],
RungeKutta23 (rtol=1e-6, atol=1e-9) and AdamsBashforth4 (h=0.01), dense output on a 1000-point grid over
t in [0, 20].  At the full 4M trajectories the output alone is 64 GB, so HBM write bandwidth co-bounds
the run.  Reports steps/s, output GB/s against MEASURED_PEAKS.json's copy bandwidth."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles as onp  # input generators

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0
y0, p = onp.vdp_ensemble(B, seed=1)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
ts = np.linspace(0, 20, m)
CASES = [("RungeKutta23", dict(rtol=1e-6, atol=1e-9)), ("AdamsBashforth4", dict(h=0.01))]
if len(sys.argv) > 3:  # e.g. "RungeKutta4": only that fixed-step class, h = 0.01
    CASES = [(sys.argv[3], dict(h=0.01))]
for cls, kw in CASES:
    times = []
    for it in range(6):
        s = getattr(nbk, cls)("vdp", 0.0, y0d, pd, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t, y = s.run(ts); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        steps = int(s._nacc.sum().item())
        del y
    ms = min(times)
    out_gb = B * m * 2 * 8 / 1e9
    print(json.dumps({"solver": cls, "B": B, "m": m, "ms": ms, "times": times, "steps": steps, "steps_per_s": steps / ms * 1e3,
                      "output_GB": out_gb, "output_GBps": out_gb / ms * 1e3, "hbm_copy_GBps_measured": hbm,
                      "frac_of_hbm_write": out_gb / ms * 1e3 / hbm, "launch": s.last_launch()}), flush=True)
