"""One configuration, two passes (warm-up + the one to capture) -- the command that is run plain and then under ncu:
    python tools/prof_one.py <case> && ncu --set full ... -k regex:nbk_run -s 1 -c 1 python tools/prof_one.py <case>
cases (full sizes of BASELINE.json unless noted):
  k1        C2   Lorenz 1M, RungeKutta45, run([10])
  k1_125k   C2 shard of an 8-GPU run: the first 125 000 trajectories
  k1_rk23   C3   Van der Pol 4M x 1000-point grid, RungeKutta23
  k2        C3   the same grid, AdamsBashforth4 h = 0.01
  k3        C5   reaction-diffusion n = 2048 x 65 536, RungeKutta45, t in [0, 1]
  k4        C4   linear n = 128 x 262 144, RungeKutta45 (DMMA)
  kg        mid-size: decay n = 32 x 262 144, RungeKutta45 (sub-warp regime)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

case = sys.argv[1]
kw45 = dict(rtol=1e-6, atol=1e-9)
if case in ("k1", "k1_125k"):
    y0, p = ensembles.lorenz_ensemble(1_000_000, seed=0)
    if case == "k1_125k":
        y0, p = y0[:125_000], p[:125_000]
    mk, ts = (lambda a, b: nbk.RungeKutta45("lorenz", 0.0, a, b, **kw45)), [10.0]
elif case in ("k1_rk23", "k2"):
    y0, p = ensembles.vdp_ensemble(4_000_000, seed=1)
    ts = np.linspace(0, 20, 1000)
    mk = (lambda a, b: nbk.RungeKutta23("vdp", 0.0, a, b, **kw45)) if case == "k1_rk23" else (lambda a, b: nbk.AdamsBashforth4("vdp", 0.0, a, b, h=0.01))
elif case == "k3":
    y0, p = ensembles.rd1d_ensemble(65_536, n=2048, seed=3)
    mk, ts = (lambda a, b: nbk.RungeKutta45("rd1d", 0.0, a, b, **kw45)), [1.0]
elif case == "k4":
    y0, p = ensembles.linear_ensemble(262_144, n=128, seed=2)
    mk, ts = (lambda a, b: nbk.RungeKutta45("linear", 0.0, a, b, **kw45)), [10.0]
elif case == "kg":
    rng = np.random.default_rng(0)
    y0, p = rng.uniform(0.5, 2, (262_144, 32)), rng.uniform(0.1, 1.0, (262_144, 32))
    mk, ts = (lambda a, b: nbk.RungeKutta45("decay", 0.0, a, b, **kw45)), [10.0]
else:
    raise SystemExit(__doc__)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
for it in range(2):
    s = mk(y0d, pd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.run(ts); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    del out
s.raise_for_status()
print(json.dumps({"case": case, "ms": ms, "attempts": int((s._nacc.sum() + s._nrej.sum()).item()), "launch": s.last_launch()}))
