"""Diagnostic (GPU box, under ncu): one RungeKutta23 Van der Pol pass with dense output (configs[2] shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles as onp  # input generators
B = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cls = sys.argv[3] if len(sys.argv) > 3 else "RungeKutta23"
y0, p = onp.vdp_ensemble(B, seed=1)
kw = dict(h=0.01) if getattr(nbk, cls).FIXED_STEP else dict(rtol=1e-6, atol=1e-9)
s = getattr(nbk, cls)("vdp", 0.0, torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda(), **kw)
t, y = s.run(np.linspace(0, 20, m))
torch.cuda.synchronize()
print(int(s._nacc.sum()), int(s._nrej.sum()), s.last_launch())
