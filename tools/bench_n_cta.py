"""Diagnostic (GPU box): the same decay ensembles as bench_n.py through the CTA regime (one small CTA per trajectory,
component-wise right-hand side) -- where does it overtake the per-thread regime?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
src = nbk.CudaRHS("NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n) { return -p[i] * y[i]; }")
for n in [int(x) for x in os.environ.get('NBK_NS', '8,12,16,24,32,48,64').split(',')]:
    y0 = torch.from_numpy(rng.uniform(0.5, 2, (B, n))).cuda()
    p = torch.from_numpy(rng.uniform(0.1, 1.0, (B, n))).cuda()
    times = []
    for it in range(3):
        s = nbk.RungeKutta45(src, 0.0, y0, p, rtol=1e-6, atol=1e-9)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run([10.0]); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    att = int((s._nacc.sum() + s._nrej.sum()).item())
    ms = min(times[1:])
    print(json.dumps({"n": n, "ms": ms, "attempts": att, "attempts_per_s": att / ms * 1e3, "regime": s._regime, "launch": s.last_launch()}), flush=True)
