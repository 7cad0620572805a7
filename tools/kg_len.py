import json, os, sys
sys.path[:0] = ["/root/repo", "/root/repo/numbakit-ode_b200"]
import numpy as np, torch
import nbkode_b200 as nbk
peak, _ = nbk.fp64_peak(0, 20000)
rng = np.random.default_rng(0)
B, n = 262144, 32
y0 = torch.from_numpy(rng.uniform(0.5, 2, (B, n))).cuda(); p = torch.from_numpy(rng.uniform(0.1, 1.0, (B, n))).cuda()
for T, rtol in ((10.0, 1e-6), (10.0, 1e-9), (40.0, 1e-9)):
    best = None
    for _ in range(3):
        s = nbk.RungeKutta45("decay", 0.0, y0, p, rtol=rtol, atol=rtol * 1e-3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run([T]); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    att = int((s._nacc.sum() + s._nrej.sum()).item())
    print(T, rtol, "attempts/traj", att / B, "ms", round(best, 3), "frac", round(att * (72 * n + 18) / best * 1e3 / 1e12 / peak, 3))
