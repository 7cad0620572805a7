"""Diagnostic (GPU box): per-thread regime as the state size grows (RungeKutta45, element-wise decay, 1 M trajectories):
where register residency ends (7 stage vectors of n doubles per thread)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
B = 1_000_000
rng = np.random.default_rng(0)
for n in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64):
    y0 = torch.from_numpy(rng.uniform(0.5, 2, (B, n))).cuda()
    p = torch.from_numpy(rng.uniform(0.1, 1.0, (B, n))).cuda()
    times = []
    for it in range(4):
        s = nbk.RungeKutta45("decay", 0.0, y0, p, rtol=1e-6, atol=1e-9)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run([10.0]); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    att = int((s._nacc.sum() + s._nrej.sum()).item())
    ms = min(times[1:])
    flop = 6 * n + 66 * n + 18
    print(json.dumps({"n": n, "ms": ms, "attempts": att, "attempts_per_s": att / ms * 1e3, "tflops": att * flop / ms * 1e9 / 1e12 / 1e6 * 1e0,
                      "launch": s.last_launch()}), flush=True)
