"""Diagnostic (GPU box): kernel time of the Lorenz RK45 pass for the library in NBK_LIB_PATH."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp

B = int(os.environ.get("NBK_TUNE_B", 1_000_000)); reps = int(os.environ.get("NBK_TUNE_REPS", 6))
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
times, att = [], 0
for it in range(reps + 3):
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.run([10.0]); e1.record(); torch.cuda.synchronize()
    if it >= 3: times.append(e0.elapsed_time(e1))
    att = int((s._nacc.sum() + s._nrej.sum()).item()); acc = int(s._nacc.sum().item())
ms = float(np.median(times))
print(json.dumps({"lib": os.path.basename(os.environ.get("NBK_LIB_PATH", "default")), "ms_median": ms, "ms_min": min(times),
                  "attempts_per_s": att / ms * 1e3, "accepted_per_s": acc / ms * 1e3, "tflops": att * 264 / ms * 1e3 / 1e12,
                  "launch": s.last_launch()}))
