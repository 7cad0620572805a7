"""Diagnostic (GPU box): counters of the K1 tail compaction (library built with -DNBK_DEBUG_COMPACT, see NBK_LIB_PATH)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

for B in (125_000, 1_000_000):
    y0, p = ensembles.lorenz_ensemble(B, 0)
    y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
    for it in range(3):
        s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run([10.0]); e1.record(); torch.cuda.synchronize()
    q = s._queue.cpu().tolist()
    if len(q) > 8:
        import numpy as np
        nw = s.last_launch()["grid"] * 4
        t0 = q[6]
        ex = (np.array(q[8:8 + nw]) - t0) / 1e6
        qe = (np.array(q[8 + 4096:8 + 4096 + nw]) - t0) / 1e6
        reps = np.array(q[8 + 8192:8 + 8192 + nw], dtype=float)
        ok = (qe > 0) & (reps > 16)
        us = 1e3 * (ex[ok] - qe[ok]) / reps[ok]
        late = ok & (ex >= np.percentile(ex, 90))
        print("  tail phase: us per attempt of a warp, percentiles", [0, 5, 25, 50, 75, 95, 100], ":",
              np.round(np.percentile(us, [0, 5, 25, 50, 75, 95, 100]), 3).tolist(),
              "| the 10 % of warps that exit last:", np.round(np.percentile(1e3 * (ex[late] - qe[late]) / reps[late], [0, 50, 100]), 3).tolist(),
              "their tail attempts:", np.percentile(reps[late], [0, 50, 100]).tolist())
        qe = qe[qe > 0]
        pc = [0, 1, 5, 25, 50, 75, 95, 99, 100]
        print("  warp exit time (ms after the first CTA started), percentiles", pc, ":", np.round(np.percentile(ex, pc), 3).tolist())
        print("  queue seen empty (ms), percentiles:", np.round(np.percentile(qe, pc), 3).tolist(), "warps that saw it:", len(qe))
    print(f"B={B}: {e0.elapsed_time(e1):.3f} ms; queue head {q[0]}, failures {q[1]}, lock acquisitions {q[2]}, taken {q[3]}, "
          f"parked {q[4]}, warps retired in the tail phase {q[5]} (of {s.last_launch()['grid'] * 4})")
