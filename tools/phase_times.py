"""Diagnostic (GPU box): host-side cost of one pass (constructor + run) of the Lorenz benchmark."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
sync = torch.cuda.synchronize
for it in range(5):
    sync(); t0 = time.perf_counter()
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    t1 = time.perf_counter()
    t, y = s.run([10.0])
    t2 = time.perf_counter(); sync(); t3 = time.perf_counter()
    print(f"ctor cpu {1e3*(t1-t0):.3f} ms | run cpu (incl. its syncs) {1e3*(t2-t1):.3f} ms | tail sync {1e3*(t3-t2):.3f} ms")
pr = cProfile.Profile(); sync(); pr.enable()
for _ in range(5):
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9); s.run([10.0])
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(18)
