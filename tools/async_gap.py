"""Diagnostic (GPU box): per-pass wall time of back-to-back async passes (device-resident)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp
B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); cpu = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        c0 = time.perf_counter()
        s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
        c1 = time.perf_counter()
        s.run([10.0])
        cpu.append((1e3 * (c1 - c0), 1e3 * (time.perf_counter() - c1)))
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"10 passes: cpu enqueue {1e3*(t1-t0):.2f} ms, total {1e3*(t2-t0):.2f} ms, events {e0.elapsed_time(e1):.2f} ms")
    print("   per-pass cpu (ctor, run):", " ".join(f"({a:.2f},{b:.2f})" for a, b in cpu))
