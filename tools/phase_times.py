"""Diagnostic (GPU box): where does one pass of the Lorenz benchmark spend its time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
sync = torch.cuda.synchronize

def tick(label, t0):
    sync(); t1 = time.perf_counter(); print(f"  {label:28s} {1e3*(t1-t0):8.3f} ms"); return t1

for it in range(4):
    print("iteration", it)
    sync(); t0 = time.perf_counter()
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    t0 = tick("constructor (incl. init kernel)", t0)
    t, y = s.run([10.0])
    t0 = tick("run([10.0])", t0)
    na = s._nacc.sum().item()
    t0 = tick("n_accepted.sum()", t0)
    del s
    t0 = tick("del solver", t0)
# finer: time pieces of the constructor
import ctypes as C
from nbkode_b200 import _lib
sync(); t0 = time.perf_counter()
bufs = [torch.empty((2, B)), ]
for _ in range(3):
    sync(); t0 = time.perf_counter()
    a = torch.empty((7, 3, B), dtype=torch.float64, device="cuda"); b = torch.empty((2, 3, B), dtype=torch.float64, device="cuda")
    t0 = tick("torch.empty x2", t0)
    z = torch.zeros((B,), dtype=torch.int32, device="cuda")
    t0 = tick("torch.zeros", t0)
    del a, b, z
print("fp64 peak", nbk.fp64_peak(0, 20000))
import cProfile, pstats
pr = cProfile.Profile()
sync()
pr.enable()
for _ in range(5):
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    sync()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
