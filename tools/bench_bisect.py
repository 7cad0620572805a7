import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp
B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0_host = torch.from_numpy(y0).pin_memory(); p_host = torch.from_numpy(p).pin_memory()
y0d, pd = y0_host.cuda(), p_host.cuda()
def loop(n, events=False, sums=False, keep=False):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = torch.zeros((), dtype=torch.int64, device="cuda"); evs = []
    e0.record()
    for _ in range(n):
        s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9, device=0)
        if events:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
        t, y = s.run([10.0])
        if events: b.record(); evs.append((a, b))
        if sums:
            acc += s._nacc.sum(); acc += s._nacc.sum() + s._nrej.sum()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
loop(5)
print("plain        %.3f" % loop(20))
print("events       %.3f" % loop(20, events=True))
print("sums         %.3f" % loop(20, sums=True))
print("events+sums  %.3f" % loop(20, events=True, sums=True))
print("plain        %.3f" % loop(20))
# after a host-buffer pass (as the bench warm-up of the e2e arm does not precede, but check)
s = nbk.RungeKutta45("lorenz", 0.0, y0_host, p_host, rtol=1e-6, atol=1e-9, device=0); s.run([10.0]); del s
print("plain after host pass %.3f" % loop(20))
import torch.cuda
print("peak", nbk.fp64_peak(0, 20000))
print("plain after peak      %.3f" % loop(20))
