import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp
B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
def stat():
    s = torch.cuda.memory_stats(); return s.get("num_device_alloc", -1), s.get("num_device_free", -1), s["reserved_bytes.all.current"] >> 20, s["allocated_bytes.all.current"] >> 20
orig = torch.empty
log = []
def timed_empty(*a, **k):
    t0 = time.perf_counter(); r = orig(*a, **k); log.append((1e3 * (time.perf_counter() - t0), tuple(a[0]) if a else None)); return r
torch.empty = timed_empty
for it in range(4):
    log.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(it, "ctor cpu %.3f ms, +sync %.3f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)), stat())
    print("   empties:", ["%.3f" % x[0] for x in log])
    s.run([10.0]); torch.cuda.synchronize()
    del s
    torch.cuda.synchronize(); print("   after del", stat())
