import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp
B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0h, ph = torch.from_numpy(y0).pin_memory(), torch.from_numpy(p).pin_memory()
dev = torch.device("cuda", 0)
# raw PCIe numbers
x = torch.empty(24_000_000 // 8 * 4, dtype=torch.float64, pin_memory=True); xd = x.cuda()
for name, fn in (("H2D 96MB", lambda: xd.copy_(x, non_blocking=True)), ("D2H 96MB", lambda: x.copy_(xd, non_blocking=True))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:.2f} ms  {x.numel()*8/dt/1e9:.1f} GB/s")
t0 = time.perf_counter(); z = torch.empty(3_000_000, dtype=torch.float64, pin_memory=True); t1 = time.perf_counter(); del z
z = torch.empty(3_000_000, dtype=torch.float64, pin_memory=True); t2 = time.perf_counter()
print(f"pinned alloc 24MB: first {1e3*(t1-t0):.2f} ms, again {1e3*(t2-t1):.2f} ms")
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
def e2e_pass(i):
    with torch.cuda.stream(streams[i % 2]):
        s = nbk.RungeKutta45("lorenz", 0.0, y0h, ph, rtol=1e-6, atol=1e-9, device=0)
        t, y = s.run([10.0], wait=False)
        nacc, nrej = s.counts(wait=False)
    return s, y, nacc
for i in range(3):
    it = e2e_pass(i); it[0].synchronize()
torch.cuda.synchronize()
T0 = time.perf_counter(); pend = None; log = []
for i in range(10):
    a = time.perf_counter(); item = e2e_pass(i); b = time.perf_counter()
    if pend is not None: pend[0].synchronize(); n = int(pend[2].sum())
    c = time.perf_counter(); log.append((1e3*(b-a), 1e3*(c-b))); pend = item
pend[0].synchronize(); torch.cuda.synchronize(); T1 = time.perf_counter()
print("pipelined: %.2f ms/pass" % (1e3*(T1-T0)/10)); print("  (enqueue, wait) per pass:", " ".join(f"({x:.2f},{y:.2f})" for x, y in log))
# sequential, single stream, blocking
torch.cuda.synchronize(); T0 = time.perf_counter()
for i in range(10):
    s = nbk.RungeKutta45("lorenz", 0.0, y0h, ph, rtol=1e-6, atol=1e-9, device=0); t, y = s.run([10.0]); na, nr = s.counts()
T1 = time.perf_counter(); print("sequential: %.2f ms/pass" % (1e3*(T1-T0)/10))
