"""Diagnostic (GPU box): where one pass (constructor + run([T])) of a small Lorenz ensemble spends its time.
  host   -- enqueue cost of constructor and run() with the GPU drained before every pass (no back-pressure)
  init   -- device time between events placed around the constructor / around run()
  b2b    -- device time per pass of back-to-back passes, with 1 and with R rotating inputs / live states"""
import collections, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

B = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
y0, p = ensembles.lorenz_ensemble(B, 0)
mk_in = lambda: (torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda())
ev = lambda: torch.cuda.Event(enable_timing=True)


def measure(ring):
    inputs = [mk_in() for _ in range(ring)]
    live = collections.deque(maxlen=ring)
    for i in range(8):
        s = nbk.RungeKutta45("lorenz", 0.0, *inputs[i % ring], rtol=1e-6, atol=1e-9); s.run([10.0]); live.append(s)
    torch.cuda.synchronize()
    # (1) host cost, GPU drained before every pass
    hc, hr, di, dr = [], [], [], []
    for i in range(reps):
        torch.cuda.synchronize()
        e0, e1, e2 = ev(), ev(), ev()
        t0 = time.perf_counter(); e0.record()
        s = nbk.RungeKutta45("lorenz", 0.0, *inputs[i % ring], rtol=1e-6, atol=1e-9)
        t1 = time.perf_counter(); e1.record()
        s.run([10.0])
        t2 = time.perf_counter(); e2.record()
        live.append(s)
        torch.cuda.synchronize()
        hc.append(1e3 * (t1 - t0)); hr.append(1e3 * (t2 - t1)); di.append(e0.elapsed_time(e1)); dr.append(e1.elapsed_time(e2))
    med = lambda v: float(np.median(v))
    print(f"ring={ring} B={B}: host ctor {med(hc):.3f} ms, host run {med(hr):.3f} ms | device: ctor span {med(di):.3f} ms, run span {med(dr):.3f} ms")
    # (2) back to back
    torch.cuda.synchronize()
    a, b = ev(), ev()
    t0 = time.perf_counter(); a.record()
    for i in range(reps):
        s = nbk.RungeKutta45("lorenz", 0.0, *inputs[i % ring], rtol=1e-6, atol=1e-9); s.run([10.0]); live.append(s)
    b.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"ring={ring} B={B}: back-to-back {a.elapsed_time(b) / reps:.3f} ms per pass on the device, host loop {1e3 * (t1 - t0) / reps:.3f} ms per pass")


for ring in (1, 5):
    measure(ring)
