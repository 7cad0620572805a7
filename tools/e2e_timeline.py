"""Diagnostic (GPU box): GPU- and host-side timeline of the two-stream end-to-end pipeline of bench.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles as onp  # input generators
B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0h, ph = torch.from_numpy(y0).pin_memory(), torch.from_numpy(p).pin_memory()
dev = torch.device("cuda", 0)
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
base = torch.cuda.Event(enable_timing=True)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
def e2e_pass(i, log):
    with torch.cuda.stream(streams[i % 2]):
        h0 = time.perf_counter(); e0 = ev()
        s = nbk.RungeKutta45("lorenz", 0.0, y0h, ph, rtol=1e-6, atol=1e-9, device=0)
        h1 = time.perf_counter(); e1 = ev()
        t, y = s.run([10.0], wait=False)
        h2 = time.perf_counter(); e2 = ev()
        nacc, nrej = s.counts(wait=False)
        h3 = time.perf_counter(); e3 = ev()
    log.append(dict(i=i, host=(h0, h1, h2, h3), ev=(e0, e1, e2, e3)))
    return s, y, nacc
for rep in range(2):
    log = []; pend = None
    torch.cuda.synchronize(); base.record(); T0 = time.perf_counter()
    fin = []
    for i in range(8):
        item = e2e_pass(i, log)
        if pend is not None:
            a = time.perf_counter(); pend[0].synchronize(); n = int(pend[2].sum()); fin.append((a, time.perf_counter()))
        pend = item
    a = time.perf_counter(); pend[0].synchronize(); fin.append((a, time.perf_counter()))
    torch.cuda.synchronize(); T1 = time.perf_counter()
print("pipelined: %.2f ms/pass" % (1e3 * (T1 - T0) / 8))
for r in log:
    h = [1e3 * (x - T0) for x in r["host"]]; g = [base.elapsed_time(e) for e in r["ev"]]
    print("pass %d host: ctor %.2f-%.2f run->%.2f counts->%.2f | gpu: start %.2f ctor-done %.2f run-done+D2H(y) %.2f counts-D2H %.2f" % (r["i"], h[0], h[1], h[2], h[3], *g))
print("finish calls (host ms):", " ".join("(%.2f-%.2f)" % (1e3 * (a - T0), 1e3 * (b - T0)) for a, b in fin))
try:
    st = torch.cuda.host_memory_stats(); print({k: v for k, v in st.items() if "num_host_alloc" in k or "allocated_bytes.current" in k or "host_alloc_time" in k})
except Exception as e: print("no host stats", e)
