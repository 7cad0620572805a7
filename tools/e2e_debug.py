"""Diagnostic (GPU box): the e2e arm of bench.py pass by pass -- wall time per pass, device / pinned-host allocator activity."""
import gc, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

dev = torch.device("cuda", 0)
y0, p = ensembles.lorenz_ensemble(1_000_000, seed=0)
y0_host, p_host = torch.from_numpy(y0).pin_memory(), torch.from_numpy(p).pin_memory()
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]


def e2e_pass(i):
    with torch.cuda.stream(streams[i % 2]):
        s = nbk.RungeKutta45("lorenz", 0.0, y0_host, p_host, rtol=1e-6, atol=1e-9, device=0)
        t, y = s.run([10.0], wait=False)
        nacc, nrej = s.counts(wait=False)
    return s, y, nacc


def stats():
    d = torch.cuda.memory_stats(dev)
    h = torch.cuda.host_memory_stats() if hasattr(torch.cuda, "host_memory_stats") else {}
    return (d.get("num_device_alloc", 0), d.get("num_alloc_retries", 0), h.get("num_host_alloc", h.get("allocation.all.allocated", 0)),
            sum(gc.get_count()))


pending = None
rows = []
for i in range(40):
    t0 = time.perf_counter()
    item = e2e_pass(i)
    t1 = time.perf_counter()
    if pending is not None:
        pending[0].synchronize()
        n = int(pending[2].sum())
    pending = item
    t2 = time.perf_counter()
    rows.append((i, round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2), stats()))
for r in rows:
    print(r)
print("gc thresholds", gc.get_threshold(), "collections", [g["collections"] for g in gc.get_stats()])
