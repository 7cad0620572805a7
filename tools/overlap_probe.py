"""Diagnostic (GPU box): throughput of back-to-back 1M-trajectory Lorenz RK45 passes on 1, 2 and 3 CUDA streams
(the tail of one persistent kernel overlaps the head of the next when they sit on different streams)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles as onp  # input generators

B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
for nstreams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    K = 12
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solvers = []
        for k in range(K):
            st = streams[k % nstreams]
            st.wait_event(e0)
            with torch.cuda.stream(st):
                s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
                s.run([10.0])
                solvers.append(s)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    acc = int(solvers[-1]._nacc.sum().item())
    print(json.dumps({"streams": nstreams, "ms_per_pass": ms, "accepted_per_s": acc / ms * 1e3}), flush=True)
