"""Diagnostic (GPU box): host-side cost of one pass (constructor + run([T])) of a small Lorenz ensemble, where the
kernels are short enough for the Python shell to become the limiter (strong scaling: 125 000 trajectories per GPU).
Prints wall time per pass with the GPU idle-waiting excluded (enqueue only) and a cProfile listing."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles

B = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
y0, p = ensembles.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()


def one():
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    s.run([10.0])
    return s


for _ in range(20):
    one()
torch.cuda.synchronize()
# enqueue-only cost: B tiny so that the GPU never back-pressures the host
t0 = time.perf_counter()
for _ in range(reps):
    one()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: host enqueue {1e3 * (t1 - t0) / reps:.3f} ms per pass; incl. drain {1e3 * (t2 - t0) / reps:.3f} ms per pass")
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    one()
pr.disable()
torch.cuda.synchronize()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(35)
print(out.getvalue())
