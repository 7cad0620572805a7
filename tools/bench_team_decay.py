"""Diagnostic (GPU box): element-wise decay under the classes without a dedicated sub-warp kernel -- per-thread kernels
("decay:thread") against sub-warp teams (forced from n = 4 by NBK_GROUP_DECAY_MINN), to place the crossover."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
os.environ["NBK_GROUP_DECAY_MINN"] = "4"
rng = np.random.default_rng(0)
for cls, kw, T in (("AdamsBashforth4", dict(h=0.01), 2.0), ("RungeKutta4", dict(h=0.02), 2.0), ("DOP853", dict(rtol=1e-8, atol=1e-11), 10.0)):
    for n in (4, 8, 12, 16, 24, 32):
        y0 = torch.from_numpy(rng.uniform(0.5, 2, (B, n))).cuda()
        p = torch.from_numpy(rng.uniform(0.1, 1.0, (B, n))).cuda()
        row = {"cls": cls, "n": n}
        for label, name in (("thread", "decay:thread"), ("team", "decay")):
            try:
                best = None
                for _ in range(3):
                    s = getattr(nbk, cls)(name, 0.0, y0, p, **kw)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); s.run([T]); e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                    best = ms if best is None else min(best, ms)
                row[label] = {"ms": round(best, 3), "regime": s._regime, "lanes": s._group}
            except Exception as exc:  # noqa: BLE001
                row[label] = {"error": str(exc)[:120]}
        if "ms" in row["thread"] and "ms" in row["team"]:
            row["speedup"] = round(row["thread"]["ms"] / row["team"]["ms"], 2)
        print(json.dumps(row), flush=True)
