"""Diagnostic (GPU box): the sub-warp regime (csrc/nbk_group.cuh) against the regimes it replaces, as the state size grows.
RungeKutta45, rtol 1e-6 / atol 1e-9, t in [0, 10]; right-hand sides: element-wise decay (per-component rates) and the
periodic reaction-diffusion stencil.  Per row: time and FP64 fraction with the sub-warp kernels, the same with
NBK_NO_GROUP=1 (per-thread kernels for decay, one CTA per trajectory for rd1d), parity of a 256-trajectory prefix against
the C port of the reference (test infrastructure).
usage: bench_group.py [B] [n,n,...] [rhs,rhs]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from nbkode_b200 import ensembles
from oracle import nbk_oracle_c as oc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
NS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 12, 16, 24, 32, 48, 64]
RHS = sys.argv[3].split(",") if len(sys.argv) > 3 else ["decay", "rd1d"]
RTOL, ATOL, T = 1e-6, 1e-9, 10.0
peak, _ = nbk.fp64_peak(0, 20000)
rng = np.random.default_rng(0)


def problem(rhs, n, B):
    if rhs == "decay":
        return rng.uniform(0.5, 2, (B, n)), rng.uniform(0.1, 1.0, (B, n)), 6 * n + 66 * n + 18
    y0, p = ensembles.rd1d_ensemble(B, n=n, seed=3)
    return y0, p, 6 * 8 * n + 66 * n + 18


def timed(rhs, y0d, pd, reps=3):
    best, s = None, None
    for _ in range(reps):
        s = nbk.RungeKutta45(rhs, 0.0, y0d, pd, rtol=RTOL, atol=ATOL)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t, y = s.run([T]); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    s.raise_for_status()
    return best, s, y


for rhs in RHS:
    for n in NS:
        y0, p, flop = problem(rhs, n, B)
        y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
        row = {"rhs": rhs, "n": n, "B": B}
        for label, env in (("group", None), ("before", "1")):
            if env:
                os.environ["NBK_NO_GROUP"] = env
            else:
                os.environ.pop("NBK_NO_GROUP", None)
            try:
                ms, s, y = timed(rhs, y0d, pd)
            except Exception as exc:  # noqa: BLE001
                row[label] = {"error": str(exc)[:200]}
                continue
            att = int((s._nacc.sum() + s._nrej.sum()).item())
            tf = att * flop / ms * 1e3 / 1e12
            row[label] = {"ms": round(ms, 3), "tflops": round(tf, 2), "frac": round(tf / peak, 3), "launch": s.last_launch(),
                          "regime": s._regime}
            if True:
                k = 256
                ref = oc.run_ensemble("RungeKutta45", rhs, y0[:k], p[:k], [T], rtol=RTOL, atol=ATOL, nthreads=8)
                yy = y[:k, 0].cpu().numpy()
                err = np.abs(yy - ref["ys"][:, 0]) / (RTOL * np.abs(ref["ys"][:, 0]) + ATOL)
                dn = np.abs(s._nacc[:k].cpu().numpy() - ref["n_accepted"])
                row[label]["parity"] = {"max_err_over_tol": float(err.max()), "max_dsteps": int(dn.max())}
        os.environ.pop("NBK_NO_GROUP", None)
        if "ms" in row.get("group", {}) and "ms" in row.get("before", {}):
            row["speedup"] = round(row["before"]["ms"] / row["group"]["ms"], 2)
        print(json.dumps(row), flush=True)
        del y0d, pd
        torch.cuda.empty_cache()
