"""Measured parity distributions behind every test bar that is not the plain north-star bar (GPU box).

Writes gpurun_out/r02_parity.json (copied to profiles/ by hand).  For each relaxed bar of tests/ the MEASURED
distribution is recorded, so that the tests can assert "measured + margin" instead of a round number:
  c5_*        K3 (one CTA per trajectory) on the stability-limited reaction-diffusion grid: |dy| / (rtol |y| + atol)
              against the bit-exact NumPy restatement of the reference AND against the C port, plus the two CPU
              restatements against each other (the noise floor of the problem itself)
  fixed_*     Lorenz fixed-step golden rows (chaotic growth of rounding differences): relative error
  dop853      20 000 Lorenz trajectories, DOP853, dense output on a grid: fraction within tolerance
  smoke       the __graft_entry__.smoke() case
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import nbkode_b200 as nbk
from oracle import nbk_oracle_c as oc, nbk_oracle_np as onp

RTOL, ATOL = 1e-6, 1e-9
out = {}


def ratio(a, b, rtol=RTOL, atol=ATOL):
    return np.abs(np.asarray(a) - np.asarray(b)) / (rtol * np.abs(np.asarray(b)) + atol)


def dist(r):
    r = np.asarray(r).ravel()
    q = np.quantile(r, [0.5, 0.99, 0.9999])
    return {"max": float(r.max()), "p9999": float(q[2]), "p99": float(q[1]), "p50": float(q[0]), "n": int(r.size)}


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b))))


# ---------------------------------------------------------------- C5 / K3
t0 = time.time()
for cls, n, B, nnp in [("RungeKutta45", 2048, 48, 12), ("RungeKutta23", 2048, 8, 4), ("RungeKutta45", 100, 40, 8),
                       ("RungeKutta45", 700, 12, 4), ("RungeKutta45", 3000, 6, 2), ("RungeKutta45", 48, 16, 8),
                       ("RungeKutta45", 1024, 8, 4), ("RungeKutta23", 1000, 8, 4)]:
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    ts = [0.25, 0.5]
    ref = oc.run_ensemble(cls, "rd1d", y0, p, ts, rtol=RTOL, atol=ATOL, nthreads=os.cpu_count())
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, rtol=RTOL, atol=ATOL)
    t, y = s.run(ts)
    row = {"gpu_vs_c_port": dist(ratio(y, ref["ys"])), "d_steps_vs_c_port_max": int(np.abs(s.n_accepted - ref["n_accepted"]).max())}
    ynp, nnp_acc = [], []
    for i in range(nnp):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[i], p[i], rtol=RTOL, atol=ATOL)
        _, yo = onp.run(o, ts)
        ynp.append(yo)
        nnp_acc.append(o.n_accepted)
    ynp = np.array(ynp)
    row["gpu_vs_numpy_restatement"] = dist(ratio(y[:nnp], ynp))
    row["c_port_vs_numpy_restatement"] = dist(ratio(ref["ys"][:nnp], ynp))
    row["d_steps_vs_numpy_max"] = int(np.abs(s.n_accepted[:nnp] - np.array(nnp_acc)).max())
    row["numpy_trajectories"] = nnp
    out[f"c5_{cls}_n{n}_B{B}"] = row
    print(cls, n, B, json.dumps(row), f"[{time.time() - t0:.0f} s]", flush=True)

# the BASELINE C5 configuration itself: run([1.0]), 12 trajectories against the NumPy restatement
y0, p = onp.rd1d_ensemble(12, n=2048, seed=3)
s = nbk.RungeKutta45("rd1d", 0.0, y0, p, rtol=RTOL, atol=ATOL)
_, y = s.run([1.0])
ynp, acc = [], []
for i in range(12):
    o = onp.AdaptiveRK("RungeKutta45", onp.rhs_rd1d, 0.0, y0[i], p[i], rtol=RTOL, atol=ATOL)
    _, yo = onp.run(o, [1.0])
    ynp.append(yo); acc.append(o.n_accepted)
ref = oc.run_ensemble("RungeKutta45", "rd1d", y0, p, [1.0], rtol=RTOL, atol=ATOL, nthreads=os.cpu_count())
out["c5_config_t1"] = {"gpu_vs_numpy_restatement": dist(ratio(y, np.array(ynp))), "gpu_vs_c_port": dist(ratio(y, ref["ys"])),
                       "c_port_vs_numpy_restatement": dist(ratio(ref["ys"], np.array(ynp))),
                       "d_steps_vs_numpy_max": int(np.abs(s.n_accepted - np.array(acc)).max()),
                       "accepted_steps_mean": float(np.mean(acc))}
print("c5 config", json.dumps(out["c5_config_t1"]), flush=True)

# ---------------------------------------------------------------- Lorenz fixed-step golden rows
from conftest import load_golden

worst = {}
for fname in ("fixed_single.json", "erk_fixed.json"):
    for c in load_golden(fname):
        if c["rhs"] != "lorenz":
            continue
        kw = {k: (None if v == "None" else v) for k, v in c.get("kw", {}).items()}
        s = getattr(nbk, c["cls"])(c["rhs"], 0.0, np.array(c["y0"]), np.array(c["p"]), h=c["h"], **kw)
        t, y = s.run(c["ts"])
        e = {"ys": rel(y, c["ys"]), "y_end": rel(s.y, c["y_end"]), "f_end": rel(s.f, c["f_end"]), "t_final": float(c["ts"][-1])}
        worst[f"{fname}:{c['cls']}:{kw}"] = e
out["fixed_lorenz"] = {"cases": worst, "max_rel": max(max(v["ys"], v["y_end"], v["f_end"]) for v in worst.values())}
print("fixed lorenz max rel", out["fixed_lorenz"]["max_rel"], flush=True)

# ---------------------------------------------------------------- DOP853, 20k Lorenz trajectories
B = 20000
y0, p = onp.lorenz_ensemble(B, seed=321)
ts = np.linspace(0.5, 10, 20)
ref = oc.run_ensemble("DOP853", "lorenz", y0, p, ts, rtol=RTOL, atol=ATOL, nthreads=os.cpu_count())
s = nbk.DOP853("lorenz", 0.0, y0, p, rtol=RTOL, atol=ATOL)
t, y = s.run(ts)
r = ratio(y, ref["ys"]).max(axis=2)  # (B, m)
dn = np.abs(s.n_accepted - ref["n_accepted"])
out["dop853_lorenz_20k"] = {
    "frac_within_tol_all_grid": float((r <= 1).all(axis=1).mean()), "frac_within_tol_t_le_4": float((r[:, :8] <= 1).all(axis=1).mean()),
    "frac_steps_identical": float((dn == 0).mean()), "frac_steps_within_1": float((dn <= 1).mean()),
    "err_over_tol_final": dist(r[:, -1]), "err_over_tol_t4": dist(r[:, 7]), "n_outliers": int((~(r <= 1).all(axis=1)).sum())}
print("dop853", json.dumps(out["dop853_lorenz_20k"]), flush=True)

# ---------------------------------------------------------------- smoke case
y0, p = onp.lorenz_ensemble(4096, seed=0)
ref = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [0.5, 2.0], rtol=RTOL, atol=ATOL, nthreads=os.cpu_count())
s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=RTOL, atol=ATOL)
t, y = s.run([0.5, 2.0])
out["smoke_lorenz_4096_t2"] = {"err_over_tol": dist(ratio(y, ref["ys"])),
                               "d_steps_max": int(np.abs(s.n_accepted - ref["n_accepted"]).max())}
print("smoke", json.dumps(out["smoke_lorenz_4096_t2"]), flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_parity.json"), "w"), indent=1)
