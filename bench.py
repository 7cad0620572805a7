#!/usr/bin/env python
"""bench.py -- headline benchmark of BASELINE.json: accepted RK45 steps/s on the Lorenz-63
ensemble (configs[1]: ONE ensemble of 1M trajectories, random y0, sigma/rho/beta sweep,
rtol=1e-6, atol=1e-9, t in [0, 10], FP64).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over the ensemble: constructor kernel (f0 + initial
step size) and the persistent RK45 kernel to t = 10.  Under torchrun the ONE 1M-trajectory
ensemble of the north star is sharded by trajectory over the N ranks (strong scaling, no
data-path collective); the weak-scaling figure (1M per GPU) of the same
run sits in `extra.weak_scaling`.

Printed JSON line (rank 0): the contract of the task statement; additionally
  roofline      FP64 vector pipe: algorithmic FLOPs (264 per attempted step)
                / run-kernel time, against the DFMA peak measured in this run (nbk_fp64_peak)
  cpu_baseline  N = 1: the LIVE reference (nbkode's numba path from baseline/_ref, one Solver
                object per trajectory, fork()ed workers on all host cores) over a 65 536-trajectory
                prefix of the same ensemble (kind "reference"); `cpu_baseline_port` = the C port
                (oracle/nbk_oracle.c, pthreads) over the full ensemble
  parity        N = 1: the GPU result of the full 1M ensemble against the C port, and of the
                prefix against the live reference: fraction within rtol*|y|+atol, fraction with
                accepted-step counts within +-1, outliers listed
  e2e           the same metric through the public API with pinned HOST buffers in and out
  extra.configs N = 1: one timed pass each of BASELINE configs[2..4] at full size and of four mid-size systems
                (sub-warp regime); N > 1: extra.weak_scaling (1M trajectories per GPU) and extra.c5_strong_scaling
                (configs[4]'s 65 536 grids of n = 2048 sharded over the ranks)
`--impl reference` times the live reference (numba) on the host cores, each step a bounded
sample of the same ensemble (the C port only if baseline/_ref is absent).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "numbakit-ode_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

B_TOTAL = 1_000_000
N_PREFIX = 65_536  # trajectories the live reference integrates beside the kernels
T_END = 10.0
RTOL, ATOL = 1e-6, 1e-9
FLOP_PER_ATTEMPT = 264  # 6R + 66n + 18 with R = 8, n = 3
WARMUP_MIN_S = float(os.environ.get("NBK_BENCH_WARMUP_S", 1.0))  # shortest warm-up of the kernel arm, seconds of wall time
METRIC = "accepted RK45 steps/s, 1M-traj Lorenz FP64"
UNIT = "accepted steps/s"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the integrator kernel, per launch, from the committed
    `ncu --set full` capture of this same command (profiles/); None if no capture is present."""
    import glob

    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full_K1_nbk_run_rk45_lorenz.json")), reverse=True):
        try:
            d = json.load(open(path))
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            return float(rd[0]) * unit[rd[1]] + float(wr[0]) * unit[wr[1]]
        except Exception:  # noqa: BLE001
            continue
    return None


def workload_config(n_gpus, b_total=B_TOTAL, l2=None):
    per = -(-b_total // n_gpus)
    mb = per * (368 + 48) / 1e6
    return {
        "workload": "Lorenz-63 ensemble, RungeKutta45 rtol=1e-6 atol=1e-9, t in [0,10], run([10.0])",
        "trajectories_total": b_total,
        "trajectories_per_gpu": per,
        "parallelism": f"one ensemble sharded by trajectory x{n_gpus}, no collective",
        "l2": l2 or l2_policy(per)[1],
    }


def l2_policy(per_gpu):
    """Working set of one pass per GPU (inputs 48 B + state 368 B per trajectory) against the 126 MB L2: when it
    does not exceed the L2 by itself, consecutive passes rotate over R distinct input copies and R live solver
    states, so that no pass finds its data in L2."""
    mb = per_gpu * (368 + 48) / 1e6
    ring = 1 if mb > 252 else int(252 // mb) + 1
    if ring == 1:
        return 1, f"working set ({mb:.0f} MB per GPU: inputs + resumable state) exceeds the 126 MB L2; no explicit flush"
    return ring, (f"working set is {mb:.0f} MB per GPU: passes rotate over {ring} input copies and {ring} live states "
                  f"({ring * mb:.0f} MB > 2 x 126 MB L2); no explicit flush")


# ----------------------------------------------------------------------------------------
class no_gc:
    """Timed regions run with the cyclic garbage collector off (after one full collection), so that no collection pass
    lands between two launches while the host has not yet run ahead of the device."""

    def __enter__(self):
        import gc

        gc.collect()
        self.was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        import gc

        if self.was:
            gc.enable()
        return False


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs, with the
    profiling guide's recipe: `nvidia-smi --query-gpu=... -lms 200` in a separate process (NVML
    calls from a thread of this process were measured to delay kernel launches by ~2 ms/step)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=None):
        self.index, self.proc, self.first = index, None, None
        self.period_ms = int(period_ms or os.environ.get("NBK_BENCH_SMI_MS", 100))

    def start(self):
        import shutil
        import subprocess

        exe = shutil.which("nvidia-smi")
        if exe is None or self.period_ms <= 0:
            return
        try:
            self.proc = subprocess.Popen(
                [exe, f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 f"-lms={self.period_ms}"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            # nvidia-smi's start-up (NVML initialisation, ~100 ms) must not overlap the timed region: wait for its
            # first sample before going on
            self.first = self.proc.stdout.readline()
        except Exception:  # noqa: BLE001
            self.proc = None

    def finish(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.25)  # let at least one more sample land
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        out = (self.first or "") + out
        sm, mx, power, reasons = [], None, [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": [], "note": "no samples"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------------------
# CPU legs (never on the product path): the live reference and the C port
# ----------------------------------------------------------------------------------------
def live_reference_pool():
    """fork()ed numba workers of the LIVE reference (baseline/nbkode_ref.py); None + reason if baseline/_ref is
    absent.  Must be called before CUDA is initialised in this process (the workers are fork()ed)."""
    try:
        from baseline import nbkode_ref

        if not nbkode_ref.installed():
            return None, nbkode_ref.install()  # build container only; "unavailable: ..." on the GPU box
        return nbkode_ref.Pool(os.cpu_count() or 1, "lorenz", "RungeKutta45", rtol=RTOL, atol=ATOL), "ok"
    except Exception as exc:  # noqa: BLE001
        return None, f"unavailable: {type(exc).__name__}: {exc}"


def parity_stats(y, nacc, y_ref, nacc_ref, max_list=16):
    """Final states within rtol*|y|+atol and accepted-step counts within +-1 (the north-star bars), as fractions
    over the ensemble, with the outliers listed."""
    y, y_ref = np.asarray(y).reshape(len(y_ref), -1), np.asarray(y_ref).reshape(len(y_ref), -1)
    ratio = np.max(np.abs(y - y_ref) / (RTOL * np.abs(y_ref) + ATOL), axis=1)
    dn = np.abs(np.asarray(nacc, dtype=np.int64) - np.asarray(nacc_ref, dtype=np.int64))
    bad = np.nonzero((ratio > 1.0) | (dn > 1) | ~np.isfinite(ratio))[0]
    order = bad[np.argsort(-np.nan_to_num(ratio[bad], nan=np.inf))][:max_list]
    q = np.quantile(ratio, [0.5, 0.99, 0.9999])
    return {
        "trajectories": int(len(y_ref)),
        "frac_within_tol": float(np.mean(ratio <= 1.0)),
        "frac_steps_within_1": float(np.mean(dn <= 1)),
        "frac_steps_identical": float(np.mean(dn == 0)),
        "err_over_tol": {"p50": float(q[0]), "p99": float(q[1]), "p9999": float(q[2]), "max": float(np.max(ratio))},
        "n_outliers": int(len(bad)),
        "outliers": [{"trajectory": int(i), "err_over_tol": float(ratio[i]), "d_steps": int(dn[i])} for i in order],
        "bars": "final state |dy| <= rtol*|y_ref| + atol (rtol=1e-6, atol=1e-9); accepted steps within +-1; a flipped "
                "accept/reject decision on a chaotic trajectory is amplified beyond the bar -- those are the outliers",
    }


def cpu_port(y0, p, threads=None):
    """C port of the reference algorithm (oracle/nbk_oracle.c, pthreads) over the whole ensemble."""
    from oracle import nbk_oracle_c as oc

    threads = threads or os.cpu_count() or 1
    t0 = time.perf_counter()
    r = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [T_END], rtol=RTOL, atol=ATOL, nthreads=threads, dense=True)
    dt = time.perf_counter() - t0
    return r, {
        "value": float(r["n_accepted"].sum() / dt), "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"all {len(y0)} trajectories of the ensemble, {dt:.2f} s wall, C port (oracle/nbk_oracle.c) with pthreads",
    }


def cpu_live(pool, y0, p, n=N_PREFIX):
    """nbkode's numba path over a prefix of the ensemble: throughput = accepted steps / slowest worker's time."""
    n = min(n, len(y0))
    r = pool.run(y0[:n], p[:n], [T_END], count=True)
    return r, {
        "value": float(r["n_accepted"].sum() / r["wall"]), "unit": UNIT, "cores": pool.workers, "kind": "reference",
        "sample": (f"first {n} trajectories of the ensemble, one nbkode.RungeKutta45 object per trajectory, run([10.0]), "
                   f"{pool.workers} fork()ed workers, slowest worker {r['wall']:.2f} s; numba JIT ({pool.jit_s:.1f} s, once, "
                   "in the parent) and the accepted-step recount are outside the timed region"),
        "jit_s": pool.jit_s,
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from nbkode_b200 import ensembles

    y0, p = ensembles.lorenz_ensemble(B_TOTAL, seed=0)
    pool, why = live_reference_pool()
    n_iter = args.warmup + args.steps
    if pool is not None:
        # bounded sample per step: ~2.5 s of wall time on this host, a fixed prefix of the ensemble
        probe = 64 * pool.workers
        pool.run(y0[:probe], p[:probe], [T_END])  # first call in every worker: page faults after fork(), cold caches
        r = pool.run(y0[:probe], p[:probe], [T_END])
        n = int(min(B_TOTAL, max(probe, 2.5 * probe / r["wall"])))
        times = []
        for it in range(n_iter):
            r = pool.run(y0[:n], p[:n], [T_END])
            if it >= args.warmup:
                times.append(r["wall"])
        steps_per_pass = int(pool.run(y0[:n], p[:n], [T_END], count=True)["n_accepted"].sum())
        kind, cores = "reference", pool.workers
        sample = (f"first {n} trajectories of the ensemble per step, the unmodified nbkode (numba) from baseline/_ref: one "
                  f"RungeKutta45 object per trajectory, run([10.0]), {cores} fork()ed workers; time = slowest worker; JIT "
                  f"({pool.jit_s:.1f} s) excluded")
        pool.close()
    else:
        from oracle import nbk_oracle_c as oc

        cores = os.cpu_count() or 1
        n, times = B_TOTAL, []
        for it in range(n_iter):
            t0 = time.perf_counter()
            r = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [T_END], rtol=RTOL, atol=ATOL, nthreads=cores, dense=True)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        steps_per_pass = int(r["n_accepted"].sum())
        kind = "port"
        sample = f"all {n} trajectories per step, C port of nbkode's RK45 path (pthreads); live reference {why}"
    value = steps_per_pass * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# BASELINE configs[2..4], one timed pass each at full size (N = 1)
# ----------------------------------------------------------------------------------------
def extra_configs(nbk, torch, dev, peak_tf, sizes=None):
    from nbkode_b200 import _lib, ensembles

    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001
        hbm, hbm_src = 6650.0, "B200_PROFILING.md fallback"
    sizes = sizes or {"c3": 4_000_000, "c4": 262_144, "c5": 65_536}
    rows = []

    def timed(make, run, reps=2):
        best = None
        for _ in range(reps):  # the first pass also warms the allocator (and NVRTC, where a program is not ahead-of-time)
            s = make()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with no_gc():
                e0.record()
                out = run(s)
                e1.record()
                torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
            acc, att = int(s._nacc.sum().item()), int((s._nacc.sum() + s._nrej.sum()).item())
            s.raise_for_status()
            launch = s.last_launch()
            del out, s
        return best, acc, att, launch

    def guarded(name, fn):
        try:
            rows.append(fn())
        except Exception as exc:  # noqa: BLE001  (a secondary config must not take the headline line down)
            rows.append({"config": name, "error": f"{type(exc).__name__}: {exc}"[:300]})
        torch.cuda.empty_cache()

    # ---- C3: Van der Pol, 1000-point dense output (64 GB per solver at 4M trajectories)
    B3, m = sizes["c3"], 1000
    y0, p = ensembles.vdp_ensemble(B3, seed=1)
    y0d, pd = torch.from_numpy(y0).to(dev), torch.from_numpy(p).to(dev)
    ts = np.linspace(0, 20, m)
    out_gb = B3 * m * 2 * 8 / 1e9

    def c3(cls, kw, flop, kernel, what):
        def fn():
            ms, acc, att, launch = timed(lambda: getattr(nbk, cls)("vdp", 0.0, y0d, pd, **kw), lambda s: s.run(ts))
            tf = att * flop / ms * 1e3 / 1e12
            return {"config": f"C3 Van der Pol {B3} x {m}-point grid, {what}", "kernel": kernel, "ms": ms, "accepted_steps": acc,
                    "steps_per_s": acc / ms * 1e3,
                    "roofline": {"bound": "hbm", "achieved": out_gb / ms * 1e3, "peak": hbm, "unit": "GB/s",
                                 "frac": out_gb / ms * 1e3 / hbm, "peak_source": hbm_src, "algorithmic_bytes": out_gb * 1e9},
                    "fp64": {"achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "flop_per_attempt": flop},
                    "launch": launch}
        return fn

    # FLOP per attempt: RK23 3R + 27n + 18, Adams-Bashforth-k R + 2kn + k + 1; Van der Pol R = 5, n = 2
    guarded("C3 AdamsBashforth4", c3("AdamsBashforth4", dict(h=0.01), 5 + 16 + 5, "nbk_run_ab4_vdp (K2)", "AdamsBashforth4 h=0.01"))
    guarded("C3 RungeKutta23", c3("RungeKutta23", dict(rtol=RTOL, atol=ATOL), 15 + 54 + 18, "nbk_run_rk23_vdp (K1)",
                                   "RungeKutta23 rtol=1e-6 atol=1e-9"))
    del y0d, pd

    # ---- C4: linear y' = A y, n = 128, shared dense A: right-hand side on the FP64 tensor pipe (DMMA)
    def c4():
        B4, n = sizes["c4"], 128
        y0, A = ensembles.linear_ensemble(B4, n=n, seed=2)
        y0d, Ad = torch.from_numpy(y0).to(dev), torch.from_numpy(A).to(dev)
        flop = 6 * 2 * n * n + 66 * n + 18
        ms, acc, att, launch = timed(lambda: nbk.RungeKutta45("linear", 0.0, y0d, Ad, rtol=RTOL, atol=ATOL), lambda s: s.run([10.0]))
        mma_peak, _ = _lib.fp64_mma_peak(dev.index or 0, 4000)
        tf = att * flop / ms * 1e3 / 1e12
        return {"config": f"C4 linear n={n} x {B4}, RungeKutta45, shared dense A", "kernel": "nbk_run_rk45_linear_mma (K4, DMMA)",
                "ms": ms, "accepted_steps": acc, "steps_per_s": acc / ms * 1e3,
                "roofline": {"bound": "fp64 tensor pipe (DMMA m8n8k4)", "achieved": tf, "peak": mma_peak, "unit": "TFLOP/s",
                             "frac": tf / mma_peak, "frac_of_dfma_peak": tf / peak_tf, "flop_per_attempt": flop},
                "launch": launch}

    guarded("C4 linear", c4)

    # ---- C5: reaction-diffusion method of lines, n = 2048, one CTA per trajectory
    def c5():
        B5, n = sizes["c5"], 2048
        y0, p5 = ensembles.rd1d_ensemble(B5, n=n, seed=3)
        y0d, p5d = torch.from_numpy(y0).to(dev), torch.from_numpy(p5).to(dev)
        flop = 6 * 8 * n + 66 * n + 18
        ms, acc, att, launch = timed(lambda: nbk.RungeKutta45("rd1d", 0.0, y0d, p5d, rtol=RTOL, atol=ATOL), lambda s: s.run([1.0]))
        tf = att * flop / ms * 1e3 / 1e12
        return {"config": f"C5 reaction-diffusion n={n} x {B5}, RungeKutta45, t in [0,1]", "kernel": "nbk_run_rk45_rd1d_cta (K3)",
                "ms": ms, "accepted_steps": acc, "steps_per_s": acc / ms * 1e3,
                "roofline": {"bound": "fp64", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                             "flop_per_attempt": flop},
                "launch": launch}

    guarded("C5 rd1d", c5)

    # ---- mid-size systems (the north star's register regime up to n = 32 and the band above it): sub-warp kernels,
    # G lanes of a warp per trajectory (nbk_group.cuh).  Element-wise decay with per-component rates (R = n) and the
    # periodic reaction-diffusion stencil (R = 8 per component), RungeKutta45, 262 144 trajectories to t = 10.
    def mid(rhs_name, n):
        def fn():
            Bm = sizes.get("mid", 262_144)
            if rhs_name == "decay":
                rng = np.random.default_rng(0)
                y0, pm, flop = rng.uniform(0.5, 2, (Bm, n)), rng.uniform(0.1, 1.0, (Bm, n)), 6 * n + 66 * n + 18
            else:
                y0, pm = ensembles.rd1d_ensemble(Bm, n=n, seed=3)
                flop = 6 * 8 * n + 66 * n + 18
            y0d, pmd = torch.from_numpy(y0).to(dev), torch.from_numpy(pm).to(dev)
            ms, acc, att, launch = timed(lambda: nbk.RungeKutta45(rhs_name, 0.0, y0d, pmd, rtol=RTOL, atol=ATOL), lambda s: s.run([10.0]), reps=3)
            tf = att * flop / ms * 1e3 / 1e12
            return {"config": f"mid-size {rhs_name} n={n} x {Bm}, RungeKutta45, t in [0,10]",
                    "kernel": f"nbk_run_jit (sub-warp regime, {launch.get('lanes_per_trajectory', 0)} lanes per trajectory)", "ms": ms,
                    "accepted_steps": acc, "steps_per_s": acc / ms * 1e3,
                    "roofline": {"bound": "fp64", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                                 "flop_per_attempt": flop},
                    "launch": launch}
        return fn

    for rhs_name, n in (("decay", 16), ("decay", 32), ("rd1d", 32), ("rd1d", 64)):
        guarded(f"mid-size {rhs_name} n={n}", mid(rhs_name, n))
    return rows


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_TOTAL, help="trajectories of the whole ensemble (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (live reference, C port, parity block)")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.configs / extra.weak_scaling")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # the live reference's workers are fork()ed BEFORE this process touches CUDA (N = 1 only, rank 0)
    cpu_legs = world == 1 and not args.no_cpu_baseline
    pool, pool_note = live_reference_pool() if cpu_legs else (None, "not requested")

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import nbkode_b200 as nbk
    from nbkode_b200 import distributed as nd
    from nbkode_b200 import ensembles

    # one process per GPU: keep the process (and the pinned buffers it is about to allocate) on the GPU's NUMA node
    numa_bound = world > 1 and nd.bind_to_gpu_numa(local_rank)

    # ONE ensemble, sharded by trajectory: every rank draws the same B_total rows and keeps its contiguous slice
    B_total = args.batch
    lo, hi = nd.shard_bounds(B_total, rank, world)
    y0_all, p_all = ensembles.lorenz_ensemble(B_total, seed=0)
    y0_np, p_np = np.ascontiguousarray(y0_all[lo:hi]), np.ascontiguousarray(p_all[lo:hi])
    B = hi - lo
    ring_n, l2_note = l2_policy(-(-B_total // world))
    # host inputs live in pinned memory (e2e arm); device-resident copies for the kernel-only arm
    y0_host = torch.from_numpy(y0_np).pin_memory()
    p_host = torch.from_numpy(p_np).pin_memory()
    inputs = [(y0_host.to(dev), p_host.to(dev)) for _ in range(ring_n)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peak_tf, _ = nbk.fp64_peak(local_rank, 20000)

    def kernel_arm(inputs, steps, warmup):
        """Inputs resident in HBM; nothing but the path itself is queued inside the timed region."""
        import collections

        live = collections.deque(maxlen=len(inputs))  # keeps the last R states alive: the allocator hands out other memory
        # Warm-up: at least `warmup` untimed passes, at least WARMUP_MIN_S seconds, and on until three consecutive passes
        # are within 3 % of each other (at most 3 s); the count actually done is reported as "warmup".  The warm-up loop
        # has to have the SAME object lifetimes as the timed loop: an earlier version dropped the result of a warm-up pass
        # at once while the timed loop keeps it bound until the next pass returns -- the second timed pass then needed one
        # more 24 MB output block, and that cudaMalloc, issued while the previous pass's kernel occupied the GPU, took
        # anywhere from 0.8 to 140 ms (per-step CUDA events showed a single 45-140 ms "kernel" among 6.7 ms ones, always at
        # timed step 1; NBK_BENCH_DEBUG=1 prints per-step device and host times and the allocator's cudaMalloc count).
        wu_ms, t_wu, i = [], time.perf_counter(), 0
        while True:
            s = nbk.RungeKutta45("lorenz", 0.0, *inputs[i % len(inputs)], rtol=RTOL, atol=ATOL, device=local_rank)
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            t, y = s.run([T_END])  # (bound to the same names as in the timed loop: the result of the previous pass stays alive
            w1.record()            # while the next one is allocated, so the second output block exists before the clock starts)
            live.append(s)
            torch.cuda.synchronize(dev)
            wu_ms.append(w0.elapsed_time(w1))
            i += 1
            settled = len(wu_ms) >= 3 and max(wu_ms[-3:]) <= 1.03 * min(wu_ms[-3:])
            elapsed = time.perf_counter() - t_wu
            if (i >= warmup and settled and elapsed >= WARMUP_MIN_S) or elapsed > 3.0:
                break
        kernel_arm.warmup_done = i
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        run_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        gc_off = no_gc()
        gc_off.__enter__()
        barrier()
        ev[0].record()
        host_marks, alloc0 = [], torch.cuda.memory_stats(dev).get("num_device_alloc")
        for i in range(steps):
            h0 = time.perf_counter()
            s = nbk.RungeKutta45("lorenz", 0.0, *inputs[(i + kernel_arm.warmup_done) % len(inputs)], rtol=RTOL, atol=ATOL, device=local_rank)
            h1 = time.perf_counter()
            e0, e1 = run_ev[i]
            e0.record()
            t, y = s.run([T_END])
            e1.record()
            h2 = time.perf_counter()
            live.append(s)
            host_marks.append((round(1e3 * (h1 - h0), 3), round(1e3 * (h2 - h1), 3), round(1e3 * (time.perf_counter() - h2), 3)))
        ev[1].record()
        barrier()
        gc_off.__exit__()
        s.raise_for_status()  # device-resident results are checked lazily: once, outside the timed region
        # every step integrates the same ensemble (deterministic kernels): count once, after the clock stopped
        acc = float(s._nacc.sum().item())
        att = acc + float(s._nrej.sum().item())
        per_step = [a.elapsed_time(b) for a, b in run_ev]
        if os.environ.get("NBK_BENCH_DEBUG"):
            print("run-kernel ms per step:", [round(x, 3) for x in per_step], "host ms (ctor, run, append):", host_marks,
                  "device allocs", alloc0, "->", torch.cuda.memory_stats(dev).get("num_device_alloc"), file=sys.stderr)
        return s, y, ev[0].elapsed_time(ev[1]), sum(per_step), acc * steps, att * steps

    # ------------------------------------------------------------ kernel-only arm
    sampler = ClockSampler(local_rank)
    sampler.start()
    s, y_dev, ms_total, ms_run, steps_acc, attempts = kernel_arm(inputs, args.steps, max(args.warmup, 3))
    warmup_done = kernel_arm.warmup_done
    clocks = sampler.finish()
    launch = s.last_launch()
    # device-timed, max over ranks; work summed over ranks (the driver computes efficiency itself)
    (ms_total, ms_run), (total_steps, total_attempts) = nd.reduce_max_sum([ms_total, ms_run], [steps_acc, attempts], device=dev)
    value = total_steps / (ms_total * 1e-3)
    y_gpu = y_dev.cpu().numpy() if cpu_legs else None
    nacc_gpu = s._nacc.cpu().numpy() if cpu_legs else None
    del s, y_dev

    # ------------------------------------------------------------ e2e arm (pinned host buffers in and out)
    # Every pass copies its inputs host->device and its results (final states, step counts) device->host.
    # Consecutive passes alternate between two CUDA streams, so the copies of one pass overlap the
    # integrator kernel of the other -- the way a user would feed a sequence of ensembles.
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def e2e_pass(i):
        with torch.cuda.stream(streams[i % 2]):
            s = nbk.RungeKutta45("lorenz", 0.0, y0_host, p_host, rtol=RTOL, atol=ATOL, device=local_rank)
            t, y = s.run([T_END], wait=False)       # pinned host tensor, filled asynchronously
            nacc, nrej = s.counts(wait=False)
        return s, y, nacc

    def e2e_finish(item):
        s, y, nacc = item
        s.synchronize()                              # results are now in host memory (and checked)
        return int(nacc.sum()), y, nacc

    # warm-up with the same two-deep pattern (pinned result buffers of two passes are alive at once)
    # and the same variable lifetimes as the timed loop: the pinned-host caching allocator then already owns every
    # buffer the steady state needs (a fresh cudaHostAlloc of 24 MB costs ~14 ms and synchronises the device)
    # ... and on until the pipeline has settled (three consecutive passes within 5 % of each other, at most 3 s): the
    # first passes also pay for the pinned buffers and, on some boxes, a slow first few hundred milliseconds (see kernel_arm)
    pending, wall, t_wu, i = None, [], time.perf_counter(), 0
    while True:
        t_p = time.perf_counter()
        item = e2e_pass(i)
        if pending is not None:
            n_i, y, nacc = e2e_finish(pending)
        pending = item
        wall.append(time.perf_counter() - t_p)
        i += 1
        settled = len(wall) >= 4 and max(wall[-3:]) <= 1.05 * min(wall[-3:])
        if (i >= max(args.warmup, 3) + 3 and settled) or time.perf_counter() - t_wu > 3.0:
            break
    n_i, y, nacc = e2e_finish(pending)
    # no stale reference may keep a warm-up solver (416 MB of device state) alive into the timed region: its blocks would be
    # missing from the stream's pool and the first timed pass would cudaMalloc nine new ones (measured: 3.5-35 ms)
    item = pending = None
    barrier()
    dbg = os.environ.get("NBK_BENCH_DEBUG")
    hstat = (lambda: dict(torch.cuda.host_memory_stats()).get("num_host_alloc")) if hasattr(torch.cuda, "host_memory_stats") else (lambda: None)
    h_before, d_before, marks = hstat(), torch.cuda.memory_stats(dev).get("num_device_alloc"), []
    import gc

    gc.collect()
    gc.disable()  # (a generation-2 collection inside a 35 ms timed region is a 10 ms outlier)
    t0 = time.perf_counter()
    e2e_steps, pending = 0, None
    for i in range(args.steps):
        item = e2e_pass(i)
        if pending is not None:
            n_i, y, nacc = e2e_finish(pending)
            e2e_steps += n_i
        pending = item
        item = None
        marks.append(time.perf_counter())
    n_i, y, nacc = e2e_finish(pending)
    e2e_steps += n_i
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    gc.enable()
    if dbg:
        print("e2e warm-up passes:", len(wall), "timed pass marks (ms):", [round(1e3 * (m - t0), 2) for m in marks], "total",
              round(1e3 * e2e_s, 2), "host allocs", h_before, "->", hstat(), "device allocs", d_before, "->",
              torch.cuda.memory_stats(dev).get("num_device_alloc"), file=sys.stderr)
    h2d = y0_host.numel() * 8 + p_host.numel() * 8
    d2h = y.numel() * 8 + 2 * nacc.numel() * 8
    # latency of ONE pass, nothing overlapped: host buffers -> result in host memory, wall clock
    lat = []
    for i in range(max(3, min(args.steps, 10))):
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        e2e_finish(e2e_pass(i))
        lat.append(1e3 * (time.perf_counter() - t1))
    (e2e_s, lat_ms), (e2e_total,) = nd.reduce_max_sum([e2e_s, float(np.median(lat))], [float(e2e_steps)], device=dev)
    e2e_value = e2e_total / e2e_s

    # ------------------------------------------------------------ weak-scaling figure of the same run (N > 1)
    extra = {}
    if not args.no_extra:
        # The same passes (device-resident inputs) issued alternately on two streams: the tail of one pass -- the last
        # trajectories of a shard run with most lanes idle -- overlaps the head of the next.  This is the throughput of a
        # SEQUENCE of ensembles, not the time of one; `value` above stays the single-stream figure.
        t2, acc2, n2, err2 = 0.0, 0.0, max(args.steps, 10), None
        try:  # (no collective inside: a failing rank must not leave the others waiting)
            st2 = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
            keep = []

            def two_stream(n_pass):
                for i in range(n_pass):
                    with torch.cuda.stream(st2[i % 2]):
                        s2 = nbk.RungeKutta45("lorenz", 0.0, *inputs[i % len(inputs)], rtol=RTOL, atol=ATOL, device=local_rank)
                        s2.run([T_END])
                    keep.append(s2)
                    del keep[:-4]

            two_stream(6)
            torch.cuda.synchronize(dev)
            with no_gc():
                t2 = time.perf_counter()
                two_stream(n2)
                torch.cuda.synchronize(dev)
                t2 = time.perf_counter() - t2
            keep[-1].raise_for_status()
            acc2 = float(keep[-1]._nacc.sum().item())
            del keep[:]
        except Exception as exc:  # noqa: BLE001
            err2 = f"{type(exc).__name__}: {exc}"[:300]
        (t2, bad2), (acc2,) = nd.reduce_max_sum([t2, 1.0 if err2 else 0.0], [acc2], device=dev)
        if bad2 or t2 <= 0:
            extra["two_stream_throughput"] = {"error": err2 or "failed on another rank"}
        else:
            extra["two_stream_throughput"] = {
                "value": acc2 * n2 / t2, "unit": UNIT, "ms_per_step": 1e3 * t2 / n2, "passes": n2,
                "note": "device-resident inputs, consecutive passes alternate between two CUDA streams (wall clock around the "
                        "whole sequence, max over ranks): one pass's tail overlaps the next pass's head"}
    if world > 1 and not args.no_extra:
        yw, pw = ensembles.lorenz_ensemble(B_TOTAL, seed=rank)
        w_in = [(torch.from_numpy(yw).to(dev), torch.from_numpy(pw).to(dev))]
        _s, _y, w_ms, w_run, w_steps, _att = kernel_arm(w_in, args.steps, 3)
        (w_ms,), (w_steps,) = nd.reduce_max_sum([w_ms], [w_steps], device=dev)
        extra["weak_scaling"] = {"value": w_steps / (w_ms * 1e-3), "unit": UNIT, "ms_per_step": w_ms / args.steps,
                                 "trajectories_per_gpu": B_TOTAL, "trajectories_total": B_TOTAL * world,
                                 "note": "every rank integrates its own 1M-trajectory ensemble (seed = rank)"}
        del _s, _y, w_in
        # BASELINE configs[4] "sharded across 8 GPUs": the 65 536 reaction-diffusion grids of n = 2048, one contiguous slice
        # per rank, one CTA per trajectory (K3); strong scaling, no collective in the data path
        B5, n5 = 65_536, 2048
        ms5, acc5, att5, err5 = 0.0, 0.0, 0.0, None
        try:  # (no collective inside: a failing rank must not leave the others waiting)
            lo5, hi5 = nd.shard_bounds(B5, rank, world)
            y5, p5 = ensembles.rd1d_ensemble(B5, n=n5, seed=3)
            y5d, p5d = torch.from_numpy(np.ascontiguousarray(y5[lo5:hi5])).to(dev), torch.from_numpy(np.ascontiguousarray(p5[lo5:hi5])).to(dev)
            del y5, p5
            best = None
            for _ in range(2):
                s5 = nbk.RungeKutta45("rd1d", 0.0, y5d, p5d, rtol=RTOL, atol=ATOL, device=local_rank)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with no_gc():
                    e0.record()
                    s5.run([1.0])
                    e1.record()
                    torch.cuda.synchronize(dev)
                t5 = e0.elapsed_time(e1)
                best = t5 if best is None else min(best, t5)
            s5.raise_for_status()
            ms5, acc5 = best, float(s5._nacc.sum().item())
            att5 = acc5 + float(s5._nrej.sum().item())
            del s5, y5d, p5d
        except Exception as exc:  # noqa: BLE001  (a secondary config must not take the headline line down)
            err5 = f"{type(exc).__name__}: {exc}"[:300]
        (ms5, bad5), (acc5, att5) = nd.reduce_max_sum([ms5, 1.0 if err5 else 0.0], [acc5, att5], device=dev)
        flop5 = 6 * 8 * n5 + 66 * n5 + 18
        if bad5 or ms5 <= 0:
            extra["c5_strong_scaling"] = {"error": err5 or "failed on another rank"}
        else:
            extra["c5_strong_scaling"] = {
                "config": f"C5 reaction-diffusion n={n5} x {B5} in total, RungeKutta45, t in [0,1], {B5 // world} trajectories per GPU",
                "kernel": "nbk_run_rk45_rd1d_cta (K3)", "ms": ms5, "accepted_steps": acc5, "steps_per_s": acc5 / ms5 * 1e3,
                "fp64_tflops_per_gpu": att5 / world * flop5 / ms5 * 1e3 / 1e12, "fp64_frac_per_gpu": att5 / world * flop5 / ms5 * 1e3 / 1e12 / peak_tf,
                "note": "device-timed per rank (CUDA events), max over ranks; strong scaling of BASELINE configs[4]"}

    if rank == 0:
        achieved_tf = (total_attempts / world) * FLOP_PER_ATTEMPT / (ms_run * 1e-3) / 1e12  # per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup_done,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, B_total, l2_note),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "host_numa_bound": bool(numa_bound),
                    "note": "throughput of back-to-back passes alternating between two streams (copies overlap the other "
                            "pass's kernel); one_pass_latency_ms = one pass alone, host buffers in -> results in host memory",
                    "one_pass_latency_ms": lat_ms},
            "gpu_launches": 2 * args.steps,
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": ncu_traffic_bytes() if B == B_TOTAL else None,
                "traffic_unit": "bytes of DRAM traffic per launch (ncu, 1M trajectories); algorithmic: ~100 B in + ~330 B out per trajectory",
                "note": "FP64 vector (DFMA) pipe; achieved = attempted steps x 264 algorithmic FLOP / persistent-kernel time "
                        "(CUDA events); peak = nbk_fp64_peak DFMA chain measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "kernel": "nbk_run_rk45_lorenz", "kernel_ms_per_step": ms_run / args.steps,
                "kernel_share_of_step": ms_run / ms_total, "attempts_per_step": total_attempts / world / args.steps,
                "launch": launch,
            },
            "accepted_steps_per_pass": total_steps / args.steps,
        }
        if cpu_legs:
            ref_c, line["cpu_baseline_port"] = cpu_port(y0_np, p_np)
            line["parity"] = {"vs_c_port_full_ensemble": parity_stats(y_gpu, nacc_gpu, ref_c["ys"], ref_c["n_accepted"])}
            if pool is not None:
                ref_l, live = cpu_live(pool, y0_np, p_np)
                pool.close()
                n = len(ref_l["ys"])
                line["cpu_baseline"] = live
                line["cpu_baseline_numba"] = live
                line["parity"]["vs_live_nbkode_prefix"] = parity_stats(y_gpu[:n], nacc_gpu[:n], ref_l["ys"], ref_l["n_accepted"])
                line["parity"]["c_port_vs_live_nbkode_prefix"] = parity_stats(
                    ref_c["ys"][:n], ref_c["n_accepted"][:n], ref_l["ys"], ref_l["n_accepted"])
            else:
                line["cpu_baseline"] = line["cpu_baseline_port"]
                line["cpu_baseline_numba"] = {"unavailable": pool_note}
        if world == 1 and not args.no_extra:
            extra["configs"] = extra_configs(nbk, torch, dev, peak_tf)
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
