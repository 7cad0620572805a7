#!/usr/bin/env python
"""bench.py -- headline benchmark of BASELINE.json: accepted RK45 steps/s on the Lorenz-63
ensemble (configs[1]: 1M trajectories per GPU, random y0, sigma/rho/beta sweep, rtol=1e-6,
atol=1e-9, t in [0, 10], FP64).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over the ensemble: constructor kernel (f0 + initial
step size) and the persistent RK45 kernel to t = 10.  Under torchrun every rank integrates its
own 1M-trajectory ensemble (weak scaling, no data-path collective).

Printed JSON line (rank 0): see the contract in the task statement; additionally
  roofline     FP64 vector pipe: algorithmic FLOPs (264 per attempted step)
               / run-kernel time, against the DFMA peak measured in this run (nbk_fp64_peak)
  cpu_baseline the C oracle (oracle/nbk_oracle.c, a port of the reference's algorithm) on all
               host cores over a bounded sample of the same ensemble
  e2e          the same metric through the public API with pinned HOST buffers in and out
`--impl reference` times the oracle port alone (the reference is pure Python + numba and
cannot travel to the GPU box).
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

B_PER_GPU = 1_000_000
T_END = 10.0
RTOL, ATOL = 1e-6, 1e-9
FLOP_PER_ATTEMPT = 264  # 6R + 66n + 18 with R = 8, n = 3
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


def workload_config(n_gpus):
    return {
        "workload": "Lorenz-63 ensemble, RungeKutta45 rtol=1e-6 atol=1e-9, t in [0,10], run([10.0])",
        "trajectories_per_gpu": B_PER_GPU,
        "trajectories_total": B_PER_GPU * n_gpus,
        "parallelism": f"trajectory sharding x{n_gpus}, no collective",
        "l2": "working set (inputs 48 MB + state 344 MB per GPU) exceeds the 126 MB L2; no explicit flush",
    }


# ----------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs, with the
    profiling guide's recipe: `nvidia-smi --query-gpu=... -lms 200` in a separate process (NVML
    calls from a thread of this process were measured to delay kernel launches by ~2 ms/step)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index, self.period_ms, self.proc = index, period_ms, None

    def start(self):
        import shutil
        import subprocess

        exe = shutil.which("nvidia-smi")
        if exe is None:
            return
        try:
            self.proc = subprocess.Popen(
                [exe, f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 f"-lms={self.period_ms}"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
def cpu_oracle_throughput(y0, p, budget_s=12.0, threads=None):
    """C oracle (port of the reference algorithm) on all host cores over a bounded sample."""
    from oracle import nbk_oracle_c as oc

    threads = threads or os.cpu_count() or 1
    probe = min(len(y0), 4096 * max(1, threads // 4))
    t0 = time.perf_counter()
    r = oc.run_ensemble("RungeKutta45", "lorenz", y0[:probe], p[:probe], [T_END], rtol=RTOL, atol=ATOL, nthreads=threads, dense=False)
    dt = time.perf_counter() - t0
    rate = r["n_accepted"].sum() / dt
    n = int(min(len(y0), max(probe, budget_s * probe / max(dt, 1e-6))))
    t0 = time.perf_counter()
    r = oc.run_ensemble("RungeKutta45", "lorenz", y0[:n], p[:n], [T_END], rtol=RTOL, atol=ATOL, nthreads=threads, dense=False)
    dt = time.perf_counter() - t0
    return {
        "value": float(r["n_accepted"].sum() / dt), "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"first {n} trajectories of the rank-0 ensemble, {dt:.2f} s wall, C oracle (oracle/nbk_oracle.c) with pthreads",
        "probe_rate": float(rate),
    }


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from oracle import nbk_oracle_c as oc
    from oracle import nbk_oracle_np as onp

    threads = os.cpu_count() or 1
    y0, p = onp.lorenz_ensemble(B_PER_GPU, seed=0)
    # bounded sample per step: ~3 s of CPU work
    probe = 8192
    t0 = time.perf_counter()
    oc.run_ensemble("RungeKutta45", "lorenz", y0[:probe], p[:probe], [T_END], rtol=RTOL, atol=ATOL, nthreads=threads, dense=False)
    dt = time.perf_counter() - t0
    n = int(min(B_PER_GPU, max(probe, 3.0 * probe / dt)))
    total_steps, times = 0, []
    for it in range(args.warmup + args.steps):
        lo = (it * n) % max(1, B_PER_GPU - n)
        t0 = time.perf_counter()
        r = oc.run_ensemble("RungeKutta45", "lorenz", y0[lo:lo + n], p[lo:lo + n], [T_END], rtol=RTOL, atol=ATOL,
                            nthreads=threads, dense=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            total_steps += int(r["n_accepted"].sum())
    value = total_steps / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {**workload_config(args.gpus), "sample_trajectories_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} trajectories per step, C oracle port of nbkode's RK45 path (pthreads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

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
    from oracle import nbk_oracle_np as onp  # input generator only (SURVEY.md 8(d) distribution)

    # one process per GPU: keep the process (and the pinned buffers it is about to allocate) on the GPU's NUMA node
    from nbkode_b200 import distributed as nd

    numa_bound = world > 1 and nd.bind_to_gpu_numa(local_rank)

    B = args.batch
    y0_np, p_np = onp.lorenz_ensemble(B, seed=rank)
    # host inputs live in pinned memory (e2e arm); device-resident copies for the kernel-only arm
    y0_host = torch.from_numpy(y0_np).pin_memory()
    p_host = torch.from_numpy(p_np).pin_memory()
    y0_dev, p_dev = y0_host.to(dev), p_host.to(dev)

    def one_pass(y0, p):
        s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=RTOL, atol=ATOL, device=local_rank)
        return s, s.run([T_END])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    peak_tf, _ = nbk.fp64_peak(local_rank, 20000)

    # ------------------------------------------------------------ kernel-only arm (inputs resident in HBM)
    for _ in range(max(args.warmup, 3)):
        s, (t, y) = one_pass(y0_dev, p_dev)
    launch = s.last_launch()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    run_ev = []
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        # nothing but the path itself is queued inside the timed region: no host sync, no reductions
        s = nbk.RungeKutta45("lorenz", 0.0, y0_dev, p_dev, rtol=RTOL, atol=ATOL, device=local_rank)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t, y = s.run([T_END])
        e1.record()
        run_ev.append((e0, e1))
    ev[1].record()
    barrier()
    s.raise_for_status()  # device-resident results are checked lazily: do it once, outside the timed region
    # every step integrates the same ensemble (deterministic kernels): count once, after the clock stopped
    steps_acc = s._nacc.sum() * args.steps
    attempts = (s._nacc.sum() + s._nrej.sum()) * args.steps
    clocks = sampler.finish()
    ms_total = ev[0].elapsed_time(ev[1])
    ms_run = sum(a.elapsed_time(b) for a, b in run_ev)
    from nbkode_b200 import distributed as nd

    # device-timed, max over ranks; work summed over ranks (the driver computes efficiency itself)
    (ms_total, ms_run), (total_steps, total_attempts) = nd.reduce_max_sum(
        [ms_total, ms_run], [float(steps_acc.item()), float(attempts.item())], device=dev)
    value = total_steps / (ms_total * 1e-3)

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
    pending = None
    for i in range(max(args.warmup, 3) + 3):
        item = e2e_pass(i)
        if pending is not None:
            n_i, y, nacc = e2e_finish(pending)
        pending = item
    n_i, y, nacc = e2e_finish(pending)
    barrier()
    t0 = time.perf_counter()
    e2e_steps, pending = 0, None
    dbg = []
    for i in range(args.steps):
        ta = time.perf_counter()
        item = e2e_pass(i)
        tb = time.perf_counter()
        if pending is not None:
            n_i, y, nacc = e2e_finish(pending)
            e2e_steps += n_i
        dbg.append((1e3 * (tb - ta), 1e3 * (time.perf_counter() - tb)))
        pending = item
    n_i, y, nacc = e2e_finish(pending)
    e2e_steps += n_i
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if os.environ.get("NBK_BENCH_DEBUG"):
        print("e2e (enqueue, finish) ms per pass:", " ".join(f"({a:.2f},{b:.2f})" for a, b in dbg), file=sys.stderr)
    h2d = y0_host.numel() * 8 + p_host.numel() * 8
    d2h = y.numel() * 8 + 2 * nacc.numel() * 8
    (e2e_s,), (e2e_total,) = nd.reduce_max_sum([e2e_s], [float(e2e_steps)], device=dev)
    e2e_value = e2e_total / e2e_s
    if rank == 0:
        achieved_tf = (total_attempts / world) * FLOP_PER_ATTEMPT / (ms_run * 1e-3) / 1e12  # per GPU
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {**workload_config(world), "trajectories_per_gpu": B},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "host_numa_bound": bool(numa_bound)},
            "gpu_launches": 2 * args.steps,
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": ncu_traffic_bytes(), "traffic_unit": "bytes of DRAM traffic per launch (ncu); algorithmic: ~100 B in + ~330 B out per trajectory",
                "note": "FP64 vector (DFMA) pipe; achieved = attempted steps x 264 algorithmic FLOP / persistent-kernel time "
                        "(CUDA events); peak = nbk_fp64_peak DFMA chain measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "kernel": "nbk_run_rk45_lorenz", "kernel_ms_per_step": ms_run / args.steps,
                "kernel_share_of_step": ms_run / ms_total, "attempts_per_step": total_attempts / world / args.steps,
                "launch": launch,
            },
            "accepted_steps_per_pass": total_steps / world / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_throughput(y0_np, p_np)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
