"""Diagnostic (GPU box): how much does clock sampling perturb back-to-back passes?"""
import os, sys, time, subprocess, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "numbakit-ode_b200")]
import numpy as np, torch
import nbkode_b200 as nbk
from oracle import nbk_oracle_np as onp
B = 1_000_000
y0, p = onp.lorenz_ensemble(B, 0)
y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
def passes(n=40):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9); s.run([10.0])
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
passes(5)
print("no sampler            %.3f ms/pass" % passes())
F_ALL = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
for fields, lms in ((F_ALL, 100), (F_ALL, 250), (F_ALL, 1000), ("clocks.sm", 100), ("clocks.sm,clocks_event_reasons.active", 100)):
    pr = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={fields}", "--format=csv,noheader,nounits", f"-lms={lms}"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    time.sleep(0.3)
    ms = passes()
    pr.terminate(); out, _ = pr.communicate()
    print(f"nvidia-smi -lms={lms:<5d} {len(fields.split(',')):d} fields  {ms:.3f} ms/pass  ({len(out.splitlines())} samples)")
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for what in ("clock", "clock+reasons"):
    stop = threading.Event(); n = [0]
    def loop():
        while not stop.is_set():
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            if what != "clock": pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            n[0] += 1; time.sleep(0.05)
    th = threading.Thread(target=loop, daemon=True); th.start(); time.sleep(0.2)
    ms = passes(); stop.set(); th.join()
    print(f"pynvml thread {what:14s} {ms:.3f} ms/pass ({n[0]} samples)")
t0 = time.perf_counter(); pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); t1 = time.perf_counter(); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h); t2 = time.perf_counter(); pynvml.nvmlDeviceGetPowerUsage(h); t3 = time.perf_counter()
print("nvml call latency idle: clock %.3f ms, reasons %.3f ms, power %.3f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
