"""Diagnostic (GPU box): time every tuning variant under numbakit-ode_b200/variants/."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sorted(glob.glob(os.path.join(ROOT, "numbakit-ode_b200", "variants", "*.so")))
for lib in [None] + libs:
    env = dict(os.environ)
    if lib: env["NBK_LIB_PATH"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "time_run.py")], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip() or r.stderr[-500:], flush=True)
