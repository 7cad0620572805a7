"""Compile the RK45/Lorenz program with extra -D flags and print registers + the opcode mix of the attempt loop
(diagnostic).  usage: sass_count.py [-DNBK_OPT_ABS ...]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "numbakit-ode_b200", "csrc")
flags = sys.argv[1:]
base = ["-DNBK_BLOCK=128", "-DNBK_MIN_BLOCKS=4"]
if any(f.startswith("-DNBK_BLOCK") for f in flags): base = []
with tempfile.TemporaryDirectory() as td:
    obj = os.path.join(td, "p.o")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
           "-DNBK_METHOD=45", "-DNBK_RHS_T=nbk::rhs_lorenz", "-DNBK_TAG=rk45_lorenz", *base, *flags, "-c",
           os.path.join(CSRC, "nbk_program.cu"), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    info = [l for l in r.stderr.split("\n") if "nbk_run_rk45_lorenz" in l or "registers" in l]
    for i, l in enumerate(r.stderr.split("\n")):
        if "nbk_run_rk45_lorenz" in l and "Compiling" in l:
            print(" ".join(r.stderr.split("\n")[i + 1:i + 4]).strip()[:300])
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "nbk_run_rk45_lorenz", obj], capture_output=True, text=True).stdout
ins = []
for l in sass.split("\n"):
    m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
# the attempt loop: the innermost backward BRA.U (uniform counter loop) containing most DFMA
best = None
for a, t in ins:
    m = re.search(r"BRA(?:\.\w+)*\s+(?:U?P\d+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        n = sum(1 for b, u in ins if lo <= b <= a and u.split()[0].startswith("DFMA"))
        size = a - lo
        if n >= 90 and (best is None or size < best[2]): best = (lo, a, size)
lo, hi, _ = best
body = [(a, re.sub(r"^@!?U?P\d+\s+", "", t)) for a, t in ins if lo <= a <= hi]
c = collections.Counter(t.split()[0].split(".")[0] for a, t in body)
fp = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"loop {lo:#x}..{hi:#x}: {len(body)} instr (incl. emission/retire code inside the loop), FP64-pipe {fp}")
print(dict(c.most_common()))
