"""Count the opcode mix of the hot attempt region of a kernel's SASS (largest branch-free-ish
span containing the most DFMA).  usage: sass_hot.py file.sass"""
import re, sys, collections
lines = open(sys.argv[1]).read().split("\n")
ins = []
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4})\*/\s+((?:@!?U?P\d+\s+)?)([A-Z0-9_.]+)(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(3), m.group(2) + m.group(3) + m.group(4)))
# split at BSSY/BSYNC/BRA targets: regions between control instructions; merge forward-only spans
ctrl = [i for i, (a, op, t) in enumerate(ins) if op.split(".")[0] in ("BRA", "EXIT", "BSSY", "WARPSYNC", "RET")]
best = None
bounds = [0] + ctrl + [len(ins)]
# greedy: find window [i, j) maximizing DFMA count where no backward BRA and no EXIT inside, allow BSSY/fwd BRA
def score(i, j): return sum(1 for k in range(i, j) if ins[k][1].startswith(("DFMA", "DMUL", "DADD")))
spans = []
start = 0
for k, (a, op, t) in enumerate(ins):
    base = op.split(".")[0]
    back = False
    if base == "BRA":
        m = re.search(r"0x([0-9a-f]+)", t)
        back = bool(m) and int(m.group(1), 16) <= a
    if base in ("EXIT", "WARPSYNC") or back or (base == "BRA" and "@" not in t and not back and score(start, k) < 50):
        spans.append((start, k + 1)); start = k + 1
spans.append((start, len(ins)))
i, j = max(spans, key=lambda s: score(*s))
# trim to the part from the first FP64 op region start
ops = collections.Counter(op.split(".")[0] for a, op, t in ins[i:j])
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"region {ins[i][0]:#x}..{ins[j-1][0]:#x}: {j-i} instr, FP64-pipe {fp64}")
print(dict(ops.most_common()))
