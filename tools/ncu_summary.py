"""Turn an .ncu-rep into the committed summaries under profiles/ (json of key metrics, details text,
top stalled instructions).  usage: ncu_summary.py <rep> <out-prefix>"""
import csv, io, json, re, subprocess, sys
rep, prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]
out = {k: d[k] for k in keys if k in d}
for k in sorted(d):
    if re.search(r"smsp__average_warps_issue_stalled.*_per_issue_active", k):
        out[k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")] = d[k]
json.dump(out, open(prefix + ".json", "w"), indent=1)
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(prefix + "_details.txt", "w").write("\n".join(l for l in det.splitlines() if l.strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
sk = ("stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_long_sb", "stall_math", "stall_mio", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait")
with open(prefix + "_top_stalls.txt", "w") as fh:
    fh.write(f"{d.get('Kernel Name', ('?',))[0]}: ncu --set full --import-source on, SASS view; warp-state samples {tot}\n")
    agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in sk if k in ix}
    fh.write("share of samples by state: " + ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])) + "\n\n")
    fh.write("samples  share  warp-instr-executed  avg-threads  dominant-state  instruction\n")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:40]:
        s = int(r[ix["# Samples"]]); main = max(agg, key=lambda k: int(r[ix[k]] or 0))
        fh.write(f"{s:6d} {100*s/tot:5.2f}% {r[ix['Instructions Executed']]:>10} {r[ix['Avg. Threads Executed']]:>5} {main[6:]:18s} {r[ix['Source']].strip()[:80]}\n")
print(json.dumps({k: v[0] for k, v in out.items()}, indent=0)[:1800])
