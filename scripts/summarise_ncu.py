"""Summarise an ncu report (raw page + SASS source page) into a small text file for profiles/."""
import collections, csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second"]
lines = ["# ncu summary of " + rep, ""]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        lines.append(f"{h} [{u}] = {v}")
lines.append("")
lines.append("## warp stall reasons (per issue-active)")
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        lines.append(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} = {float(v):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = [r for r in rows[2:] if len(r) == len(h2)]
iS, iE, iP = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iP]) for r in data)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    toks = r[iS].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    ops[op] += int(r[iE]); samp[op] += int(r[iP])
lines += ["", f"## SASS instruction mix (warp-level instructions executed: {tot}, stall samples: {totS})"]
for op, c in ops.most_common(22):
    lines.append(f"{op:10s} {100*c/tot:6.2f}% of instructions  {100*samp[op]/max(totS,1):6.2f}% of samples")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
