"""Per-source-line view of an ncu report captured with --import-source on: stall samples, executed instructions and the
dominant stall reasons of the hottest lines. usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
        out.append((int(d["# Samples"]), int(d["Instructions Executed"]), fname, int(r[0]), r[1].strip()[:90], stalls))
tot_s = sum(o[0] for o in out) or 1
tot_i = sum(o[1] for o in out) or 1
print(f"# {rep}: {tot_s} samples, {tot_i} warp instructions")
agg = {}
for s, i, f, *_ in out:
    a = agg.setdefault(f, [0, 0]); a[0] += s; a[1] += i
for f, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"## {f}: {100*s/tot_s:.1f}% of samples, {100*i/tot_i:.1f}% of instructions")
for s, i, f, ln, src, st in sorted(out, reverse=True)[:top]:
    main = ", ".join(f"{k} {100*v/max(s,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln}  [{main}]  {src}")
