"""per-source-line instruction counts / stall samples of one kernel from an .ncu-rep
usage: ncu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout.splitlines()
# first kernel instance only: locate header line
start = next(i for i, l in enumerate(out) if l.startswith('"Line No"'))
end = next((i for i in range(start + 1, len(out)) if out[i].startswith('"Kernel Name"') or out[i].startswith('"File Name"')), len(out))
rows = list(csv.reader(out[start:end]))
hdr = rows[0]
iL, iS, iI, iSm = hdr.index("Line No"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[1:]:
    if len(r) <= max(iI, iSm) or r[iL] == "":
        continue  # SASS rows below a source line
    try:
        data.append((int(r[iI] or 0), int(r[iSm] or 0), r[iL], r[iS]))
    except (ValueError, IndexError):
        pass
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total inst {ti:,}  samples {ts:,}")
for inst, smp, ln, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{ln:>5s} inst {100*inst/max(ti,1):5.1f}%  stall-samples {100*smp/max(ts,1):5.1f}%  {src.strip()[:110]}")
