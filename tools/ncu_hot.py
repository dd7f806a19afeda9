"""hot SASS instructions of a kernel in an .ncu-rep (source page): address, executed warp instructions, stall samples
usage: ncu_hot.py report.ncu-rep kernel-regex [top]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks[:1]:
    h = b["hdr"]
    ia, isrc, iex, ismp, ithr = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
    tot_ex = sum(int(r[iex]) for r in b["rows"])
    tot_s = sum(int(r[ismp]) for r in b["rows"])
    print(b["name"], "instructions", tot_ex, "samples", tot_s)
    print("--- in program order (instructions executed % / samples % / avg threads)")
    for k, r in enumerate(b["rows"]):
        ex, s = int(r[iex]), int(r[ismp])
        if ex * 100.0 / max(1, tot_ex) >= 0.4 or s * 100.0 / max(1, tot_s) >= 0.6:
            print(f"{k:5d} {100.0 * ex / tot_ex:5.2f}% {100.0 * s / max(1, tot_s):5.2f}% {r[ithr]:>5s}  {r[isrc].strip()[:90]}")
