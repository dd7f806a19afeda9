"""stall samples and executed warp instructions of a kernel in an .ncu-rep, summed over buckets of consecutive SASS instructions
usage: ncu_regions.py report.ncu-rep kernel-regex [bucket]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if hdr is not None:
            break
        hdr = "next"
    elif hdr == "next":
        hdr = r
    elif hdr is not None and r:
        data.append(r)
iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot_ex = sum(int(r[iex]) for r in data); tot_s = sum(int(r[ismp]) for r in data)
print("instructions", len(data), "executed", tot_ex, "samples", tot_s)
for b in range(0, len(data), bucket):
    ch = data[b:b + bucket]
    ex = sum(int(r[iex]) for r in ch); s = sum(int(r[ismp]) for r in ch)
    ops = {}
    for r in ch:
        o = r[isrc].strip().split()[0] if r[isrc].strip() else "?"
        if o.startswith("@"):
            o = r[isrc].strip().split()[1]
        ops[o.split(".")[0]] = ops.get(o.split(".")[0], 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print(f"{b:6d} exec {100.0 * ex / tot_ex:6.2f}%  samples {100.0 * s / max(1, tot_s):6.2f}%  {top}")
