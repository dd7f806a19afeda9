"""key metrics of every kernel in an .ncu-rep (ncu --set full) as a text table
usage: ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("sm__inst_executed_pipe_fp64.sum", "fp64_inst")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    print(r[ik][:110])
    parts = []
    for key, short in WANT:
        if key in hdr:
            i = hdr.index(key)
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.3f}" if f < 1e6 else f"{f:,.0f}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{(' ' + units[i]) if units[i] and units[i] != '%' else ''}")
    print("   " + "  ".join(parts))
