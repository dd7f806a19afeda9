"""DRAM traffic per pipeline phase from an `ncu --set full` report of ONE config-2 step
(tools/ncu_step.py): dram__bytes_read.sum + dram__bytes_write.sum of the phase's kernels, summed.
Writes profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) and prints a table.

usage: ncu_traffic.py report.ncu-rep n [out.json]
"""
import csv
import json
import subprocess
import sys

PHASE = [("k_classify", "classify"), ("k_cut_ratio", "classify"),
         ("k_count_dofs", "restrict_plan"), ("k_active_", "restrict_plan"), ("k_gather_fill", "restrict_plan"), ("k_adj_", "restrict_plan"),
         ("k_dofs_of_els", "restrict"), ("k_bits_rank", "restrict"), ("k_gather_rows", "restrict"),
         ("k_elmat_cut", "elmat_cut"), ("k_affine_coefs", "affine_coefs"),
         ("k_cq_", "elmat"), ("k_elmat", "elmat"),
         ("k_rows_small_pattern", "csr_count_pass"), ("k_rows_warp_pattern", "csr_count_pass"),
         ("k_rows_small_values", "csr_rows_fused"), ("k_rows_warp_values", "csr_rows_fused"),
         ("k_pattern", "csr_pattern_pass"), ("k_compact", "csr_pattern"), ("k_scatter", "csr_scatter")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main():
    rep, n = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else "profiles/ncu_traffic.json"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    col = {k: hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * UNIT[units[col[k]]]

    # one step = the launches from one k_classify to the next (the capture window may start in the middle of a step)
    body = rows[2:]
    starts = [i for i, r in enumerate(body) if "k_classify" in r[ik]]
    if starts:
        body = body[starts[0]:(starts[1] if len(starts) > 1 else len(body))]
    phases, kernels = {}, {}
    for r in body:
        name = r[ik]
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        ms = val(r, "gpu__time_duration.sum")
        k = kernels.setdefault(short, [0, 0.0, 0.0])
        k[0] += 1; k[1] += b; k[2] += ms
        for pref, ph in PHASE:
            if pref in short:
                phases[ph] = phases.get(ph, 0.0) + b
                break
    json.dump({"n": n, "source": rep.split("/")[-1], "note": "dram__bytes_read.sum + dram__bytes_write.sum per step, summed over the phase's kernels (cold cache per replay)",
               "phases": {k: int(v) for k, v in phases.items()},
               "kernels": {k: {"launches": v[0], "dram_bytes": int(v[1]), "ms_under_ncu": round(v[2], 4)} for k, v in kernels.items()}},
              open(out, "w"), indent=1)
    for k, v in sorted(kernels.items(), key=lambda x: -x[1][2]):
        print(f"{k:40s} n={v[0]:3d} dram={v[1] / 1e9:8.3f} GB  t={v[2]:8.3f} ms")
    for k, v in phases.items():
        print(f"phase {k:20s} {v / 1e9:8.3f} GB")


if __name__ == "__main__":
    main()
