"""Measured FP64 peaks of the GPU -> profiles/fp64_peak.json (+ a line on stdout).

    python tools/peak_fp64.py --build     # here (nvcc cross-compiles tools/_bin/peak_fp64 for sm_100a; the binary travels)
    python tools/peak_fp64.py             # on the GPU box: DFMA chain, DMMA chain (tools/peak_fp64.cu), cuBLAS DGEMM 8192^3

bench.py reads profiles/fp64_peak.json for the roofline of the contraction kernels (k_elmat_ms: CUDA-core DFMA;
k_elmat_dmma / k_facetpatch_dmma: the FP64 tensor instruction)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "_bin", "peak_fp64")


def build():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-ccbin", "/usr/bin/g++", os.path.join(HERE, "peak_fp64.cu"), "-o", BIN])
    return BIN


def main():
    if "--build" in sys.argv:
        print(build())
        return
    if not os.path.exists(BIN):
        build()
    res = json.loads(subprocess.check_output([BIN]).decode())
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        (a @ b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(); (a @ b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res["dgemm_tflops"] = round(2.0 * n ** 3 / (best * 1e-3) / 1e12, 3)
    res["dgemm_shape"] = f"{n}^3 (cuBLAS through torch.matmul, float64)"
    res["note"] = "best of 5 (chains) / 3 (DGEMM) launches, CUDA events; FP64 has no tcgen05 kind: the tensor instruction is DMMA m8n8k4"
    with open(os.path.join(ROOT, "profiles", "fp64_peak.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
