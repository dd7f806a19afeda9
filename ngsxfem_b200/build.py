"""Build libxgpu.so (hand-written sm_100a CUDA kernels + the C ABI of include/xgpu.h) in-tree.

    python -m ngsxfem_b200.build [--force]

nvcc cross-compiles without a GPU; the .so lands next to this file so that it travels with the
repository snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libxgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# Geometry kernels take discrete decisions (which sub-simplices survive a clamped re-evaluation,
# straightcutrule.cpp:548-556) from floating-point values at the 1e-14 threshold; they are compiled
# without FMA contraction so that those decisions are taken from the same IEEE operations as the
# CPU restatement (gcc -ffp-contract=off).  The dense contraction kernels keep FMA.
NO_FMA = {"cutquad.cu", "cutquad_ext.cu", "classify.cu", "deform.cu"}
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "--expt-relaxed-constexpr"]


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(src, seen=None):
    """newest mtime among the files `src` includes (recursively, quoted includes only)"""
    seen = set() if seen is None else seen
    t = os.path.getmtime(src)
    with open(src) as f:
        text = f.read()
    for inc in _INC.findall(text):
        path = os.path.normpath(os.path.join(os.path.dirname(src), inc))
        if path in seen or not os.path.exists(path):
            continue
        seen.add(path)
        t = max(t, _deps(path, seen))
    return t


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < _deps(src):
            jobs.append((src, obj))

    # longest translation units first (measured seconds on this image), so that the pool's tail is short
    heavy = {"facetpatch.cu": 136, "elmat_st_d3.cu": 90, "facetpatch_dmma_33.cu": 74, "facetpatch_dmma_34.cu": 72, "elmat_st_d2.cu": 65,
             "elmat_ms_3.cu": 62, "elmat.cu": 52, "elmat_st_d1.cu": 35, "elmat_dmma_3.cu": 31, "cutquad_ext.cu": 29}
    jobs.sort(key=lambda j: -heavy.get(os.path.basename(j[0]), 10))

    def cc(job):
        src, obj = job
        extra = ["-fmad=false"] if os.path.basename(src) in NO_FMA else []
        cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    failed = False
    with ThreadPoolExecutor(max_workers=8) as ex:
        for src, rc, out in ex.map(cc, jobs):
            if rc != 0 or verbose:
                sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
            failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
