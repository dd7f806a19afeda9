"""The C-ABI library loads on a CPU-only host and exports exactly the entry points include/xgpu.h declares;
compute entry points fail loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "xgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with python -m ngsxfem_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = header_symbols()
    assert len(decl) >= 40
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/xgpu.h but not exported"
    # the ctypes signature table covers the header exactly
    assert sorted(_lib.exported_symbols()) == decl
    # and nothing else with the xgpu_ prefix is exported with C linkage
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if re.search(r" T xgpu_", l))
    assert exported == decl


def test_host_only_entry_points_work_without_gpu():
    lib = _lib.load()
    assert lib.xgpu_version().startswith(b"xgpu")
    assert [lib.xgpu_h1_ndof(3, p) for p in (1, 2, 3, 4)] == [4, 10, 20, 35]
    sh, dsh = xf.ops.h1_shapes_host(3, 2, [0, 1, 2, 3], [0.2, 0.3, 0.1])
    assert abs(float(sh[:4].sum()) - 1.0) < 1e-15


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a host without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(xf.XgpuError):
        _lib.Ctx()
    mesh = xf.MakeStructured3DMesh(nx=2)
    with pytest.raises(xf.XgpuError):
        xf.CutInfo(mesh, xf.InterpolateToP1(lambda x, y, z: x - 0.3, mesh))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ngsxfem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"
