"""The NGSolve-side adapter sources (integration/*.cpp) must stay in step with include/xgpu.h: they are
syntax-checked with g++ against the C ABI header and small stand-ins for the NGSolve / ngsxfem declarations they touch
(integration/stubs/ — NGSolve itself is not installable in this image).  A changed signature in xgpu.h breaks this test."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("src", ["gpucutinfo.cpp", "gpurestrictedblf.cpp"])
def test_adapter_sources_compile_against_the_abi_header(src):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    cuda_inc = "/usr/local/cuda/include"
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-DXGPU_ADAPTER_STUBS", "-I", os.path.join(ROOT, "integration", "stubs"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "integration"), "-I", cuda_inc,
           os.path.join(ROOT, "integration", src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
