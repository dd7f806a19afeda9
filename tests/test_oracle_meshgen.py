"""oracle/meshgen.py (numpy input producers of the CPU arm of bench.py) against the product's mesh / dof tables:
identical numbering, so both arms of the bench work on the same inputs."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import sphere_levelset
from oracle import meshgen, xoracle as xo


@pytest.mark.parametrize("n", [3, 6])
def test_structured_mesh_and_dof_tables_match(n):
    mesh = xf.MakeStructured3DMesh(nx=n, device="cpu")
    pts, els = meshgen.structured_tets(n)
    assert np.array_equal(els, mesh.elements.numpy()) and np.allclose(pts, mesh.points.numpy(), rtol=0, atol=0)
    for p in (1, 2):
        V = xf.H1(mesh, order=p)
        eld, ndof = meshgen.h1_dofs(els, len(pts), p)
        assert ndof == V.ndof and np.array_equal(eld, V.el_dofs.numpy())
    phi = meshgen.sphere_levelset(pts, (.5, .5, .5), 0.4)
    ref = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4).numpy()
    # torch's CPU sqrt is not correctly rounded (1 ulp off for some arguments); numpy's and CUDA's are IEEE
    assert np.allclose(phi, ref, rtol=0, atol=2.3e-16)


def test_graph_does_not_depend_on_the_thread_count():
    rng = np.random.default_rng(3)
    table = rng.integers(0, 200, size=(500, 6)).astype(np.int32)
    for r in range(len(table)):   # distinct dofs per row
        table[r] = rng.permutation(200)[:6]
    n0 = xo.num_threads()
    xo.set_num_threads(1)
    rp1, c1 = xo.graph(200, table)
    xo.set_num_threads(max(2, n0))
    rp2, c2 = xo.graph(200, table)
    xo.set_num_threads(n0)
    assert np.array_equal(rp1, rp2) and np.array_equal(c1, c2)
    assert xo.num_threads() == n0 >= 1
