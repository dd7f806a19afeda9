"""Periodic facet identification in GetFacetsWithNeighborTypes (xfem/cutinfo.cpp:509-518): the oracle restates the lookup
(GetPeriodicFacet, the facet with the smaller number takes the partner's element); the product hands the kernel a facet table
that already carries that neighbour (Mesh.facet_els_periodic).  Closed-form answer on a structured periodic mesh, then
oracle == table transform on random markers (CPU), then the device result through the public function (GPU)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import meshes
from oracle import xoracle as xo
from util import unpack_bits


def pack(mask):
    return np.packbits(np.asarray(mask, dtype=bool), bitorder="little")


def test_identification_of_a_structured_mesh():
    nx, ny = 4, 3
    m = meshes.MakeStructured2DMesh(nx=nx, ny=ny, periodic_x=True)
    ents, _, fe = m.facets()
    part = m.periodic_facets().numpy()
    nf = len(part)
    ident = np.flatnonzero(part != np.arange(nf))
    assert len(ident) == 2 * ny and np.array_equal(part[part[ident]], ident)
    P = m.points.numpy()
    for f in ident:
        assert fe[f, 1] < 0
        c, c2 = P[ents[f].numpy()].mean(axis=0), P[ents[part[f]].numpy()].mean(axis=0)
        assert abs(abs(c[0] - c2[0]) - 1.0) < 1e-12 and abs(c[1] - c2[1]) < 1e-12
    assert meshes.MakeStructured2DMesh(nx=3, ny=3).periodic_facets() is None
    m3 = meshes.MakeStructured3DMesh(nx=2, ny=3, nz=2, periodic_x=True, periodic_z=True)
    p3 = m3.periodic_facets().numpy()
    assert (p3 != np.arange(len(p3))).sum() == 2 * (2 * 3 * 2) + 2 * (2 * 2 * 3)   # two triangles per boundary cell face


@pytest.mark.parametrize("use_and", [True, False])
def test_first_and_last_column_meet_through_the_identification(use_and):
    nx, ny = 5, 3
    m = meshes.MakeStructured2DMesh(nx=nx, ny=ny, periodic_x=True)
    _, _, fe = m.facets()
    fe = fe.numpy()
    part = m.periodic_facets().numpy()
    cx = m.points[m.elements.long()].mean(dim=1)[:, 0].numpy()
    a, b = pack(cx < 1.0 / nx), pack(cx > 1.0 - 1.0 / nx)
    got = unpack_bits(xo.facets_with_neighbor_types(fe, a, b, False, False, use_and, periodic_facet=part), len(fe))
    plain = unpack_bits(xo.facets_with_neighbor_types(fe, a, b, False, False, use_and), len(fe))
    if use_and:
        # (a_left and b_right) or (a_right and b_left): only across the identified boundary, on the smaller facet of each pair
        assert plain.sum() == 0
        want = (fe[:, 1] < 0) & (part > np.arange(len(fe)))
        assert np.array_equal(got, want) and got.sum() == ny
    else:
        # "or" with boundary values False: a facet of an a- or b-element on either side
        assert np.array_equal(got, plain | ((fe[:, 1] < 0) & (part > np.arange(len(fe)))))


@pytest.mark.parametrize("dim", [2, 3])
def test_oracle_lookup_equals_the_table_transform(dim):
    m = meshes.MakeStructured2DMesh(nx=4, ny=5, periodic_x=True, periodic_y=True) if dim == 2 else \
        meshes.MakeStructured3DMesh(nx=3, ny=2, nz=3, periodic_y=True)
    _, _, fe = m.facets()
    fe2 = m.facet_els_periodic().numpy()
    part = m.periodic_facets().numpy()
    rng = np.random.default_rng(dim)
    for use_and in (True, False):
        for bva, bvb in ((True, True), (False, True), (False, False)):
            a, b = pack(rng.random(m.ne) < 0.4), pack(rng.random(m.ne) < 0.5)
            lhs = xo.facets_with_neighbor_types(fe.numpy(), a, b, bva, bvb, use_and, periodic_facet=part)
            rhs = xo.facets_with_neighbor_types(fe2, a, b, bva, bvb, use_and)
            assert np.array_equal(lhs, rhs)
    assert (fe2 != fe.numpy()).sum() == (part > np.arange(len(part))).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_device_markers_on_a_periodic_mesh_match_oracle(dim):
    dev = "cuda"
    m = xf.MakeStructured2DMesh(nx=6, ny=5, periodic_x=True, device=dev) if dim == 2 else \
        xf.MakeStructured3DMesh(nx=3, ny=4, nz=3, periodic_x=True, periodic_z=True, device=dev)
    _, _, fe = m.facets()
    part = m.periodic_facets().cpu().numpy()
    rng = np.random.default_rng(10 + dim)
    for use_and in (True, False):
        for bva, bvb in ((True, True), (False, False), (True, False)):
            ma, mb = rng.random(m.ne) < 0.4, rng.random(m.ne) < 0.5
            A, B = xf.BitArray.from_numpy(ma, dev), xf.BitArray.from_numpy(mb, dev)
            got = xf.GetFacetsWithNeighborTypes(m, a=A, b=B, bnd_val_a=bva, bnd_val_b=bvb, use_and=use_and)
            want = xo.facets_with_neighbor_types(fe.cpu().numpy(), pack(ma), pack(mb), bva, bvb, use_and, periodic_facet=part)
            assert np.array_equal(got.to_numpy(), unpack_bits(want, m.nfacet))

