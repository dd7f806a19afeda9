"""ngsxfem_b200.mlset (DomainTypeArray, TensorUnion, TensorIntersection) against the reference's own known answers:
tests/pytests/test_mlset_module_functionality.py (boundary / complement / operators / ANY expansion / tensor operators, set
equalities) and the compression part of tests/pytests/test_mlset_integration.py:170-253 (z-shaped region of a disc cut by
three planes), the measures coming from the oracle's multi-level-set rule (the product's Integrate needs a GPU)."""
import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import meshes
from ngsxfem_b200.mlset import DomainTypeArray, TensorIntersection, TensorUnion
from oracle import xoracle as xo
from util import interpolate_to_p1, mesh_np

NEG, POS, IF, ANY = xf.NEG, xf.POS, xf.IF, xf.ANY


# ---- tests/pytests/test_mlset_module_functionality.py:19-48
def test_triangle():
    triangle = DomainTypeArray([(NEG, NEG, NEG)])
    assert ~triangle == DomainTypeArray([(NEG, NEG, POS), (NEG, POS, NEG), (NEG, POS, POS), (POS, NEG, NEG), (POS, NEG, POS),
                                         (POS, POS, NEG), (POS, POS, POS)])
    bnd = triangle.Boundary()
    target = DomainTypeArray([(IF, NEG, NEG), (NEG, IF, NEG), (NEG, NEG, IF)])
    assert bnd == target and bnd.codim == 1
    assert (target | bnd) == bnd
    assert (target & bnd) == bnd
    bnd2 = bnd.Boundary()
    assert bnd2 == DomainTypeArray([(IF, IF, NEG), (NEG, IF, IF), (IF, NEG, IF)]) and bnd2.codim == 2
    for i in range(3):
        assert DomainTypeArray(tuple(IF if k == i else NEG for k in range(3))).codim == 1


# ---- :54-102
def test_operators():
    dta1, dta2 = DomainTypeArray((POS, NEG)), DomainTypeArray((POS, POS))
    dta3 = dta1 | dta2
    assert (POS, NEG) in dta3 and (POS, POS) in dta3 and len(dta3) == 2
    assert dta3 == DomainTypeArray([(POS, NEG), (POS, POS)])
    dta3 |= DomainTypeArray([(NEG, POS)])
    assert dta3 == DomainTypeArray([(POS, NEG), (POS, POS), (NEG, POS)])
    assert (dta3 & dta1) == DomainTypeArray([(POS, NEG)])
    dta3 &= dta1 | dta2
    assert dta3 == DomainTypeArray([(POS, NEG), (POS, POS)])
    dta5 = DomainTypeArray([(IF, NEG)]) & DomainTypeArray([(POS, NEG)])
    assert dta5 == DomainTypeArray([(IF, NEG)])
    dta5 &= DomainTypeArray([(POS, IF)])
    assert dta5 == DomainTypeArray([(IF, IF)])
    assert (~dta3) == DomainTypeArray([(NEG, POS), (NEG, NEG)])
    dta6 = DomainTypeArray([(IF, NEG, POS)])
    dta6 &= DomainTypeArray([(POS, IF, POS)])
    assert (~dta6) == DomainTypeArray([(IF, NEG, IF), (POS, IF, IF), (IF, POS, IF), (NEG, IF, IF), (IF, IF, NEG)])
    assert DomainTypeArray([(NEG, NEG, NEG)]).Boundary() == DomainTypeArray([(NEG, IF, NEG), (NEG, NEG, IF), (IF, NEG, NEG)])
    assert DomainTypeArray([(POS, IF, NEG)]).Boundary() == DomainTypeArray([(POS, IF, IF), (IF, IF, NEG)])


# ---- :108-122
def test_any_expansion():
    assert DomainTypeArray((NEG, ANY, ANY)) == DomainTypeArray([(NEG, NEG, NEG), (NEG, NEG, POS), (NEG, POS, NEG), (NEG, POS, POS)])
    assert DomainTypeArray([(NEG, ANY, ANY), (POS, ANY, ANY)]) == DomainTypeArray([(ANY, ANY, ANY)])
    assert DomainTypeArray([(IF, ANY, ANY), (POS, IF, ANY)]) == DomainTypeArray(
        [(IF, NEG, NEG), (IF, NEG, POS), (IF, POS, NEG), (IF, POS, POS), (POS, IF, NEG), (POS, IF, POS)])


# ---- :128-147
def test_tensor_operators():
    dta1, dta2, dta3 = DomainTypeArray((POS,)), DomainTypeArray([(POS, NEG), (POS, POS)]), DomainTypeArray((NEG, NEG, IF))
    assert TensorUnion(dta1, dta2) == DomainTypeArray([(POS, ANY, ANY), (ANY, POS, NEG), (ANY, POS, POS)])
    with pytest.raises(Exception):
        TensorUnion(dta1, dta2, dta3)
    assert TensorIntersection(dta1, dta2) == DomainTypeArray([(POS, POS, NEG), (POS, POS, POS)])
    assert TensorIntersection(dta1, dta2, dta3) == DomainTypeArray([(POS, POS, NEG, NEG, NEG, IF), (POS, POS, POS, NEG, NEG, IF)])


def test_input_checks_and_container_protocol():
    with pytest.raises(TypeError):
        DomainTypeArray([NEG, POS])                        # a list of entries, not of tuples
    with pytest.raises(TypeError):
        DomainTypeArray("NEG")
    with pytest.raises(xf.XgpuError):
        DomainTypeArray([(NEG, IF), (NEG, NEG)])           # different co-dimension
    with pytest.raises(xf.XgpuError):
        DomainTypeArray([(NEG, POS), (NEG, NEG, POS)])     # different length
    with pytest.raises(xf.XgpuError):
        DomainTypeArray([(NEG, 5)])
    with pytest.raises(xf.XgpuError):
        DomainTypeArray((NEG, xf.HASNEG))                  # a combined domain type other than ANY
    with pytest.raises(xf.XgpuError):
        DomainTypeArray((NEG, NEG)) | DomainTypeArray((IF, NEG))   # union of different co-dimension
    with pytest.raises(xf.XgpuError):
        DomainTypeArray((IF, IF, IF)).Boundary()
    with pytest.raises(xf.XgpuError):
        DomainTypeArray((IF, NEG)).GetOuterNormals((None, None))
    d = DomainTypeArray([(NEG, ANY)])
    assert len(d) == 2 and sorted(d) == sorted([(NEG, POS), (NEG, NEG)]) and (NEG, POS) in d and (POS, POS) not in d
    assert d != DomainTypeArray((NEG, POS)) and "Co-dim : 0" in str(d) and "(NEG, POS)" in str(d)
    assert DomainTypeArray([(NEG, POS), (NEG, POS)]).as_list == [(NEG, POS)]      # duplicates collapse
    # the forms layer reads .as_list (cutinfo._normalize_dtlist)
    from ngsxfem_b200.cutinfo import _normalize_dtlist
    assert sorted(_normalize_dtlist(d, 2)) == sorted([(NEG, POS), (NEG, NEG)])
    with pytest.raises(xf.XgpuError):
        _normalize_dtlist(d, 3)


# ---- tests/pytests/test_mlset_integration.py:170-253: compression to the pieces of positive measure
def oracle_measure(mesh):
    pts, els = mesh_np(mesh)

    def measure(lsets, dtt):
        phis = np.stack([l.vec.numpy() for l in lsets], axis=1)         # (nv, M)
        total = 0.0
        for el in range(els.shape[0]):
            T = xo.make_trafo(pts[els[el]], els[el])
            _, w = xo.mlset_cut_rule(T, phis[els[el]], [list(dtt)], 0)
            if len(w):
                J = (pts[els[el]][:2] - pts[els[el]][2]).T
                total += float(w.sum()) * abs(np.linalg.det(J)) if dtt.count(IF) == 0 else float(w.sum())
        return total

    return measure


def test_compression_of_the_z_shaped_disc():
    mesh = meshes.MakeStructured2DMesh(nx=12, ny=12, mapping=lambda x, y: (2.2 * x - 1.1, 2.2 * y - 1.1))
    pts, _ = mesh_np(mesh)
    fns = [lambda x, y: x * x + y * y - 1, lambda x, y: -x - 1 / 3, lambda x, y: x - 1 / 3, lambda x, y: y - 0.5]
    lsets = tuple(xf.GridFunction(mesh, torch.from_numpy(interpolate_to_p1(f, pts))) for f in fns)
    integ = oracle_measure(mesh)
    z1 = DomainTypeArray([(NEG, NEG, NEG, POS), (NEG, POS, NEG, POS), (NEG, POS, NEG, NEG), (NEG, NEG, POS, NEG), (NEG, NEG, POS, POS)])
    z2 = DomainTypeArray([(NEG, ANY, ANY, POS), (NEG, POS, ANY, ANY), (NEG, ANY, POS, ANY)], integrator=integ)
    z3 = DomainTypeArray((NEG, ANY, ANY, ANY), integrator=integ) & ~DomainTypeArray((NEG, NEG, NEG, NEG))
    z4 = DomainTypeArray(z3.as_list, lsets, persistent_compress=False, integrator=integ)
    z5 = DomainTypeArray(z3.as_list, lsets, persistent_compress=True, integrator=integ)
    assert z4.lsets is None and z4.persistent_compress is False
    assert z2 != z1 and z3 != z1                       # before the compression they hold pieces without measure
    z2.Compress(lsets, persistent=True)
    z3.Compress(lsets)
    assert z1 == z2 and z1 == z3 and z1 == z4 and z1 == z5
    z2_bnd, z3_bnd, z5_bnd = z2.Boundary(), z3.Boundary(), z5.Boundary()
    assert z3_bnd != z2_bnd                            # uncompressed boundary: pieces that do not exist geometrically
    z3_bnd.Compress(lsets)
    assert z2_bnd == z5_bnd and z2_bnd == z3_bnd
    assert z2_bnd.persistent_compress is True and all(a is b for a, b in zip(z2_bnd.lsets, lsets))
    assert z3_bnd.persistent_compress is False and z3_bnd.lsets is None
    # operators keep the persistent compression only if both operands carry it, with the same level sets
    assert (z2 | z5).persistent_compress is True and (z2 | z1).persistent_compress is False
    other = tuple(xf.GridFunction(mesh, l.vec.clone()) for l in lsets)
    with pytest.raises(xf.XgpuError):
        z2 | DomainTypeArray(z1.as_list, other, persistent_compress=True, integrator=integ)
    with pytest.raises(xf.XgpuError):
        z1.Compress(lsets[:3])


def test_indicators_and_outer_normals():
    square = DomainTypeArray([(NEG, NEG, NEG, NEG)])
    lsets = [lambda x, y: -y, lambda x, y: x - 1, lambda x, y: y - 1, lambda x, y: -x]          # test_mlset_integration.py:145
    X = torch.tensor([0.5, 0.5, 1.5, -0.2, 0.999], dtype=torch.float64)
    Y = torch.tensor([0.5, 1.5, 0.5, 0.5, 0.001], dtype=torch.float64)
    ind = square.Indicator(lsets)
    fn, gfs = ind.factors[0]
    assert gfs == () and fn(X, Y).tolist() == [1.0, 0.0, 0.0, 0.0, 1.0]
    assert (~square).Indicator(lsets).factors[0][0](X, Y).tolist() == [0.0, 1.0, 1.0, 1.0, 0.0]
    edge = DomainTypeArray([(IF, NEG, NEG, NEG)])                                             # the bottom edge y = 0
    sm = edge.IndicatorSmoothed(lsets, eps=0.05).factors[0][0]
    Xe = torch.tensor([0.5, 0.5, 0.5, 1.2], dtype=torch.float64)
    Ye = torch.tensor([0.01, -0.04, 0.2, 0.0], dtype=torch.float64)
    assert sm(Xe, Ye).tolist() == [1.0, 1.0, 0.0, 0.0]
    assert square.IndicatorSmoothed(lsets).factors[0][0](X, Y).tolist() == [1.0, 0.0, 0.0, 0.0, 1.0]   # codim 0 -> Indicator
    # outer normals: + grad(lset_i) / |.| where the region is NEG in entry i, the negative where it is POS
    mesh = meshes.MakeStructured2DMesh(nx=3, ny=3)
    pts, _ = mesh_np(mesh)
    gf = [xf.GridFunction(mesh, torch.from_numpy(interpolate_to_p1(f, pts))) for f in lsets]
    n = square.GetOuterNormals(gf)
    assert set(n) == set(square.Boundary().as_list) and all(v.sign == 1.0 for v in n.values())
    assert n[(NEG, IF, NEG, NEG)].lset is gf[1]
    outside = DomainTypeArray([(POS, NEG, NEG, NEG)])                                         # below the bottom edge
    n2 = outside.GetOuterNormals(gf)
    assert n2[(IF, NEG, NEG, NEG)].sign == -1.0 and n2[(POS, IF, NEG, NEG)].sign == 1.0
    # the sign ends up in the coefficient of grad(u) * n
    V = xf.H1(mesh, order=1)
    u, v = V.TnT()
    t_plus = (xf.grad(u) * n[(IF, NEG, NEG, NEG)]) * v
    t_minus = (xf.grad(u) * n2[(IF, NEG, NEG, NEG)]) * v
    assert t_plus.terms[0][2] == 1.0 and t_minus.terms[0][2] == -1.0 and t_minus.nlset is gf[0]
    assert (-n[(IF, NEG, NEG, NEG)]).sign == -1.0
