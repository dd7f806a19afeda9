"""ngsxfem_b200.utils (xfem.utils of the reference, python/utils.py): the golden BitArrays of the reference's own
tests/pytests/test_elementlayers.py:18-40 through AdjacencyMatrix / AddNeighborhood / IsEqual.  On a host without a GPU the four
marker kernels the module calls are replaced by numpy stand-ins with the kernels' contracts (those kernels have their own GPU
parity tests: test_marker_helpers_match_oracle, GetElementsWithSharedVertex); what is checked here is the host logic — which
incidence table a neighbour type means, the layer loop, inplace — against the reference's known answer."""
import json
import os

import numpy as np
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200 import meshes, ops
from ngsxfem_b200.utils import AddNeighborhood, AdjacencyMatrix, IsEqual, LevelsetExamples, BoundingBoxes

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "elementlayers.json")))


def unpack(words, n):
    return np.unpackbits(words.numpy().view(np.uint8), bitorder="little")[:n].astype(bool)


def pack(mask):
    n = len(mask)
    buf = np.zeros(ops.nwords(n) * 4, np.uint8)
    p = np.packbits(np.asarray(mask, dtype=bool), bitorder="little")
    buf[:p.size] = p
    return torch.from_numpy(buf.view(np.int32).copy())


@pytest.fixture
def numpy_marker_kernels(monkeypatch):
    def dofs_of_elements(el_dofs, el_bits, ndof, ctx=None):          # xgpu_dofs_of_elements: entries of the marked rows
        t = el_dofs.numpy()
        m = unpack(el_bits, t.shape[0])
        out = np.zeros(ndof, bool)
        d = t[m].reshape(-1)
        out[d[d >= 0]] = True
        return pack(out)

    def elements_with_marked(table, marks, ctx=None):                # xgpu_elements_with_marked: rows with a marked entry
        t = table.numpy()
        mk = unpack(marks, int(t.max()) + 1)
        return pack(np.array([any(mk[c] for c in row if c >= 0) for row in t]))

    def bitset_op(op, a, b, nbits, out=None, ctx=None):              # xgpu_bitset_op (bits beyond nbits stay clear)
        x, y = unpack(a, nbits), unpack(b, nbits)
        r = {ops.BIT_OR: x | y, ops.BIT_AND: x & y, ops.BIT_ANDNOT: x & ~y, ops.BIT_NOT: ~x, ops.BIT_COPY: x,
             ops.BIT_CLEAR: np.zeros_like(x), ops.BIT_SET: np.ones_like(x)}[op]
        w = pack(r)
        if out is not None:
            out.copy_(w)
            return out
        return w

    def bitset_count(bits, nbits, ctx=None):
        return int(unpack(bits, nbits).sum())

    for name, fn in (("dofs_of_elements", dofs_of_elements), ("elements_with_marked", elements_with_marked),
                     ("bitset_op", bitset_op), ("bitset_count", bitset_count)):
        monkeypatch.setattr(ops, name, fn)


def test_element_layers_golden(numpy_marker_kernels):
    mesh = meshes.MakeStructured2DMesh(nx=7, ny=7)
    # the IF elements of x - 1/2 (tests/pytests/test_elementlayers.py:13-14, 27-28): sign change of the nodal values
    phi = mesh.points[:, 0] - 0.5
    v = phi[mesh.elements.long()]
    is_if = ((v > 0).any(dim=1) & (v < 0).any(dim=1)).numpy()
    ref = {"vertex": np.array([c == "1" for c in GOLD["vertex"]]), "edge": np.array([c == "1" for c in GOLD["edge"]])}
    ref["face"] = ref["edge"]                                        # in 2D the lowest-order HDiv dofs are the edges (:22-23)
    for nbtype in ("vertex", "face", "edge"):
        marker = xf.BitArray(mesh.ne, "cpu", words=pack(is_if))
        adja = AdjacencyMatrix(mesh, nbtype)
        m2 = AddNeighborhood(marker, adja, 1, inplace=False)
        m3 = AddNeighborhood(m2, AdjacencyMatrix(mesh, nbtype), 1, inplace=False)
        m4 = AddNeighborhood(marker, adja, 2, inplace=False)
        assert m2 is not marker and np.array_equal(unpack(marker.bits, mesh.ne), is_if)       # not in place
        assert AddNeighborhood(marker, adja, 2, inplace=True) is marker
        assert IsEqual(m3, m4) and IsEqual(m3, marker) and IsEqual(m4, marker)
        assert not IsEqual(m2, m3)
        assert IsEqual(xf.BitArray(mesh.ne, "cpu", words=pack(ref[nbtype])), marker), nbtype
    with pytest.raises(xf.XgpuError):
        AdjacencyMatrix(mesh, "corner")
    with pytest.raises(xf.XgpuError):
        AddNeighborhood(xf.BitArray(mesh.ne + 1, "cpu"), AdjacencyMatrix(mesh, "vertex"))
    with pytest.raises(xf.XgpuError):
        AddNeighborhood(xf.BitArray(mesh.ne, "cpu"), "vertex")


def test_neighbour_tables_in_3d():
    m = meshes.MakeStructured3DMesh(nx=2)
    assert AdjacencyMatrix(m, "vertex").table.shape == (m.ne, 4) and AdjacencyMatrix(m, "vertex").nent == m.nv
    assert AdjacencyMatrix(m, "edge").table.shape == (m.ne, 6) and AdjacencyMatrix(m, "edge").nent == m.nedge
    assert AdjacencyMatrix(m, "face").table.shape == (m.ne, 4) and AdjacencyMatrix(m, "face").nent == m.nfacet
    with pytest.raises(xf.XgpuError):
        AdjacencyMatrix(meshes.MakeStructured1DMesh(4), "edge")


def test_levelset_examples():
    assert set(LevelsetExamples) == set(BoundingBoxes) == {"cheese", "dziukelliott", "dziuk88", "sphere", "torus", "gyroid"}
    one, zero = torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.0], dtype=torch.float64)
    assert abs(float(LevelsetExamples["sphere"](one, zero, zero))) < 1e-15                      # on the unit sphere
    assert abs(float(LevelsetExamples["torus"](1.6 * one, zero, zero))) < 1e-15                 # outer equator, R = 1, r = 0.6
    assert abs(float(LevelsetExamples["torus"](zero, one, 0.6 * one))) < 1e-15
    assert abs(float(LevelsetExamples["dziuk88"](one, zero, zero))) < 1e-15
    assert abs(float(LevelsetExamples["dziukelliott"](2 * one, zero, zero))) < 1e-15            # sqrt(x^2 / 4) = 1
    assert abs(float(LevelsetExamples["gyroid"](zero, zero, zero))) < 1e-15
    assert abs(float(LevelsetExamples["cheese"](zero, zero, zero)) - (np.sqrt(3 + 48) - 4)) < 1e-14
    # usable as InterpolateToP1 input inside its box
    lo, hi = BoundingBoxes["torus"]
    mesh = meshes.MakeStructured3DMesh(nx=8, mapping=lambda x, y, z: (lo[0] + (hi[0] - lo[0]) * x, lo[1] + (hi[1] - lo[1]) * y, lo[2] + (hi[2] - lo[2]) * z))
    g = xf.InterpolateToP1(LevelsetExamples["torus"], mesh)
    assert g.vec.shape == (mesh.nv,) and float(g.vec.min()) < 0 < float(g.vec.max())
