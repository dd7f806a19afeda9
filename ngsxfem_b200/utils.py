"""xfem.utils on the accelerated path (python/utils.py of the reference): element neighbourhoods (`AdjacencyMatrix`,
`AddNeighborhood` :9-54), `IsEqual` (:5-7) and the named level sets of the demos (`LevelsetExamples` :58-72).

The reference multiplies a marker vector with B B^T, B = the element x (vertex | edge | face) incidence matrix; here one layer
is the same two steps on the device's marker kernels: the entities of the marked elements (xgpu_dofs_of_elements on the
incidence table), then the elements holding a marked entity (xgpu_elements_with_marked)."""
import math

import torch

from . import ops
from ._lib import XgpuError
from .cutinfo import BitArray


def IsEqual(A, B):
    """python/utils.py:5-7: elementwise equal"""
    return (A & ~B).NumSet() == 0 and (B & ~A).NumSet() == 0


class AdjacencyMatrix:
    """AdjacencyMatrix(mesh, neighbortype="vertex" | "edge" | "face") (python/utils.py:9-29): elements are neighbours if they
    share a vertex / an edge / a face (H1 / HCurl / HDiv lowest-order dofs there: in 2D the "faces" are the edges).  Holds the
    element -> entity incidence table; AddNeighborhood applies it."""

    def __init__(self, mesh, neighbortype="vertex"):
        if neighbortype == "vertex":
            table, n = mesh.elements, mesh.nv
        elif neighbortype == "edge":
            if mesh.dim < 2:
                raise XgpuError("AdjacencyMatrix: edge neighbours need a 2D / 3D mesh")
            table, n = mesh.edges()[1], mesh.nedge
        elif neighbortype == "face":
            if mesh.dim < 2:
                raise XgpuError("AdjacencyMatrix: face neighbours need a 2D / 3D mesh")
            table, n = mesh.facets()[1], mesh.nfacet
        else:
            raise XgpuError("Unknown neighbor type: " + str(neighbortype) + "\n allowed are vertex, face and edge.")
        self.mesh, self.neighbortype, self.table, self.nent = mesh, neighbortype, table.contiguous(), int(n)


def AddNeighborhood(marker, adja, layers=1, inplace=False):
    """python/utils.py:31-54: the marker extended by `layers` layers of neighbours w.r.t. `adja`; inplace refills `marker`"""
    if not isinstance(adja, AdjacencyMatrix):
        raise XgpuError("AddNeighborhood: adja is an AdjacencyMatrix(mesh, neighbortype)")
    ne = adja.table.shape[0]
    if len(marker) != ne:
        raise XgpuError("AddNeighborhood: the marker is not a BitArray over the elements of the mesh")
    bits = marker.bits
    for _ in range(int(layers)):
        ent = ops.dofs_of_elements(adja.table, bits, adja.nent)
        bits = ops.elements_with_marked(adja.table, ent)
    if int(layers) < 1:           # adja^0 is not formed by the reference either: mw stays the zero vector
        bits = ops.new_bits(ne, marker.bits.device)
    if inplace:
        marker.bits.copy_(bits)
        return marker
    return BitArray(ne, words=bits)


def _sq(t):
    return t * t


# the named level sets of the reference's demos (python/utils.py:58-72) as functions of the coordinates (torch)
LevelsetExamples = {
    # Dziuk, Elliott, Finite element methods for surface PDEs, Acta Numerica 2013, pp. 373-374
    "cheese": lambda x, y, z: torch.sqrt(_sq(x * x - 1) + _sq(y * y - 1) + _sq(z * z - 1) + _sq(x * x + y * y - 4)
                                         + _sq(x * x + z * z - 4) + _sq(y * y + z * z - 4)) - 4,
    # ibid., pp. 318-319
    "dziukelliott": lambda x, y, z: torch.sqrt(0.25 * x * x + y * y + 4.0 * z * z / _sq(1 + 0.5 * torch.sin(math.pi * x))) - 1.0,
    # Dziuk, Finite elements for the Beltrami operator on arbitrary surfaces
    "dziuk88": lambda x, y, z: torch.sqrt(_sq(x - z * z) + y * y + z * z) - 1.0,
    "sphere": lambda x, y, z: torch.sqrt(x * x + y * y + z * z) - 1.0,
    # Grande, Reusken, SINUM 2016
    "torus": lambda x, y, z: torch.sqrt(z * z + _sq(torch.sqrt(x * x + y * y) - 1.0)) - 0.6,
    # Lehrenfeld, CMAME 2016
    "gyroid": lambda x, y, z: torch.cos(math.pi * x) * torch.sin(math.pi * y) + torch.cos(math.pi * y) * torch.sin(math.pi * z)
    + torch.cos(math.pi * z) * torch.sin(math.pi * x),
}

# the boxes the reference meshes for them (python/utils.py:74-81): (lower corner, upper corner)
BoundingBoxes = {
    "cheese": ((-2.5, -2.5, -2.5), (2.5, 2.5, 2.5)),
    "dziukelliott": ((-2.5, -1.5, -1.5), (2.5, 1.5, 1.5)),
    "dziuk88": ((-2, -2, -2), (2, 2, 2)),
    "sphere": ((-1.5, -1.5, -1.5), (1.5, 1.5, 1.5)),
    "torus": ((-2, -2, -2), (2, 2, 2)),
    "gyroid": ((-1, -1, -1), (1, 1, 1)),
}
