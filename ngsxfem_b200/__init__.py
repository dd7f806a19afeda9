"""ngsxfem_b200 — B200-native cut-element assembly behind the ngsxfem (`xfem`) surface.

Scope: ONE hot path — CutInfo classification -> straight-cut quadrature -> batched element /
ghost-penalty facet-patch matrices -> deterministic CSR assembly — as hand-written sm_100a CUDA
kernels in libxgpu.so (C ABI: include/xgpu.h).  This package is the host-side mirror of the
reference's Python surface for that path (names / argument meaning / error behaviour of
python/__init__.py and the pybind layer).  There is no CPU fallback: without the CUDA library
or a GPU every compute entry point raises XgpuError.

Importing the package does not need a GPU (tables, meshes and symbol checks work on CPU).
"""
from ._lib import XgpuError, LIB_PATH, exported_symbols, load as load_library
from .meshes import Mesh, MakeStructured1DMesh, MakeStructured2DMesh, MakeStructured3DMesh
from .fes import H1, Restrict, Deformation
from .cutinfo import (NEG, POS, IF, VOL, BND, HAS, BitArray, GridFunction, InterpolateToP1, CutInfo,
                      MultiLevelsetCutInfo, GetFacetsWithNeighborTypes, GetElementsWithNeighborFacets,
                      GetDofsOfElements, GetElementsWithSharedVertex, CombinedDomainType)
from .cutinfo import C_NO as NO, C_NEG as CDOM_NEG, C_POS as CDOM_POS, C_UNCUT as UNCUT, C_IF as CDOM_IF, \
    C_HASNEG as HASNEG, C_HASPOS as HASPOS, C_ANY as ANY
from .forms import (dCut, dFacetPatch, grad, Normalize, specialcf, BilinearForm, RestrictedBilinearForm, LinearForm,
                    Integrate, TrialFunction, TestFunction, x, y, z, dn, CoefficientFunction)
from .spacetime import (ScalarTimeFE, SpaceTimeFESpace, SpaceTimeGridFunction, SpaceTimeInterpolateToP1,
                        ReferenceTimeVariable, tref, dtref, dt, fix_tref, dxtref, dmesh)
from .lsetcurv import LevelSetMeshAdaptation, ProjectShift
from . import mlset
from .mlset import DomainTypeArray, TensorUnion, TensorIntersection
from . import utils, lset_smoothblend
from .compat import SymbolicBFI, SymbolicLFI, NoDeformation, IndicatorCF, CutRatioGF, kappa, IsCut, fix_t
from .forms import ElementCF

__all__ = [n for n in dir() if not n.startswith("_")]
