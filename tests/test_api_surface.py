"""The names of the reference's Python surface that belong to the accelerated path (python/__init__.py, python/mlset.py,
python/utils.py, python/lsetcurv.py, python/lset_smoothblend.py and the pybind layer xfem/python_xfem.cpp,
cutint/python_cutint.cpp, spacetime/python_spacetime.cpp) exist under the same names; the out-of-scope subsystems
(XFEM spaces, aggregates, multigrid, shifted evaluation, drawing — SURVEY.md §2) do not pretend to."""
import ngsxfem_b200 as xf

IN_SCOPE = {
    "xf": ["NEG", "POS", "IF", "NO", "CDOM_NEG", "CDOM_POS", "UNCUT", "CDOM_IF", "HASNEG", "HASPOS", "ANY", "HAS", "VOL", "BND",
           "CutInfo", "MultiLevelsetCutInfo", "InterpolateToP1", "GetFacetsWithNeighborTypes", "GetElementsWithNeighborFacets",
           "GetDofsOfElements", "GetElementsWithSharedVertex", "Restrict", "RestrictedBilinearForm", "BilinearForm", "LinearForm",
           "dCut", "dFacetPatch", "dxtref", "dmesh", "Integrate", "SymbolicBFI", "SymbolicLFI", "dn", "dt", "dtref", "fix_tref",
           "fix_t", "ScalarTimeFE", "SpaceTimeFESpace", "ReferenceTimeVariable", "tref", "NoDeformation", "IndicatorCF",
           "CutRatioGF", "kappa", "IsCut", "LevelSetMeshAdaptation", "ProjectShift", "DomainTypeArray", "TensorUnion",
           "TensorIntersection", "CoefficientFunction", "specialcf", "Normalize", "grad", "x", "y", "z"],
    "xf.mlset": ["DomainTypeArray", "TensorUnion", "TensorIntersection"],
    "xf.utils": ["IsEqual", "AdjacencyMatrix", "AddNeighborhood", "LevelsetExamples", "BoundingBoxes"],
    "xf.lsetcurv": ["LevelSetMeshAdaptation", "ProjectShift"],
    "xf.lset_smoothblend": ["LevelSet_SmoothBlending", "BLEND_FIXED_WIDTH", "BLEND_AUTODETECT"],
}
OUT_OF_SCOPE = ["XFESpace", "XStdFESpace", "AggEmbedding", "ElementAggregation", "CutMG", "DrawDC", "shifted_eval", "pos", "neg",
                "extend", "pos_grad", "neg_grad", "extend_grad", "SFESpace", "ddtref", "RefineAtLevelSet", "P2Prolongation"]


def test_in_scope_names_exist():
    missing = []
    for mod, names in IN_SCOPE.items():
        obj = eval(mod)
        missing += [f"{mod}.{n}" for n in names if not hasattr(obj, n)]
    assert not missing, missing


def test_out_of_scope_names_are_absent():
    present = [n for n in OUT_OF_SCOPE if hasattr(xf, n)]
    assert not present, present


def test_domain_type_values_match_the_reference_enums():
    # utils/ngsxstd.hpp:16-30: DOMAIN_TYPE {NEG = 0, POS = 1, IF = 2}, COMBINED_DOMAIN_TYPE bit masks 0..7
    assert (xf.NEG, xf.POS, xf.IF) == (0, 1, 2)
    assert [c.mask for c in (xf.NO, xf.CDOM_NEG, xf.CDOM_POS, xf.UNCUT, xf.CDOM_IF, xf.HASNEG, xf.HASPOS, xf.ANY)] == list(range(8))
    assert xf.HAS(xf.NEG) is xf.HASNEG and xf.HAS(xf.POS) is xf.HASPOS and xf.HAS(xf.IF) is xf.CDOM_IF
