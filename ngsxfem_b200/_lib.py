"""ctypes loader of libxgpu.so (the C ABI declared in include/xgpu.h).

There is NO fallback: if the CUDA library is missing or a call fails, an exception is raised.
torch is used only as the device allocator / stream provider; tensors are passed as raw pointers.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxgpu.so")

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_intp = C.POINTER(C.c_int)
VP = C.c_void_p


class XgpuError(RuntimeError):
    """raised where the reference raises ngcore::Exception / NgException"""


class Term(C.Structure):
    _fields_ = [("op_trial", C.c_int), ("op_test", C.c_int), ("coef", C.c_double), ("hpow", C.c_int)]


class LTerm(C.Structure):
    _fields_ = [("op_test", C.c_int), ("coef", C.c_double), ("hpow", C.c_int), ("expo", C.c_int * 3)]


class StLTerm(C.Structure):
    _fields_ = [("op_test", C.c_int), ("coef", C.c_double), ("hpow", C.c_int), ("expo", C.c_int * 4)]


class FTerm(C.Structure):
    _fields_ = [("trial_kind", C.c_int), ("trial_side", C.c_int), ("test_kind", C.c_int), ("test_side", C.c_int),
                ("coef", C.c_double), ("hpow", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("api_calls", C.c_int64), ("plans_created", C.c_int64), ("scratch_bytes", C.c_int64),
                ("status_code", C.c_int), ("status_element", C.c_int64)]


class Batch(C.Structure):
    _fields_ = [("table", VP), ("mats", VP), ("nrows", C.c_int64), ("width", C.c_int), ("coefs", VP)]


_SIGS = {
    "xgpu_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
    "xgpu_destroy": (None, [VP]),
    "xgpu_set_stream": (C.c_int, [VP, VP]),
    "xgpu_set_option": (C.c_int, [VP, C.c_char_p, C.c_int]),
    "xgpu_set_deformation": (C.c_int, [VP, C.c_int, VP, VP]),
    "xgpu_sync": (C.c_int, [VP]),
    "xgpu_last_error": (C.c_char_p, [VP]),
    "xgpu_get_status": (C.c_int, [VP, c_intp, c_i64p]),
    "xgpu_clear_status": (C.c_int, [VP]),
    "xgpu_launch_count": (C.c_int64, [VP]),
    "xgpu_get_stats": (C.c_int, [VP, C.POINTER(Stats)]),
    "xgpu_version": (C.c_char_p, []),
    "xgpu_rule_size": (C.c_int, [VP, C.c_int, C.c_int]),
    "xgpu_get_rule": (C.c_int, [VP, C.c_int, C.c_int, c_f64p, c_f64p]),
    "xgpu_set_rule": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, c_f64p, c_f64p]),
    "xgpu_h1_ndof": (C.c_int, [C.c_int, C.c_int]),
    "xgpu_h1_shapes_host": (C.c_int, [C.c_int, C.c_int, c_i32p, c_f64p, c_f64p, c_f64p]),
    "xgpu_classify": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP]),
    "xgpu_bitset_op": (C.c_int, [VP, C.c_int, VP, VP, VP, C.c_int64]),
    "xgpu_bitset_count": (C.c_int, [VP, VP, C.c_int64, c_i64p]),
    "xgpu_bitset_to_list": (C.c_int, [VP, VP, C.c_int64, VP]),
    "xgpu_threshold_elements": (C.c_int, [VP, C.c_int64, VP, C.c_int, C.c_double, VP]),
    "xgpu_classify_ml": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, C.c_int64, C.c_int, c_intp, C.c_int, VP]),
    "xgpu_facets_with_neighbor_types": (C.c_int, [VP, C.c_int64, VP, VP, VP, C.c_int, C.c_int, C.c_int, VP]),
    "xgpu_elements_with_neighbor_facets": (C.c_int, [VP, C.c_int64, VP, VP, C.c_int64, VP]),
    "xgpu_dofs_of_elements": (C.c_int, [VP, C.c_int64, C.c_int, VP, VP, C.c_int64, VP]),
    "xgpu_bitset_rank": (C.c_int, [VP, VP, C.c_int64, VP, c_i64p]),
    "xgpu_gather_rows": (C.c_int, [VP, C.c_int64, VP, C.c_int, VP, VP, VP]),
    "xgpu_cutquad_count": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, C.c_int, C.c_int, VP, c_i64p]),
    "xgpu_cutquad_fill": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int, C.c_int, VP, VP, VP]),
    "xgpu_cutquad_ml_count": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, C.c_int64, C.c_int, c_intp, C.c_int,
                                        C.c_int, VP, c_i64p]),
    "xgpu_cutquad_ml_fill": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int64, C.c_int, c_intp, C.c_int,
                                       C.c_int, VP, VP, VP]),
    "xgpu_cutquad_st_count": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, C.c_int64, C.c_int, C.c_int, C.c_int,
                                        C.c_int, VP, c_i64p]),
    "xgpu_cutquad_st_fill": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       C.c_int, VP, VP, VP, VP]),
    "xgpu_integrate_monomial": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, c_intp, c_f64p]),
    "xgpu_classify_st": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, C.c_int64, C.c_int, C.c_int, VP, VP, VP, VP]),
    "xgpu_rule_element_sums": (C.c_int, [VP, C.c_int64, VP, VP, VP]),
    "xgpu_classify_volumes": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP, VP]),
    "xgpu_integrate_monomial_st": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, c_intp, c_f64p]),
    "xgpu_integrate_monomial_bnd": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, c_intp, c_f64p]),
    "xgpu_elmat_rule": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.POINTER(Term),
                                  C.c_int, C.c_int, VP, VP]),
    "xgpu_elmat_rule_pair": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.POINTER(Term), C.c_int,
                                       VP, VP, VP, C.POINTER(Term), C.c_int, C.c_int, VP, VP]),
    "xgpu_elmat_cut_pair": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int, C.c_int, C.POINTER(Term), C.c_int,
                                      C.c_int, C.c_int, C.POINTER(Term), C.c_int, C.c_int, VP, VP]),
    "xgpu_elmat_std": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int, C.POINTER(Term),
                                 C.c_int, C.c_int, VP, VP]),
    "xgpu_elvec_rule": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.POINTER(LTerm),
                                  C.c_int, C.c_int, VP, VP]),
    "xgpu_elvec_std": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int, C.POINTER(LTerm),
                                 C.c_int, C.c_int, VP, VP]),
    "xgpu_elmat_st": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.POINTER(Term),
                                C.c_int, C.c_int, VP, VP]),
    "xgpu_elvec_st": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.POINTER(StLTerm),
                                C.c_int, C.c_int, VP, VP]),
    "xgpu_facetpatch": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int, C.c_double, C.c_int, VP]),
    "xgpu_facetpatch_set_downscale": (C.c_int, [VP, C.c_double]),
    "xgpu_facetpatch_st": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, C.c_int, C.c_int, C.c_double,
                                     C.c_double, C.c_int, VP]),
    "xgpu_csr_plan": (C.c_int, [VP, C.c_int64, C.POINTER(Batch), C.c_int, C.POINTER(VP), c_i64p]),
    "xgpu_csr_pattern": (C.c_int, [VP, VP, VP, VP]),
    "xgpu_csr_scatter": (C.c_int, [VP, VP, C.POINTER(Batch), C.c_int, VP, VP, VP]),
    "xgpu_adjacency_plan": (C.c_int, [VP, C.c_int64, C.POINTER(Batch), C.c_int, C.POINTER(VP)]),
    "xgpu_vec_scatter": (C.c_int, [VP, VP, C.POINTER(Batch), C.c_int, VP]),
    "xgpu_vec_scatter_mean": (C.c_int, [VP, VP, C.POINTER(Batch), C.c_int, VP, C.c_int, C.c_int]),
    "xgpu_rule_phys": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP]),
    "xgpu_h1_eval_rule": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP]),
    "xgpu_reduce_sum": (C.c_int, [VP, C.c_int64, VP, c_f64p]),
    "xgpu_elements_with_marked": (C.c_int, [VP, C.c_int64, C.c_int, VP, VP, VP]),
    "xgpu_bitset_count_many": (C.c_int, [VP, C.c_int, C.POINTER(VP), c_i64p, c_i64p]),
    "xgpu_band_elements": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, C.c_double, C.c_double, VP]),
    "xgpu_rule_points_phys": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP]),
    "xgpu_h1_grad_points": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP]),
    "xgpu_l2_project": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.c_int64, VP, VP, C.c_int, VP, C.c_int, VP]),
    "xgpu_shift_points": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, C.c_int, VP, VP, C.c_double, VP, c_i64p]),
    "xgpu_csr_assemble": (C.c_int, [VP, VP, C.POINTER(Batch), C.c_int, VP, VP, VP]),
    "xgpu_csr_plan_rows": (C.c_int, [VP, C.c_int64, C.POINTER(Batch), C.c_int, C.POINTER(VP), c_i64p]),
    "xgpu_csr_count_rows": (C.c_int, [VP, VP, c_i64p]),
    "xgpu_restrict_plan": (C.c_int, [VP, C.c_int, VP, C.c_int, C.POINTER(VP), c_i64p, C.c_int64, VP, c_i64p, C.POINTER(VP),
                                    C.POINTER(VP)]),
    "xgpu_csr_assemble_fused": (C.c_int, [VP, VP, C.POINTER(Batch), C.c_int, VP, C.c_int, VP, VP, VP]),
    "xgpu_csr_assemble_fused_rows": (C.c_int, [VP, VP, C.POINTER(Batch), C.c_int, VP, C.c_int, VP, VP, VP, C.c_int64, C.c_int64, C.c_int]),
    "xgpu_affine_coefs": (C.c_int, [VP, C.c_int, C.c_int64, VP, VP, VP, C.POINTER(Term), C.c_int, VP]),
    "xgpu_reftensor": (C.c_int, [VP, C.c_int, C.c_int, C.c_int, C.POINTER(VP), c_intp]),
    "xgpu_plan_destroy": (None, [VP]),
    "xgpu_plan_phase_ms": (C.c_int, [VP, c_f64p]),
    "xgpu_facet_elmat": (C.c_int, [VP, C.c_int, C.c_int, C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.POINTER(FTerm), C.c_int, C.c_int, VP]),
    "xgpu_iface_marks": (C.c_int, [VP, VP, C.c_int64, VP, VP, C.c_int64, VP]),
    "xgpu_iface_rows": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP]),
    "xgpu_iface_index": (C.c_int, [VP, C.c_int64, VP, VP, VP, C.c_int64, VP]),
    "xgpu_l2g_own": (C.c_int, [VP, C.c_int64, C.c_int64, C.c_int64, VP, C.c_int, VP]),
    "xgpu_l2g_ghosts": (C.c_int, [VP, C.c_int64, VP, VP, VP, C.c_int, VP]),
    "xgpu_rows_offsets": (C.c_int, [VP, C.c_int64, VP, VP, VP, c_i64p]),
    "xgpu_rows_pack": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP]),
    "xgpu_rows_add": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP]),
    "xgpu_csr_spmv": (C.c_int, [VP, C.c_int64, VP, VP, VP, VP, VP]),
}

_LIB = None


def exported_symbols():
    return sorted(_SIGS)


def load():
    """dlopen libxgpu.so; raises if it has not been built (python -m ngsxfem_b200.build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise XgpuError(f"{LIB_PATH} not found: build it with `python -m ngsxfem_b200.build` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def ptr(t):
    """device (or host) pointer of a tensor, None -> NULL"""
    if t is None:
        return None
    assert t.is_contiguous()
    return t.data_ptr()


class Ctx:
    """one xgpu_ctx per (device, stream)"""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise XgpuError("no CUDA device: the cut-element assembly path has no CPU fallback")
        self.lib = load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = VP()
        rc = self.lib.xgpu_create(self.device.index, C.byref(h))
        if rc != 0:
            raise XgpuError(f"xgpu_create failed ({rc})")
        self.h = h
        self.use_current_stream()

    def use_current_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != getattr(self, "_stream", None):
            self.check(self.lib.xgpu_set_stream(self.h, VP(s)))
            self._stream = s

    def stats(self):
        """xgpu_get_stats as a dict"""
        st = Stats()
        self.check(self.lib.xgpu_get_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in Stats._fields_}

    def set_option(self, key, value):
        self.check(self.lib.xgpu_set_option(self.h, key.encode(), int(value)))

    def deformation(self, defo):
        """context manager: `with ctx.deformation(d): ...` = mesh.SetDeformation(d) ... mesh.UnsetDeformation()
        (python/__init__.py: the `deformation=` argument of dCut / dFacetPatch / Integrate); d=None: no-op"""
        return _DeformationScope(self, defo)

    def check(self, rc):
        if rc != 0:
            msg = self.lib.xgpu_last_error(self.h)
            err = XgpuError(f"xgpu error {rc}: {msg.decode() if msg else ''}")
            err.code = rc
            raise err

    def status(self):
        code = C.c_int(0)
        el = C.c_int64(-1)
        self.check(self.lib.xgpu_get_status(self.h, C.byref(code), C.byref(el)))
        return code.value, el.value

    def raise_on_geometry_error(self):
        code, el = self.status()
        if code != 0:
            self.check(self.lib.xgpu_clear_status(self.h))
            msgs = {-11: "You tried to cut a simplex with a plain geometry lset function",
                    -12: "Bad length of s_cut!",
                    -13: "Cutting this part of a tetraeder is not implemented yet!"}
            raise XgpuError(f"{msgs.get(code, 'geometry error')} (element {el}, code {code})")

    @property
    def launches(self):
        return self.lib.xgpu_launch_count(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.xgpu_destroy(self.h)
                self.h = None
        except Exception:
            pass


class _DeformationScope:
    def __init__(self, ctx, defo):
        self.ctx, self.defo = ctx, defo

    def __enter__(self):
        d = self.defo
        if d is not None:
            assert d.el_dofs.is_contiguous() and d.vals.is_contiguous()
            self.ctx.check(self.ctx.lib.xgpu_set_deformation(self.ctx.h, int(d.order), d.el_dofs.data_ptr(), d.vals.data_ptr()))
        return self

    def __exit__(self, *a):
        if self.defo is not None:
            self.ctx.check(self.ctx.lib.xgpu_set_deformation(self.ctx.h, 0, None, None))
        return False


_CTX = {}


def default_ctx(device=None):
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    idx = device if isinstance(device, int) else (torch.device(device).index or 0)
    if idx not in _CTX:
        _CTX[idx] = Ctx(idx)
    c = _CTX[idx]
    c.use_current_stream()
    return c
