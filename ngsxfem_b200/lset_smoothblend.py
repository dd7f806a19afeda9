"""xfem.lset_smoothblend on the accelerated path (python/lset_smoothblend.py of the reference): the smooth blending function b
that switches a mesh deformation off away from the interface — 0 where |lset| <= w, 1 where |lset| >= 2 w, in between the
polynomial pi_s of the given order in s = (|lset| - w) / w:  2^(k-1) s^k for s < 1/2, 1 - 2^(k-1) (1 - s)^k above
(python/lset_smoothblend.py:60-62; arXiv 2311.02348).  `CF` is a function of the coordinates, the form
LevelSetMeshAdaptation.CalcDeformation(levelset, blending=...) takes.

BLEND_FIXED_WIDTH: options {"lset": function of the coordinates, "order": int, "width": float}.  BLEND_AUTODETECT estimates
the width from the cut elements of a time slab (CutInfo(time_order=...) + IntegrationPointExtrema, :63-81) and is not provided."""
from enum import Enum

import torch

from ._lib import XgpuError


class SmoothBlendingType(Enum):
    BLEND_AUTODETECT = 1
    BLEND_FIXED_WIDTH = 2


BLEND_AUTODETECT = SmoothBlendingType.BLEND_AUTODETECT
BLEND_FIXED_WIDTH = SmoothBlendingType.BLEND_FIXED_WIDTH


def blend_1d_polyn(s, order):
    """pi_s on [0, 1]: IfPos(0.5 - s, 2^(order-1) s^order, 1 - 2^(order-1) (1 - s)^order)"""
    c = 2.0 ** (order - 1)
    return torch.where(0.5 - s > 0, c * s ** order, 1.0 - c * (1.0 - s) ** order)


class LevelSet_SmoothBlending:
    """LevelSet_SmoothBlending(BLEND_FIXED_WIDTH, {"lset": fn, "order": k, "width": w}).CF(x, y[, z]) -> b"""

    def __init__(self, sb_type, option_dict):
        if sb_type not in (BLEND_AUTODETECT, BLEND_FIXED_WIDTH):
            raise XgpuError("Illegal first argument in LevelSet_SmoothBlending")
        if not isinstance(option_dict, dict):
            raise XgpuError("LevelSet_SmoothBlending needs a dictionary when called with BLEND_AUTODETECT or BLEND_FIXED_WIDTH")
        if sb_type == BLEND_AUTODETECT:
            raise XgpuError("BLEND_AUTODETECT (width from the cut elements of a time slab) is not provided: give the width "
                            "(BLEND_FIXED_WIDTH)")
        self.smooth_blend_type = sb_type
        try:
            self.lset, self.order, self.width = option_dict["lset"], int(option_dict["order"]), float(option_dict["width"])
        except KeyError as e:
            raise XgpuError(f"LevelSet_SmoothBlending(BLEND_FIXED_WIDTH): missing option {e}")
        if not callable(self.lset) or self.order < 1 or not self.width > 0:
            raise XgpuError("LevelSet_SmoothBlending: lset is a function of the coordinates, order >= 1, width > 0")
        self.Calc_blend_CF()

    def Calc_blend_CF(self):
        w, k, lset = self.width, self.order, self.lset

        def CF(*coords):
            a = torch.abs(torch.as_tensor(lset(*coords), dtype=torch.float64))
            mid = blend_1d_polyn((a - w) / w, k)
            return torch.where(w - a > 0, torch.zeros_like(a), torch.where(a - 2 * w > 0, torch.ones_like(a), mid))

        self.CF = CF
