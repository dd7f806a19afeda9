"""ngsxfem_b200.lset_smoothblend against the formula of python/lset_smoothblend.py:60-62 of the reference: b = 0 up to the width,
1 from twice the width, the polynomial pi_s in between (continuous, monotone, b(1.5 w) = 1/2)."""
import pytest
import torch

import ngsxfem_b200 as xf
from ngsxfem_b200.lset_smoothblend import BLEND_AUTODETECT, BLEND_FIXED_WIDTH, LevelSet_SmoothBlending, blend_1d_polyn


@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_fixed_width_blending(order):
    w = 0.2
    sb = LevelSet_SmoothBlending(BLEND_FIXED_WIDTH, {"lset": lambda x, y: x - 0.5, "order": order, "width": w})
    x = torch.linspace(0.0, 1.0, 2001, dtype=torch.float64)
    b = sb.CF(x, torch.zeros_like(x))
    d = (x - 0.5).abs()
    assert torch.all(b[d <= w] == 0) and torch.all(b[d >= 2 * w] == 1)
    assert float(b.min()) >= 0 and float(b.max()) <= 1
    right = x >= 0.5
    assert torch.all(b[right][1:] - b[right][:-1] >= -1e-15)                        # monotone in the distance
    at = lambda dist: float(sb.CF(torch.tensor([0.5 + dist], dtype=torch.float64), torch.zeros(1, dtype=torch.float64)))
    assert abs(at(1.5 * w) - 0.5) < 1e-14                                           # 2^(k-1) (1/2)^k = 1/2
    assert abs(at(1.25 * w) - 2.0 ** (order - 1) * 0.25 ** order) < 1e-14
    assert abs(at(1.75 * w) - (1 - 2.0 ** (order - 1) * 0.25 ** order)) < 1e-14
    assert abs(at(w + 1e-9)) < 1e-7 and abs(at(2 * w - 1e-9) - 1) < 1e-7            # continuous at both ends
    assert abs(at(-1.5 * w) - 0.5) < 1e-14                                          # even in the level set
    s = torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0], dtype=torch.float64)
    p = blend_1d_polyn(s, order)
    assert float(p[0]) == 0 and float(p[-1]) == 1 and abs(float(p[2]) - 0.5) < 1e-15


def test_options_are_checked():
    with pytest.raises(xf.XgpuError):
        LevelSet_SmoothBlending(BLEND_AUTODETECT, {"lset": lambda x, y: x, "order": 2, "mesh": None, "time_order": 1})
    with pytest.raises(xf.XgpuError):
        LevelSet_SmoothBlending(BLEND_FIXED_WIDTH, {"lset": lambda x, y: x, "order": 2})
    with pytest.raises(xf.XgpuError):
        LevelSet_SmoothBlending(BLEND_FIXED_WIDTH, [1, 2])
    with pytest.raises(xf.XgpuError):
        LevelSet_SmoothBlending(3, {})
    with pytest.raises(xf.XgpuError):
        LevelSet_SmoothBlending(BLEND_FIXED_WIDTH, {"lset": 1.0, "order": 2, "width": 0.1})
