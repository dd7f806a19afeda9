"""Config-5 kernels once warm, once to profile — the short command ncu wraps (slab CutInfo, space-time rules, element matrices)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 102
dev = "cuda"
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
tfe = xf.ScalarTimeFE(2)
ls = xf.SpaceTimeInterpolateToP1(
    lambda x, y, z, t: torch.sqrt((x - .5) ** 2 + (y - .5 - .1 * math.sin(2 * math.pi * t)) ** 2 + (z - .5) ** 2) - 0.3, mesh, tfe)
terms = [(0, 0, 1.0, 0), (3, 0, 1.0, 0), (1, 1, 1.0, 0)]
for _ in range(2):
    ci = xf.CutInfo(mesh, ls, time_order=0)
    l_A = ci.GetElementsOfType(xf.HASNEG).indices()
    off, q, tv, w = ops.cutquad_st(3, l_A, mesh.elements, mesh.points, ls.vec, 2, 0, 4, 2)
    m = ops.elmat_st(3, 1, 2, l_A, mesh.elements, mesh.points, off, q, tv, w, terms)
torch.cuda.synchronize()
print("active", l_A.numel(), "points", w.numel())
