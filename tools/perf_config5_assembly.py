"""BASELINE config 5 at full size: space-time CutFEM on one time slab — moving sphere, P1-in-space x P2-in-time level set
and FE space on 6 367 248 tets (n=102): space-time cut rules on the elements with a NEG contribution, space-time element
matrices (mass + dt + stiffness, 12 x 12), sparse assembly.  Size-independent checks: 1^T M 1 = space-time volume of the
discrete NEG domain (sum of the rule weights x |det F|), K 1 = 0, (dt-form) 1 = 0.   python tools/perf_config5_assembly.py [n]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

NEG = 0
n5 = int(sys.argv[1]) if len(sys.argv) > 1 else 102
dev = "cuda"


def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:58s} {ms:9.3f} ms", flush=True)
    return out, ms


mesh = xf.MakeStructured3DMesh(nx=n5, device=dev)
tfe = xf.ScalarTimeFE(2)
r = 0.3
ls = xf.SpaceTimeInterpolateToP1(
    lambda x, y, z, t: torch.sqrt((x - .5) ** 2 + (y - .5 - .1 * math.sin(2 * math.pi * t)) ** 2 + (z - .5) ** 2) - r, mesh, tfe)
print(f"config 5: n={n5} ne={mesh.ne}, P1 x P2 space-time FE (12 dofs / tet), time_order 4", flush=True)
(off, q, tv, w), t_rule_all = timed("space-time NEG rule on ALL elements (order 2 x 4)",
                                    lambda: ops.cutquad_st(3, None, mesh.elements, mesh.points, ls.vec, 2, NEG, 4, 2))
vol = xf.Integrate({"levelset": ls, "domain_type": NEG}, mesh=mesh, cf=1.0, order=2, time_order=4)
del off, q, tv, w
ci, t_ci = timed("CutInfo(mesh, lset_st, time_order=0) on the slab", lambda: xf.CutInfo(mesh, ls, time_order=0))
active = ci.GetElementsOfType(xf.HASNEG)
nact = active.NumSet()
V = xf.SpaceTimeFESpace(xf.Restrict(xf.H1(mesh, order=1), active), tfe)
u, v = V.TnT()


def assemble(form):
    a = xf.BilinearForm(V)
    a += form * xf.dCut(ls, NEG, time_order=4, definedonelements=active)
    return a.Assemble()


a, ta = timed("Assemble (u v + dt(u) v + grad u.grad v), NEG, one slab", lambda: assemble(u * v + xf.dt(u) * v + xf.grad(u) * xf.grad(v)))
print(f"elements with a NEG contribution {nact}, ndof {V.ndof}, nnz {a.mat.nze}: {nact / ta / 1e3:.2f} M matrices/s", flush=True)
l_A = active.indices()
(rule, t_rule) = timed("   of which: space-time rule on the active elements", lambda: ops.cutquad_st(3, l_A, mesh.elements, mesh.points, ls.vec, 2, NEG, 4, 2))
off, q, tv, w = rule
terms = [(0, 0, 1.0, 0), (3, 0, 1.0, 0), (1, 1, 1.0, 0)]
(mats, t_el) = timed("   of which: xgpu_elmat_st (3 terms)", lambda: ops.elmat_st(3, 1, 2, l_A, mesh.elements, mesh.points, off, q, tv, w, terms))
npts = w.numel()
byt = nact * (16 + 96 + 144 * 8) + npts * 40
print(f"   points {npts} ({npts / nact:.1f} / element); element matrices: {byt / t_el / 1e6:.1f} GB/s algorithmic "
      f"({nact / t_el / 1e3:.1f} M matrices/s)")
del mats, rule, off, q, tv, w
m = assemble(u * v)
k = assemble(xf.grad(u) * xf.grad(v))
d = assemble(xf.dt(u) * v)
one = torch.ones(V.ndof, dtype=torch.float64, device=dev)
tot = float(one @ (m.mat * one))
k1 = float((k.mat * one).abs().max())
d1 = float((d.mat * one).abs().max())
ball = 4.0 / 3.0 * math.pi * r ** 3
print(f"1^T M 1 = {tot:.15f}  space-time NEG volume (rule) = {vol:.15f}  rel diff {abs(tot - vol) / vol:.2e};  max|K 1| = {k1:.2e}  "
      f"max|D 1| = {d1:.2e};  ball volume {ball:.6f}")
assert abs(tot - vol) < 1e-11 * vol and k1 < 1e-9 and d1 < 1e-9
assert abs(vol - ball) < 2e-2 * ball
free, total = torch.cuda.mem_get_info()
print(f"mem used {(total - free) / 1e9:.1f} GB")
