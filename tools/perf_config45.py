"""Configs 4 and 5 of BASELINE.json at single-GPU scale: timings + size-independent checks (development tool; the
numbers go to profiles/).

config 4: multi-level-set domain cube ∩ 3 planes ∩ sphere, tuple (NEG,NEG,NEG,NEG), H1 order 3, mass + stiffness through
          RestrictedBilinearForm.Assemble.  Check: 1^T M 1 = |domain| from the order-0 multi-level-set rule; K 1 = 0.
config 5: space-time cut rules, moving sphere c(t) = (1/2, 1/2 + 0.1 sin 2 pi t, 1/2), r = 0.3, P1 in space x P2 in time
          (nodes 0, 1/2, 1), time_order 4, space order 2.  Check: the space-time measure of the NEG region is the
          time integral of the (P1-interface) ball volume: close to 4/3 pi r^3; NEG + POS = 1 exactly.
usage: perf_config45.py [n4=64] [n5=102]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 102
dev = torch.device("cuda", 0)
NEG, POS, IF = 0, 1, 2


def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {ms:9.3f} ms", flush=True)
    return out, ms


# ------------------------------------------------------------------ config 4
mesh = xf.MakeStructured3DMesh(nx=n4, device=dev)
x, y, z = (mesh.points[:, i] for i in range(3))
fns = [x + 0.31 * y - 0.83, -0.27 * x + y + 0.11 * z - 0.79, 0.13 * x - 0.21 * y + z - 0.81,
       torch.sqrt((x - .4) ** 2 + (y - .45) ** 2 + (z - .4) ** 2) - 0.37]
lsets = [xf.GridFunction(mesh, torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f).contiguous()) for f in fns]
dts = (NEG, NEG, NEG, NEG)
print(f"config 4: n={n4} ne={mesh.ne} order 3", flush=True)
mlci = xf.MultiLevelsetCutInfo(mesh, lsets)
active, _ = timed("MultiLevelsetCutInfo.GetElementsWithContribution", lambda: xf.MultiLevelsetCutInfo(mesh, lsets).GetElementsWithContribution(dts))
V0 = xf.H1(mesh, order=3)
V = xf.Restrict(V0, active)
u, v = V.TnT()


def assemble(form):
    a = xf.RestrictedBilinearForm(V, element_restriction=active)
    a += form * xf.dCut(lsets, dts, definedonelements=active)
    return a.Assemble()


a, ta = timed("Assemble (grad u.grad v + u v), ml rules", lambda: assemble(xf.grad(u) * xf.grad(v) + u * v))
nact = active.NumSet()
print(f"active elements {nact}, ndof {V.ndof}, nnz {a.mat.nze}: {nact / ta / 1e3:.2f} M matrices/s")
m = assemble(u * v)
k = assemble(xf.grad(u) * xf.grad(v))
one = torch.zeros(V.ndof, dtype=torch.float64, device=dev)
rows = V.rows(active.indices())
one[rows[:, :4].reshape(-1).long()] = 1.0
vol = xf.Integrate({"levelset": lsets, "domain_type": dts}, mesh=mesh, cf=1.0, order=0)
tot = float((one * (m.mat * one)).sum())
k1 = float((k.mat * one).abs().max())
print(f"1^T M 1 = {tot:.15f}  |domain| (order-0 ml rule) = {vol:.15f}  rel diff {abs(tot - vol) / vol:.2e};  max|K 1| = {k1:.2e}")
assert abs(tot - vol) < 1e-11 * vol and k1 < 1e-9
del a, m, k, V, V0, mlci, mesh
torch.cuda.empty_cache()

# ------------------------------------------------------------------ config 5
mesh = xf.MakeStructured3DMesh(nx=n5, device=dev)
print(f"config 5: n={n5} ne={mesh.ne} space-time P1 x P2, time_order 4, space order 2", flush=True)
r = 0.3
phis = []
for tn in (0.0, 0.5, 1.0):
    cy = 0.5 + 0.1 * math.sin(2 * math.pi * tn)
    f = torch.sqrt((mesh.points[:, 0] - .5) ** 2 + (mesh.points[:, 1] - cy) ** 2 + (mesh.points[:, 2] - .5) ** 2) - r
    phis.append(torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f))
phi_st = torch.stack(phis).contiguous()
res = {}
for dt, name in ((NEG, "NEG"), (POS, "POS"), (IF, "IF")):
    (off, q, tv, w), ms = timed(f"space-time cut rule {name}", lambda: ops.cutquad_st(3, None, mesh.elements, mesh.points, phi_st, 2, dt, 4, 2))
    val = ops.integrate_monomial(3, None, mesh.elements, mesh.points, off, q, w, (0, 0, 0))
    res[name] = val
    print(f"   points {w.numel()}  ({w.numel() / ms / 1e6:.2f} G points/s, {mesh.ne / ms / 1e3:.1f} M tets/s)  integral {val:.12f}")
ball = 4.0 / 3.0 * math.pi * r ** 3
print(f"NEG + POS = {res['NEG'] + res['POS']:.15f} (exact 1);  NEG = {res['NEG']:.8f} vs ball volume {ball:.8f}")
assert abs(res["NEG"] + res["POS"] - 1.0) < 1e-11
assert abs(res["NEG"] - ball) < 2e-2 * ball
free, total = torch.cuda.mem_get_info()
print(f"mem used {(total - free) / 1e9:.1f} GB")
