"""Config 3 of BASELINE.json at a single-GPU size: order-k isoparametric CutFEM Laplace with ghost-penalty facet patches
on a torus level set (unit cube, torus R = 0.28, r = 0.12 around the cube centre), through the public surface
(CutInfo, Restrict, dCut(deformation=...), dFacetPatch(deformation=...), RestrictedBilinearForm.Assemble) and phase by
phase through the C ABI.  The deformation is an INPUT here (synthetic smooth field at order k, small amplitude): the
producer (lsetcurving) is outside the hot path.  usage: perf_config3.py [n=40] [k=4] [reps=3]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
t0 = time.time()
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
lset = xf.InterpolateToP1(lambda x, y, z: (torch.sqrt((x - .5) ** 2 + (y - .5) ** 2) - 0.28) ** 2 + (z - .5) ** 2 - 0.12 ** 2, mesh)
V0 = xf.H1(mesh, order=k, dgjumps=True)
Vd = xf.H1(mesh, order=k)
g = torch.Generator().manual_seed(7)
defo = xf.Deformation(Vd, ((torch.rand((Vd.ndof, 3), generator=g, dtype=torch.float64) - 0.5) * (0.004 / n)).to(dev))
torch.cuda.synchronize()
print(f"setup {time.time() - t0:.1f}s  n={n} k={k} ne={mesh.ne} ndof_full={V0.ndof} nd={V0.nd}", flush=True)
ctx = ops.default_ctx()


def timed(name, fn, r=reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(r):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:34s} {ms:9.3f} ms", flush=True)
    return out, ms


ci, _ = timed("CutInfo", lambda: xf.CutInfo(mesh, lset))
hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
gpf = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif, use_and=True)
V, _ = timed("Restrict", lambda: xf.Restrict(V0, hasneg))
l_neg, l_if = ci.GetElementsOfType(xf.NEG).indices(), hasif.indices()
_, _, fe = mesh.facets()
l_f = gpf.indices()
l_f = l_f[fe[l_f.long(), 1] >= 0].contiguous()
print(f"NEG {l_neg.numel()} IF {l_if.numel()} ghost-penalty facets {l_f.numel()} ndof {V.ndof}", flush=True)
T = [(ops.OP_GRAD, ops.OP_GRAD, 1.0, 0)]
nstd = ops.get_rule(3, 2 * k - 2)[1].numel()
with ctx.deformation(defo):
    m_neg, t1 = timed("elmat uncut (curved, std rule)", lambda: ops.elmat_std(3, k, l_neg, mesh.elements, mesh.points, lset.vec, -1, T))
    (off, q, w), _ = timed("cutquad NEG", lambda: ops.cutquad(3, l_if, mesh.elements, mesh.points, lset.vec, ops.NEG, 2 * k - 2))
    m_if, t2 = timed("elmat cut (curved)", lambda: ops.elmat_rule(3, k, l_if, mesh.elements, mesh.points, lset.vec, off, q, w, T))
    (off2, q2, w2), _ = timed("cutquad IF (curved weights)", lambda: ops.cutquad(3, l_if, mesh.elements, mesh.points, lset.vec, ops.IF, 2 * k))
    _, t3 = timed("elmat IF nitsche-type (u v / h)", lambda: ops.elmat_rule(3, k, l_if, mesh.elements, mesh.points, lset.vec, off2, q2, w2,
                                                                             [(0, 0, 10.0 * k * k, -1)], out=m_if, accumulate=True))
    m_f, t4 = timed("facet patches (Newton map)", lambda: ops.facetpatch(3, k, l_f, fe, mesh.elements, mesh.points, -1, 0.1, -2))
nd = V0.nd
fl1 = 2.0 * l_neg.numel() * nstd * 4 * nd * nd
fl2 = 2.0 * w.numel() * 4 * nd * nd
npatch = 2 * ops.get_rule(3, 2 * k)[1].numel()
fl4 = 2.0 * l_f.numel() * npatch * (2 * nd) ** 2
print(f"contraction rates (full-matrix-equivalent flops): uncut {fl1 / t1 / 1e9:.2f} TFLOP/s ({nstd} pts/el), "
      f"cut {fl2 / t2 / 1e9:.2f} TFLOP/s ({w.numel() / max(1, l_if.numel()):.0f} pts/el), facet patch {fl4 / t4 / 1e9:.2f} TFLOP/s ({npatch} pts/facet)")
print(f"element matrices/s: uncut {l_neg.numel() / t1 / 1e3:.2f} M, cut {l_if.numel() / t2 / 1e3:.2f} M, facet {l_f.numel() / t4 / 1e3:.3f} M")

u, v = V.TnT()
h = xf.specialcf.mesh_size


def assemble():
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=gpf)
    a += (xf.grad(u) * xf.grad(v)) * xf.dCut(lset, xf.NEG, definedonelements=hasneg, deformation=defo)
    a += (10.0 * k * k / h) * u * v * xf.dCut(lset, xf.IF, definedonelements=hasif, deformation=defo)
    a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=gpf, deformation=defo)
    return a.Assemble()


a, ta = timed("RestrictedBilinearForm.Assemble", assemble)
units = l_neg.numel() + l_if.numel() + l_f.numel()
print(f"assembled: nnz {a.mat.nze}, {units} element/facet matrices in {ta:.2f} ms = {units / ta / 1e3:.2f} M matrices/s")
free, total = torch.cuda.mem_get_info()
print(f"mem used {(total - free) / 1e9:.1f} GB; launches {ctx.launches}")
