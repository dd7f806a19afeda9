"""config-3 shape: where the time of RestrictedBilinearForm.Assemble goes — every ops.* call of one Assemble timed with CUDA
events (synchronising: the sum exceeds the asynchronous total).  usage: perf_config3_phases.py [n=40] [k=4]"""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
lset = xf.InterpolateToP1(lambda x, y, z: (torch.sqrt((x - .5) ** 2 + (y - .5) ** 2) - 0.28) ** 2 + (z - .5) ** 2 - 0.12 ** 2, mesh)
V0 = xf.H1(mesh, order=k, dgjumps=True)
Vd = xf.H1(mesh, order=k)
g = torch.Generator().manual_seed(7)
defo = xf.Deformation(Vd, ((torch.rand((Vd.ndof, 3), generator=g, dtype=torch.float64) - 0.5) * (0.004 / n)).to(dev))
ci = xf.CutInfo(mesh, lset)
hasneg, hasif = ci.GetElementsOfType(xf.HASNEG), ci.GetElementsOfType(xf.IF)
gpf = xf.GetFacetsWithNeighborTypes(mesh, a=hasneg, b=hasif, use_and=True)
V = xf.Restrict(V0, hasneg)
u, v = V.TnT()
h = xf.specialcf.mesh_size


def assemble():
    a = xf.RestrictedBilinearForm(V, element_restriction=hasneg, facet_restriction=gpf)
    a += (xf.grad(u) * xf.grad(v)) * xf.dCut(lset, xf.NEG, definedonelements=hasneg, deformation=defo)
    a += (10.0 * k * k / h) * u * v * xf.dCut(lset, xf.IF, definedonelements=hasif, deformation=defo)
    a += (0.1 / h ** 2) * (u - u.Other()) * (v - v.Other()) * xf.dFacetPatch(definedonelements=gpf, deformation=defo)
    return a.Assemble()


for _ in range(2):
    assemble()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); assemble(); e1.record(); torch.cuda.synchronize()
print(f"Assemble (asynchronous) {e0.elapsed_time(e1):.3f} ms")

acc = {}


def wrap(name, fn):
    def w(*a, **kw):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(*a, **kw)
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    return w


for name in dir(ops):
    fn = getattr(ops, name)
    if isinstance(fn, types.FunctionType) and not name.startswith("_") and name not in ("ptr", "default_ctx", "h1_ndof", "nwords", "terms_array"):
        setattr(ops, name, wrap(name, fn))
torch.cuda.synchronize()
t0 = time.perf_counter()
assemble()
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) * 1e3
print(f"Assemble (every ops call synchronised) {tot:.3f} ms; inside ops.* {sum(acc.values()):.3f} ms, Python / torch glue {tot - sum(acc.values()):.3f} ms")
for name, ms in sorted(acc.items(), key=lambda x: -x[1]):
    print(f"  {name:28s} {ms:8.3f} ms")
