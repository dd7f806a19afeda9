"""timeline of the concurrent step: marks of the main chain and the interval of the cut-element branch on the side stream"""
import sys
import torch
sys.path.insert(0, ".")
import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import CutAssembly, PhaseTimer, sphere_levelset

dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=128, device=dev)
V0 = xf.H1(mesh, order=2)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)
for _ in range(3):
    prob.step(phi, concurrent=True)
torch.cuda.synchronize()
for rep in range(2):
    t = PhaseTimer(True)
    prob.step(phi, concurrent=True, timer=t)
    torch.cuda.synchronize()
    e0 = t.marks[0][1]
    print("main:", "  ".join(f"{n}@{e0.elapsed_time(e):.2f}" for n, e in t.marks))
    print("side:", "  ".join(f"{e0.elapsed_time(e):.2f}" for e in t.side))
