"""config-2 step with / without the concurrent element-matrix branch (development tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import CutAssembly, sphere_levelset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
V0 = xf.H1(mesh, order=2)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)
for conc in (False, True, False, True):
    for _ in range(3):
        prob.step(phi, concurrent=conc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r = prob.step(phi, concurrent=conc)
    e1.record(); torch.cuda.synchronize()
    print(f"concurrent={conc}: {e0.elapsed_time(e1) / 10:.3f} ms/step")
