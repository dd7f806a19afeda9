"""Two passes of the config-2 hot path (one warm, one to profile) — the short command ncu wraps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import CutAssembly, sphere_levelset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "twopass"
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
V0 = xf.H1(mesh, order=2)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)
torch.cuda.synchronize()
for _ in range(reps):
    r = prob.step(phi, mode=mode)
torch.cuda.synchronize()
print("units", r["units"], "nnz", r["nnz"])
