"""Timeline (CUDA events) of the copies of the pipelined e2e steps of config 2: where do H2D and D2H overlap?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import HostBuffers, drain_host, sphere_levelset, step_host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
V0 = xf.H1(mesh, order=2)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
hb = HostBuffers(mesh.elements, mesh.points, V0.el_dofs, phi)
del mesh
for _ in range(2):
    step_host(hb, 3, V0.ndof, 2, dev, pipelined=True)
drain_host(hb, dev)
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True); base.record()
traces = []
import time
t0 = time.perf_counter()
for k in range(5):
    tr = [("host_enter_%.1f" % ((time.perf_counter() - t0) * 1e3), None)]
    step_host(hb, 3, V0.ndof, 2, dev, pipelined=True, trace=tr)
    tr.append(("host_leave_%.1f" % ((time.perf_counter() - t0) * 1e3), None))
    traces.append(tr)
drain_host(hb, dev)
torch.cuda.synchronize()
for k, tr in enumerate(traces):
    print(f"step {k}: " + "  ".join(name if ev is None else f"{name}={base.elapsed_time(ev):.1f}" for name, ev in tr))
