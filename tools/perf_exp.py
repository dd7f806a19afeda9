"""config-2 experiments: whole-step time (CUDA events, median of 8) for combinations of library options
usage: perf_exp.py [key=value ...]   e.g. csr_fork=0"""
import sys
import torch

sys.path.insert(0, ".")
import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import CutAssembly, PhaseTimer, sphere_levelset

n = 128
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
V0 = xf.H1(mesh, order=2)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)
opts = dict(a.split("=") for a in sys.argv[1:])
for c in (prob.ctx, prob.ctx_hp):
    for k, v in opts.items():
        c.set_option(k, int(v))


def run(concurrent, reps=8, phases=False):
    for _ in range(3):
        r = prob.step(phi, concurrent=concurrent)
    torch.cuda.synchronize()
    ts, acc = [], {}
    for _ in range(reps):
        t = PhaseTimer(phases)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = prob.step(phi, concurrent=concurrent, timer=t)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        if phases:
            for k, v in t.result().items():
                acc[k] = acc.get(k, 0.0) + v / reps
    ts.sort()
    return r, ts[len(ts) // 2], acc


r, t1, ph = run(False, phases=True)
print(f"{opts} serial {t1:.3f} ms phases", {k: round(v, 3) for k, v in ph.items()})
_, t2, _ = run(True)
print(f"{opts} concurrent {t2:.3f} ms = {r['units'] / t2 / 1e3:.1f} M matrices/s")
