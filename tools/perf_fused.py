"""config 2 (n = 128, sphere r = 0.4, P2): the one-row-pass path against the two-pass path — per-phase CUDA-event times
(non-concurrent) and whole-step times with the cut-element branch on a side stream."""
import sys
import torch

sys.path.insert(0, ".")
import ngsxfem_b200 as xf
from ngsxfem_b200.pipeline import CutAssembly, PhaseTimer, sphere_levelset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
V0 = xf.H1(mesh, order=2)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2)


def run(mode, concurrent, reps=6, phases=False):
    for _ in range(3):
        r = prob.step(phi, mode=mode, concurrent=concurrent)
    torch.cuda.synchronize()
    ts = []
    acc = {}
    for _ in range(reps):
        t = PhaseTimer(phases)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = prob.step(phi, mode=mode, concurrent=concurrent, timer=t)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        if phases:
            for k, v in t.result().items():
                acc[k] = acc.get(k, 0.0) + v / reps
    ts.sort()
    return r, ts[len(ts) // 2], acc


if len(sys.argv) > 2:
    prob.ctx.set_option("csr_trips", int(sys.argv[2]))
ra, ta, pa = run("fused", False, phases=True)
print(f"fused   serial     : {ta:.3f} ms  units={ra['units']} nnz={ra['nnz']} ndof={ra['ndof']}")
print("   phases:", {k: round(v, 3) for k, v in pa.items()})
rb, tb, pb = run("twopass", False, phases=True)
print(f"twopass serial     : {tb:.3f} ms")
print("   phases:", {k: round(v, 3) for k, v in pb.items()})
_, tc, _ = run("fused", True)
print(f"fused   concurrent : {tc:.3f} ms  = {ra['units'] / tc / 1e3:.1f} M matrices/s")
_, tc0, _ = run("fused", False)
print(f"fused   serial (no phase events) : {tc0:.3f} ms")
print("equal bits:", torch.equal(ra["rowptr"], rb["rowptr"]), torch.equal(ra["col"], rb["col"]), torch.equal(ra["vals"], rb["vals"]))
