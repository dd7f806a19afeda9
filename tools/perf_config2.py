"""First-look per-phase timings of the config-2 pipeline (development tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = "cuda:0"
t0 = time.time()
mesh = xf.MakeStructured3DMesh(nx=N, device=dev)
lset = xf.InterpolateToP1(lambda x, y, z: torch.sqrt((x - .5) ** 2 + (y - .5) ** 2 + (z - .5) ** 2) - 0.4, mesh)
V0 = xf.H1(mesh, order=2)
torch.cuda.synchronize()
print(f"setup {time.time()-t0:.2f}s ne={mesh.ne} nv={mesh.nv} nedge={mesh.nedge} ndof_full={V0.ndof}", flush=True)
ctx = ops.default_ctx()
TV = [(1, 1, 1.0, 0), (0, 0, 1.0, 0)]
TI = [(0, 0, 1.0, 0)]


def timed(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        ev0.record(); r = fn(); ev1.record(); torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    print(f"{name:28s} {min(ts):9.3f} ms (median {sorted(ts)[len(ts)//2]:9.3f})", flush=True)
    return r

phi = lset.vec
bn, bp, bi, cr = timed("classify", lambda: ops.classify(3, mesh.elements, phi))
has = timed("bitset_or", lambda: ops.bitset_op(ops.BIT_OR, bn, bi, mesh.ne))
l_neg = timed("list NEG", lambda: ops.bitset_to_list(bn, mesh.ne))
l_if = timed("list IF", lambda: ops.bitset_to_list(bi, mesh.ne))
print("NEG", l_neg.numel(), "IF", l_if.numel())
V = timed("Restrict", lambda: xf.Restrict(V0, has))
t_neg = timed("rows NEG", lambda: V.rows(l_neg))
t_if = timed("rows IF", lambda: V.rows(l_if))
m_neg = timed("elmat affine NEG", lambda: ops.elmat_std(3, 2, l_neg, mesh.elements, mesh.points, phi, 4, TV))
gb = m_neg.numel() * 8 / 1e9
q = timed("cutquad NEG", lambda: ops.cutquad(3, l_if, mesh.elements, mesh.points, phi, 0, 4))
q2 = timed("cutquad IF", lambda: ops.cutquad(3, l_if, mesh.elements, mesh.points, phi, 2, 4))
print("points NEG", q[2].numel(), "IF", q2[2].numel())
m_if = timed("elmat cut NEG", lambda: ops.elmat_rule(3, 2, l_if, mesh.elements, mesh.points, phi, q[0], q[1], q[2], TV))
timed("elmat cut IF (acc)", lambda: ops.elmat_rule(3, 2, l_if, mesh.elements, mesh.points, phi, q2[0], q2[1], q2[2], TI, out=m_if, accumulate=True))
timed("elmat general NEG(200k)", lambda: ops.elmat_std(3, 2, l_neg[:200000].contiguous(), mesh.elements, mesh.points, phi, 4, TV + [(2, 2, 0.0, 0)]))
tables = [t_neg, t_if]
plan = timed("csr_plan", lambda: ops.csr_plan(V.ndof, tables), reps=3)
print("ndof", V.ndof, "nnz", plan.nnz, "maxcand", )
rowptr, col = timed("csr_pattern", lambda: ops.csr_pattern(plan, dev), reps=3)
vals = timed("csr_scatter", lambda: ops.csr_scatter(plan, [(t_neg, m_neg), (t_if, m_if)], rowptr, col), reps=3)
print(f"elmat bytes {gb:.2f} GB; launches {ctx.launches}")
free, total = torch.cuda.mem_get_info()
print(f"mem used {(total-free)/1e9:.1f} GB")
