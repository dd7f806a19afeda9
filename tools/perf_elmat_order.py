"""Element-matrix contraction at order p: FP64 tensor-core tiles (DMMA) vs the CUDA-core kernel
(development tool; numbers go to profiles/).  usage: perf_elmat_order.py [n] [p ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngsxfem_b200 as xf
from ngsxfem_b200 import ops
from ngsxfem_b200.pipeline import sphere_levelset

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
orders = [int(a) for a in sys.argv[2:]] or [2, 3, 4]
dev = torch.device("cuda", 0)
mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
ctx = ops.default_ctx()
bn, bp, bi, cr = ops.classify(3, mesh.elements, phi)
l_if = ops.bitset_to_list(bi, mesh.ne)
l_neg = ops.bitset_to_list(bn, mesh.ne)
TV = [(1, 1, 1.0, 0), (0, 0, 1.0, 0)]


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


print(f"n={n} ne={mesh.ne} cut={l_if.numel()} uncut NEG={l_neg.numel()}")
for p in orders:
    nd = ops.h1_ndof(3, p)
    off, q, w = ops.cutquad(3, l_if, mesh.elements, mesh.points, phi, 0, 2 * p)
    npts = w.numel()
    out = torch.empty((l_if.numel(), nd, nd), dtype=torch.float64, device=dev)
    sub = l_neg[: min(l_neg.numel(), 400000)].contiguous()
    out2 = torch.empty((sub.numel(), nd, nd), dtype=torch.float64, device=dev)
    nstd = ops.get_rule(3, 2 * p)[1].numel()
    for mode, name in ((-1, "cuda-core"), (1, "dmma")):
        ctx.set_option("elmat_dmma", mode)
        t = timed(lambda: ops.elmat_rule(3, p, l_if, mesh.elements, mesh.points, phi, off, q, w, TV, out=out))
        fl = 2.0 * npts * 4 * nd * nd  # full (unsymmetrised) contraction flops, as the reference computes them
        # force the point kernel for the uncut elements too (a non mass/stiffness-free term list would leave the fast path)
        t2 = timed(lambda: ops.elmat_rule(3, p, sub, mesh.elements, mesh.points, phi,
                                          torch.arange(0, (sub.numel() + 1) * nstd, nstd, device=dev, dtype=torch.int64),
                                          ops.get_rule(3, 2 * p)[0].to(dev).repeat(sub.numel(), 1),
                                          ops.get_rule(3, 2 * p)[1].to(dev).repeat(sub.numel()), TV, out=out2)) if p >= 2 else 0.0
        fl2 = 2.0 * sub.numel() * nstd * 4 * nd * nd
        print(f"p={p} nd={nd} {name:9s}: cut {l_if.numel()} el / {npts} pts {t:8.3f} ms = {fl / t / 1e9:8.2f} TFLOP/s(full-equiv) "
              f"{l_if.numel() / t / 1e3:8.2f} Mel/s | uncut-as-rule {sub.numel()} el x {nstd} pts {t2:8.3f} ms = {fl2 / t2 / 1e9:8.2f} TFLOP/s")
    ctx.set_option("elmat_dmma", 0)
