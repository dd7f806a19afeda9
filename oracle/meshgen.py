"""ORACLE (test infrastructure): numpy producers of the synthetic inputs of BASELINE config 2 — structured Kuhn tet mesh,
H1 order-1/2 dof table, sphere level set — so that the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) does not
import the product package.  Same numbering as ngsxfem_b200.meshes / fes (tests/test_oracle_meshgen.py compares them):
vertex id = i + (n+1)(j + (n+1)k); element id = 6*cell + path index over the 6 monotone vertex paths; order-2 dofs =
vertices, then one dof per edge, edges numbered by ascending (min vertex, max vertex); local edges {3,0},{3,1},{3,2},
{0,1},{0,2},{1,2} (NGSolve ElementTopology)."""
import itertools

import numpy as np

TET_EDGES = [(3, 0), (3, 1), (3, 2), (0, 1), (0, 2), (1, 2)]


def structured_tets(n):
    """unit cube, n^3 cells x 6 Kuhn tets: (points (nv,3) f64, elements (ne,4) int32)"""
    # torch.linspace's evaluation order (first half from the start, second half from the end), so that the coordinates
    # are bit-identical to the product's mesh generator
    step = 1.0 / n
    ii = np.arange(n + 1, dtype=np.float64)
    xs = np.where(ii < (n + 1) // 2, ii * step, 1.0 - (n - ii) * step)
    Z, Y, X = np.meshgrid(xs, xs, xs, indexing="ij")
    pts = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = (i + (n + 1) * (j + (n + 1) * k)).reshape(-1).astype(np.int64)
    stride = [1, n + 1, (n + 1) * (n + 1)]
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [base]
        for a in perm:
            v.append(v[-1] + stride[a])
        tets.append(np.stack(v, axis=1))
    els = np.stack(tets, axis=1).reshape(-1, 4)
    return np.ascontiguousarray(pts), np.ascontiguousarray(els.astype(np.int32))


def h1_dofs(els, nv, order):
    """(el_dofs (ne, nd) int32, ndof) of H1(order), order 1 or 2"""
    if order == 1:
        return np.ascontiguousarray(els.astype(np.int32)), int(nv)
    assert order == 2
    el = els.astype(np.int64)
    keys = []
    for a, b in TET_EDGES:
        lo, hi = np.minimum(el[:, a], el[:, b]), np.maximum(el[:, a], el[:, b])
        keys.append(lo * nv + hi)
    keys = np.stack(keys, axis=1)
    uniq, inv = np.unique(keys.reshape(-1), return_inverse=True)
    el_edges = inv.reshape(keys.shape)
    dofs = np.concatenate([el, nv + el_edges], axis=1)
    return np.ascontiguousarray(dofs.astype(np.int32)), int(nv + len(uniq))


def sphere_levelset(pts, center, radius, eps=1e-14):
    """nodal P1 values of |x - c| - r with the InterpolateToP1 perturbation (utils/p1interpol.cpp:105-107)"""
    acc = 0.0  # the product's evaluation order: ((dx^2 + dy^2) + dz^2)
    for i in range(pts.shape[1]):
        acc = acc + (pts[:, i] - center[i]) ** 2
    d = np.sqrt(acc) - radius
    return np.ascontiguousarray(np.where(np.abs(d) < eps, eps, d))
