"""Dump golden vectors from a REAL ngsolve + ngsxfem installation (run this on a box that has `pip install xfem`; this image has
neither, so the repository ships no such file — tests/test_ngsolve_golden.py consumes whatever is placed under
tests/golden/ngsolve/*.json and skips otherwise).  This pins the third-party arithmetic DESIGN.md lists as unpinned: the point
order of SelectIntegrationRule, the H1HighOrderFE shape functions / dof numbering, CutInfo bits and the CSR matrix of a
RestrictedBilinearForm on small structured meshes, and scalar `Integrate` values over volume and boundary (`vb=BND`, including the
codim-2 `IF` case) cut domains.

    python tools/dump_ngsolve_golden.py tests/golden/ngsolve/ngsolve_golden.json

`--from-oracle` writes the SAME schema from this repository's own oracle / mesh generators (no ngsolve needed): that file only
proves that the consuming test reads the schema and compares entry-wise (tests/golden/ngsolve_schema_selftest.json)."""
import json
import sys

import numpy as np

REF_POINTS = {1: [[0.3, 0, 0], [0.77, 0, 0]], 2: [[0.2, 0.3, 0], [0.6, 0.15, 0]], 3: [[0.1, 0.2, 0.3], [0.5, 0.2, 0.15]]}
CASES = [dict(name="cube_n3_p2", dim=3, n=3, order=2, center=[0.5, 0.5, 0.5], radius=0.4),
         dict(name="cube_n2_p1", dim=3, n=2, order=1, center=[0.45, 0.55, 0.5], radius=0.33),
         dict(name="square_n5_p2", dim=2, n=5, order=2, center=[0.5, 0.48], radius=0.31)]


# meshes for the scalar integrals: the interface crosses the boundary of the domain
INTEGRAL_CASES = [dict(name="square_n4_crossing", dim=2, n=4, center=[0.1, 0.2], radius=0.62),
                  dict(name="square_n5_centered", dim=2, n=5, center=[0.5, 0.48], radius=0.31),
                  dict(name="cube_n2_crossing", dim=3, n=2, center=[0.1, 0.2, 0.15], radius=0.6)]
# scalar known answers of Integrate on those meshes: (domain type, VorB, rule order, monomial exponents)
INTEGRALS = [("neg", "vol", 2, [0, 0, 0]), ("pos", "vol", 3, [1, 1, 0]), ("if", "vol", 2, [0, 1, 0]),
             ("neg", "bnd", 2, [0, 0, 0]), ("pos", "bnd", 3, [1, 0, 0]), ("if", "bnd", 0, [0, 0, 0]), ("if", "bnd", 1, [1, 0, 0])]


def boundary_facets(els):
    """the facets that belong to exactly one element, as sorted vertex tuples (any order: the integrals are sums)"""
    import itertools
    cnt = {}
    for e in els:
        for f in itertools.combinations(sorted(int(v) for v in e), len(e) - 1):
            cnt[f] = cnt.get(f, 0) + 1
    return sorted(f for f, c in cnt.items() if c == 1)


def oracle_integral(xo, dim, pts, els, phi, dt, vorb, order, expo):
    """sum_q w_q |measure| x_q^expo from the oracle's cut rules.  vol: the elements; bnd: the boundary facets as
    (dim-1)-simplices, NEG / POS with the facet measure, IF on the segments of a 2D mesh as the point measure of
    cutint/straightcutrule.cpp:646-656 (weight / measure, so that measure * weight = the untransformed weight)."""
    DT = {"neg": 0, "pos": 1, "if": 2}[dt]
    if vorb == "vol":
        return float(xo.integrate_monomial(dim, els, pts, phi, DT, order, expo))
    if dim == 3 and dt == "if":
        return None   # the reference has no meaningful codim-2 rule on boundary triangles
    total = 0.0
    T = xo.ref_trafo(dim - 1)
    for f in boundary_facets(els):
        f = list(f)
        X = pts[f]
        rp, rw, _ = xo.straight_cut_rule(T, phi[f], DT, order)
        if len(rw) == 0:
            continue
        if dt == "if":
            meas = 1.0            # reference-segment interface rule: the untransformed weights
        elif dim == 2:
            meas = np.linalg.norm(X[0] - X[1])
        else:
            meas = np.linalg.norm(np.cross(X[0] - X[2], X[1] - X[2]))
        lam = np.concatenate([rp[:, :dim - 1], 1.0 - rp[:, :dim - 1].sum(axis=1, keepdims=True)], axis=1)
        xq = lam @ X
        total += float((rw * meas * np.prod(xq ** np.array(expo[:dim]), axis=1)).sum())
    return total


def from_ngsolve():
    from ngsolve import (BND, H1, SEGM, TET, TRIG, VOL, BilinearForm, ElementId, GridFunction, IntegrationRule, sqrt, x, y, z)  # noqa: F401
    from ngsolve.meshes import Make1DMesh, MakeStructured2DMesh, MakeStructured3DMesh
    from xfem import (HASNEG, IF, NEG, POS, CutInfo, InterpolateToP1, Restrict, RestrictedBilinearForm, dCut)
    out = {"source": "ngsolve", "rules": {}, "shapes": [], "cases": []}
    import ngsolve
    out["ngsolve_version"] = ngsolve.__version__
    for name, et, dim in (("segm", SEGM, 1), ("trig", TRIG, 2), ("tet", TET, 3)):
        for order in range(13):
            ir = IntegrationRule(et, order)
            out["rules"][f"{dim}_{order}"] = {"points": [list(p)[:dim] + [0.0] * (3 - dim) for p in ir.points], "weights": list(ir.weights)}
    meshes = {1: Make1DMesh(2), 2: MakeStructured2DMesh(quads=False, nx=1, ny=1), 3: MakeStructured3DMesh(hexes=False, nx=1, ny=1, nz=1)}
    for dim, mesh in meshes.items():
        for p in (1, 2, 3, 4):
            fes = H1(mesh, order=p)
            for k in range(min(mesh.ne, 6)):
                el = ElementId(VOL, k)
                fe = fes.GetFE(el)
                vnum = [v.nr for v in mesh[el].vertices]
                for xh in REF_POINTS[dim]:
                    shape = np.array(fe.CalcShape(*xh[:3])).reshape(-1)
                    dshape = np.array(fe.CalcDShape(*xh[:3])).reshape(len(shape), dim)
                    out["shapes"].append(dict(dim=dim, order=p, vnum=vnum, xhat=xh, shape=shape.tolist(), dshape=dshape.tolist()))
    for c in CASES:
        dim, n, p = c["dim"], c["n"], c["order"]
        mesh = MakeStructured3DMesh(hexes=False, nx=n, ny=n, nz=n) if dim == 3 else MakeStructured2DMesh(quads=False, nx=n, ny=n)
        cc, r = c["center"], c["radius"]
        lset = sqrt((x - cc[0]) ** 2 + (y - cc[1]) ** 2 + ((z - cc[2]) ** 2 if dim == 3 else 0)) - r
        lsetp1 = GridFunction(H1(mesh, order=1))
        InterpolateToP1(lset, lsetp1)
        ci = CutInfo(mesh, lsetp1)
        hasneg = ci.GetElementsOfType(HASNEG)
        Vb = H1(mesh, order=p)
        Vh = Restrict(Vb, hasneg)
        u, v = Vh.TnT()
        a = RestrictedBilinearForm(Vh, element_restriction=hasneg, check_unused=False)
        a += (ngsolve.grad(u) * ngsolve.grad(v) + u * v) * dCut(lsetp1, NEG, definedonelements=hasneg)
        a += u * v * dCut(lsetp1, IF, definedonelements=ci.GetElementsOfType(IF))
        a.Assemble()
        vals, cols, rowptr = a.mat.CSR()
        els = [[vv.nr for vv in mesh[ElementId(VOL, k)].vertices] for k in range(mesh.ne)]
        pts = [list(mesh[vv].point) for vv in mesh.vertices]

        out["cases"].append(dict(
            name=c["name"], dim=dim, order=p, points=pts, elements=els, lset_p1=list(lsetp1.vec),
            el_dofs=[list(Vb.GetDofNrs(ElementId(VOL, k))) for k in range(mesh.ne)], ndof_full=Vb.ndof,
            bits={"neg": [int(b) for b in ci.GetElementsOfType(NEG)], "pos": [int(b) for b in ci.GetElementsOfType(POS)],
                  "if": [int(b) for b in ci.GetElementsOfType(IF)]},
            restricted_el_dofs=[list(Vh.GetDofNrs(ElementId(VOL, k))) for k in range(mesh.ne)], ndof=Vh.ndof,
            csr={"rowptr": [int(i) for i in rowptr], "cols": [int(i) for i in cols], "vals": [float(t) for t in vals]}))
    from xfem import Integrate as XIntegrate
    coords = (x, y, z)
    out["integral_cases"] = []
    for c in INTEGRAL_CASES:
        dim, n = c["dim"], c["n"]
        mesh = MakeStructured3DMesh(hexes=False, nx=n, ny=n, nz=n) if dim == 3 else MakeStructured2DMesh(quads=False, nx=n, ny=n)
        cc, r = c["center"], c["radius"]
        lset = sqrt((x - cc[0]) ** 2 + (y - cc[1]) ** 2 + ((z - cc[2]) ** 2 if dim == 3 else 0)) - r
        lsetp1 = GridFunction(H1(mesh, order=1))
        InterpolateToP1(lset, lsetp1)
        integrals = []
        for dt_name, vorb, order, expo in INTEGRALS:
            if dim == 3 and vorb == "bnd" and dt_name == "if":
                continue
            cf = ngsolve.CoefficientFunction(1.0)
            for i in range(dim):
                for _ in range(expo[i]):
                    cf = cf * coords[i]
            dtv = {"neg": NEG, "pos": POS, "if": IF}[dt_name]
            dx = dCut(lsetp1, dtv, order=order, vb=BND) if vorb == "bnd" else dCut(lsetp1, dtv, order=order)
            integrals.append(dict(dt=dt_name, vorb=vorb, order=order, expo=expo, value=float(XIntegrate(cf * dx, mesh))))
        out["integral_cases"].append(dict(
            name=c["name"], dim=dim, points=[list(mesh[vv].point) for vv in mesh.vertices],
            elements=[[vv.nr for vv in mesh[ElementId(VOL, k)].vertices] for k in range(mesh.ne)],
            lset_p1=list(lsetp1.vec), integrals=integrals))
    return out


def from_oracle():
    """the same schema from this repository's own restatement (schema self-test only — pins nothing)"""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import ngsxfem_b200 as xf
    from oracle import xoracle as xo
    from oracle.pipeline_ref import oracle_step
    out = {"source": "oracle-selftest", "rules": {}, "shapes": [], "cases": []}
    for dim in (1, 2, 3):
        for order in range(13):
            p, w = xo.select_rule(dim, order)
            out["rules"][f"{dim}_{order}"] = {"points": p.tolist(), "weights": w.tolist()}
    for dim in (1, 2, 3):
        for p in (1, 2, 3, 4):
            for vnum in ([0, 1, 2, 3][:dim + 1], [5, 2, 9, 4][:dim + 1]):
                for xh in REF_POINTS[dim]:
                    sh, dsh = xo.h1_shapes(dim, p, vnum, xh)
                    out["shapes"].append(dict(dim=dim, order=p, vnum=vnum, xhat=xh, shape=sh.tolist(), dshape=dsh.tolist()))
    for c in CASES:
        dim, n, p = c["dim"], c["n"], c["order"]
        mesh = xf.MakeStructured3DMesh(nx=n) if dim == 3 else xf.MakeStructured2DMesh(nx=n, ny=n)
        V = xf.H1(mesh, order=p)
        d2 = sum((mesh.points[:, i] - c["center"][i]) ** 2 for i in range(dim))
        phi = torch.sqrt(d2) - c["radius"]
        phi = torch.where(phi.abs() < 1e-14, torch.full_like(phi, 1e-14), phi)
        o, _ = oracle_step(dim, mesh.elements.numpy(), mesh.points.numpy(), V.el_dofs.numpy(), V.ndof, p, phi.numpy())
        act = (np.unpackbits(o["bits"][0], bitorder="little")[:mesh.ne] | np.unpackbits(o["bits"][2], bitorder="little")[:mesh.ne]).astype(bool)
        eld = V.el_dofs.numpy()
        used = np.zeros(V.ndof, dtype=bool)
        used[eld[act].reshape(-1)] = True
        dofmap = np.full(V.ndof, -1, dtype=np.int64)
        dofmap[used] = np.arange(int(used.sum()))
        red = dofmap[eld]
        red[~act] = -1
        out["cases"].append(dict(
            name=c["name"], dim=dim, order=p, points=mesh.points.tolist(), elements=mesh.elements.tolist(), lset_p1=phi.tolist(),
            el_dofs=V.el_dofs.tolist(), ndof_full=V.ndof,
            bits={k: np.unpackbits(o["bits"][i], bitorder="little")[:mesh.ne].tolist() for i, k in enumerate(("neg", "pos", "if"))},
            restricted_el_dofs=red.tolist(), ndof=int(o["ndof"]),
            csr={"rowptr": o["rowptr"].tolist(), "cols": o["col"].tolist(), "vals": o["vals"].tolist()}))
    out["integral_cases"] = integral_cases_from_oracle(xo)
    return out


def integral_cases_from_oracle(xo):
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import meshgen
    out = []
    for c in INTEGRAL_CASES:
        dim, n = c["dim"], c["n"]
        if dim == 3:
            pts, els = meshgen.structured_tets(n)
        else:   # unit square, n x n cells, two triangles per cell (lower-right, upper-left)
            xs = np.linspace(0.0, 1.0, n + 1)
            pts = np.array([[xs[i], xs[j]] for j in range(n + 1) for i in range(n + 1)])
            els = []
            for j in range(n):
                for i in range(n):
                    v00 = i + (n + 1) * j
                    els += [[v00, v00 + 1, v00 + n + 2], [v00, v00 + n + 1, v00 + n + 2]]
            els = np.array(els, dtype=np.int32)
        phi = np.sqrt(((pts - np.array(c["center"])) ** 2).sum(axis=1)) - c["radius"]
        phi = np.where(np.abs(phi) < 1e-14, 1e-14, phi)
        out.append(dict(name=c["name"], dim=dim, points=pts.tolist(), elements=np.asarray(els).tolist(), lset_p1=phi.tolist(),
                        integrals=case_integrals(xo, dim, pts, np.asarray(els, dtype=np.int32), phi)))
    return out


def case_integrals(xo, dim, pts, els, phi):
    out = []
    for dt_name, vorb, order, expo in INTEGRALS:
        val = oracle_integral(xo, dim, pts, els, phi, dt_name, vorb, order, expo)
        if val is not None:
            out.append(dict(dt=dt_name, vorb=vorb, order=order, expo=expo, value=val))
    return out


def add_integrals_to_selftest(path):
    """append the `integral_cases` section to an existing --from-oracle file without touching its other sections"""
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import xoracle as xo
    with open(path) as f:
        data = json.load(f)
    assert data["source"] == "oracle-selftest"
    data["integral_cases"] = integral_cases_from_oracle(xo)
    with open(path, "w") as f:
        json.dump(data, f)
    return data


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--add-integrals" in sys.argv:
        data = add_integrals_to_selftest(args[0])
    else:
        data = from_oracle() if "--from-oracle" in sys.argv else from_ngsolve()
        with open(args[0], "w") as f:
            json.dump(data, f)
    print("wrote", args[0], "rules", len(data["rules"]), "shape records", len(data["shapes"]), "cases", len(data["cases"]))
