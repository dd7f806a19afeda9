"""Golden vectors of the space-time path, made by the CPU restatement (oracle/) on fixed inputs:
tests/golden/spacetime_elmat.json — rules (points, times, weights), element matrices and vectors of two cut tets / triangles.
They freeze the oracle (drift protection): tests/test_oracle_spacetime.py compares the oracle with them at 1e-14; the GPU
parity tests compare the CUDA path with the oracle on seeded inputs.     python tools/gen_golden_spacetime.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xoracle as xo

CASES = [
    # dim, vertices, level-set values per time node (k_l+1 rows), k_l, domain, space order p, time order k_t, rule orders (time, space)
    dict(name="tet_p1_kt2_neg", dim=3, vtx=[[0.1, 0.0, 0.05], [1.0, 0.1, 0.0], [0.0, 0.9, 0.1], [0.2, 0.1, 1.1]],
         lset=[[-0.3, 0.4, 0.2, 0.5], [-0.1, 0.3, -0.2, 0.6], [0.2, 0.1, -0.4, 0.7]], kl=2, dt=0, p=1, kt=2, ot=4, os=2),
    dict(name="tet_p2_kt1_if", dim=3, vtx=[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
         lset=[[-0.5, 0.5, 0.25, 0.75], [0.25, -0.5, 0.5, 0.3]], kl=1, dt=2, p=2, kt=1, ot=2, os=4),
    dict(name="trig_p2_kt2_pos", dim=2, vtx=[[0.0, 0.1], [1.2, 0.0], [0.3, 0.9]],
         lset=[[-0.2, 0.3, 0.5], [0.1, -0.3, 0.4], [0.4, -0.6, 0.2]], kl=2, dt=1, p=2, kt=2, ot=4, os=4),
]
TERMS = [(0, 0, 1.25, 0), (1, 1, 0.5, 1), (3, 0, -0.75, 0), (0, 3, 0.3, -1)]
LTERMS = [(0, 1.5, 0, (2, 1, 0, 1)), (3, -0.5, 1, (0, 1, 0, 2))]


def compute(c):
    V = np.array(c["vtx"], dtype=float)
    T = xo.make_trafo(V, np.arange(c["dim"] + 1, dtype=np.int32))
    q, tv, w = xo.spacetime_cut_rule(T, np.array(c["lset"], dtype=float), c["kl"], c["dt"], c["ot"], c["os"])
    m = xo.elmat_st_from_rule(T, c["p"], c["kt"], q, tv, w, TERMS)
    v = xo.elvec_st_from_rule(T, c["p"], c["kt"], q, tv, w, LTERMS)
    return q, tv, w, m, v


if __name__ == "__main__":
    out = {"terms": TERMS, "lterms": [[a, b, c, list(e)] for a, b, c, e in LTERMS], "cases": []}
    for c in CASES:
        q, tv, w, m, v = compute(c)
        assert len(w) > 0
        out["cases"].append(dict(c, pts=q.tolist(), tvals=tv.tolist(), wts=w.tolist(), elmat=m.tolist(), elvec=v.tolist()))
    path = os.path.join(ROOT, "tests", "golden", "spacetime_elmat.json")
    json.dump(out, open(path, "w"), indent=0)
    print(path, [len(c["wts"]) for c in out["cases"]])
