#!/usr/bin/env python3
"""Generate the hard-coded symmetric simplex rules to full double precision.

NGSolve (third party, not vendored) hard-codes symmetric rules for low orders
(trig: 1 / 3 / 6 / 12 points, tet: 1 / 4 / 14 points) and uses Duffy-transformed
Gauss-Jacobi rules above.  The low-order symmetric rules are classical (Strang-Fix /
Dunavant / Walkington); this script re-derives their parameters from the moment
equations with mpmath (50 digits) from textbook starting values and writes
  - ngsxfem_b200/csrc/rule_tables.inc   (product; included by rules.cpp)
  - oracle/xo_rule_tables.inc           (oracle; included by xo_rules.c)
  - tests/golden/rules_sym.json         (fixture for the tests)
The files are committed; the script is only needed to regenerate them.
"""
import json
import os
from mpmath import mp, mpf, findroot, sqrt, factorial

mp.dps = 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def trig_moment(i, j):
    # int over reference triangle x^i y^j = i! j! / (i+j+2)!
    return factorial(i) * factorial(j) / factorial(i + j + 2)


def tet_moment(i, j, k):
    return factorial(i) * factorial(j) * factorial(k) / factorial(i + j + k + 3)


def trig_orbit3(a):      # (a,a,1-2a) permutations -> (x,y)
    b = 1 - 2 * a
    return [(a, a), (a, b), (b, a)]


def trig_orbit6(a, b):
    c = 1 - a - b
    return [(a, b), (b, a), (a, c), (c, a), (b, c), (c, b)]


def trig_rule12():
    # Dunavant degree 6: two 3-orbits + one 6-orbit
    def pts(a1, a2, a3, b3):
        return trig_orbit3(a1), trig_orbit3(a2), trig_orbit6(a3, b3)

    def eqs(a1, w1, a2, w2, a3, b3, w3):
        o1, o2, o3 = pts(a1, a2, a3, b3)
        res = []
        for (i, j) in [(0, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (3, 3)]:
            s = sum(w1 * x**i * y**j for x, y in o1) + sum(w2 * x**i * y**j for x, y in o2) \
                + sum(w3 * x**i * y**j for x, y in o3)
            res.append(s - trig_moment(i, j))
        return res

    x0 = [mpf('0.063089014491502'), mpf('0.0254224531851035'), mpf('0.249286745170910'),
          mpf('0.0583931378631895'), mpf('0.053145049844817'), mpf('0.310352451033784'),
          mpf('0.041425537809187')]
    sol = findroot(eqs, x0, tol=1e-45, maxsteps=100)
    a1, w1, a2, w2, a3, b3, w3 = [sol[i] for i in range(7)]
    o1, o2, o3 = pts(a1, a2, a3, b3)
    P = o1 + o2 + o3
    W = [w1] * 3 + [w2] * 3 + [w3] * 6
    return P, W


def trig_rule6():
    def eqs(a1, w1, a2, w2):
        o1, o2 = trig_orbit3(a1), trig_orbit3(a2)
        res = []
        for (i, j) in [(0, 0), (2, 0), (3, 0), (4, 0)]:
            s = sum(w1 * x**i * y**j for x, y in o1) + sum(w2 * x**i * y**j for x, y in o2)
            res.append(s - trig_moment(i, j))
        return res
    x0 = [mpf('0.091576213509771'), mpf('0.054975871827661'), mpf('0.445948490915965'),
          mpf('0.1116907948390055')]
    sol = findroot(eqs, x0, tol=1e-45, maxsteps=100)
    a1, w1, a2, w2 = [sol[i] for i in range(4)]
    return trig_orbit3(a1) + trig_orbit3(a2), [w1] * 3 + [w2] * 3


def tet_orbit4(a):
    b = 1 - 3 * a
    return [(a, a, a), (a, a, b), (a, b, a), (b, a, a)]


def tet_orbit6(c):
    d = mpf(1) / 2 - c
    return [(c, c, d), (c, d, c), (d, c, c), (c, d, d), (d, c, d), (d, d, c)]


def tet_rule14():
    def eqs(a, wa, b, wb, c, wc):
        P = [(p, wa) for p in tet_orbit4(a)] + [(p, wb) for p in tet_orbit4(b)] + \
            [(p, wc) for p in tet_orbit6(c)]
        res = []
        for (i, j, k) in [(0, 0, 0), (2, 0, 0), (3, 0, 0), (4, 0, 0), (5, 0, 0), (2, 2, 0)]:
            s = sum(w * p[0]**i * p[1]**j * p[2]**k for p, w in P)
            res.append(s - tet_moment(i, j, k))
        return res
    x0 = [mpf('0.092735250310891'), mpf('0.01224884051939366'), mpf('0.310885919263300'),
          mpf('0.01878132095300264'), mpf('0.045503704125650'), mpf('0.00709100346284691')]
    sol = findroot(eqs, x0, tol=1e-45, maxsteps=100)
    a, wa, b, wb, c, wc = [sol[i] for i in range(6)]
    return tet_orbit4(a) + tet_orbit4(b) + tet_orbit6(c), [wa] * 4 + [wb] * 4 + [wc] * 6


def check(P, W, dim, deg):
    import itertools
    worst = mpf(0)
    for e in itertools.product(range(deg + 1), repeat=dim):
        if sum(e) > deg:
            continue
        s = mpf(0)
        for p, w in zip(P, W):
            t = w
            for d in range(dim):
                t *= p[d] ** e[d]
            s += t
        ex = trig_moment(*e) if dim == 2 else tet_moment(*e)
        worst = max(worst, abs(s - ex))
    return worst


def fmt(v):
    return mp.nstr(v, 25, strip_zeros=False)


def main():
    rules = {}
    third = mpf(1) / 3
    rules['trig1'] = ([(third, third)], [mpf(1) / 2], 1)
    h = mpf(1) / 2
    rules['trig3'] = ([(0, h), (h, 0), (h, h)], [mpf(1) / 6] * 3, 2)
    P, W = trig_rule6(); rules['trig6'] = (P, W, 4)
    P, W = trig_rule12(); rules['trig12'] = (P, W, 6)
    q = mpf(1) / 4
    rules['tet1'] = ([(q, q, q)], [mpf(1) / 6], 1)
    a = (5 - sqrt(5)) / 20
    b = (5 + 3 * sqrt(5)) / 20
    rules['tet4'] = ([(b, a, a), (a, b, a), (a, a, b), (a, a, a)], [mpf(1) / 24] * 4, 2)
    P, W = tet_rule14(); rules['tet14'] = (P, W, 5)

    lines = ["/* generated by tools/gen_rule_tables.py (mpmath, 50 digits) -- do not edit */"]
    js = {}
    for name, (P, W, deg) in rules.items():
        dim = len(P[0])
        err = check(P, W, dim, deg)
        assert err < mpf(10) ** -40, (name, err)
        assert all(w > 0 for w in W)
        n = len(W)
        lines.append(f"static const int SYM_{name.upper()}_N = {n};")
        lines.append(f"static const double SYM_{name.upper()}_P[{n}][3] = {{")
        for p in P:
            c = [fmt(mpf(x)) for x in p] + ["0.0"] * (3 - dim)
            lines.append("  {" + ", ".join(c) + "},")
        lines.append("};")
        lines.append(f"static const double SYM_{name.upper()}_W[{n}] = {{")
        for w in W:
            lines.append("  " + fmt(w) + ",")
        lines.append("};")
        js[name] = {"degree": deg, "points": [[float(x) for x in p] for p in P],
                    "weights": [float(w) for w in W],
                    "points_str": [[fmt(mpf(x)) for x in p] for p in P],
                    "weights_str": [fmt(w) for w in W]}
    txt = "\n".join(lines) + "\n"
    for rel in ("ngsxfem_b200/csrc/rule_tables.inc", "oracle/xo_rule_tables.inc"):
        path = os.path.join(ROOT, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(txt)
    os.makedirs(os.path.join(ROOT, "tests/golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests/golden/rules_sym.json"), "w") as f:
        json.dump(js, f, indent=1)
    print("ok", {k: len(v[1]) for k, v in rules.items()})


if __name__ == "__main__":
    main()
