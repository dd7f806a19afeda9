"""shared helpers for the tests (numpy views of the torch meshes, level-set interpolation)."""
import numpy as np
import torch

from ngsxfem_b200 import meshes


def interpolate_to_p1(fn, pts, eps=1e-14):
    """InterpolateToP1 (utils/p1interpol.cpp:27-113): nodal values, |v|<eps -> +eps (:105-107)."""
    v = np.asarray(fn(*[pts[:, i] for i in range(pts.shape[1])]), dtype=np.float64)
    v = np.where(np.abs(v) < eps, eps, v)
    return v


def mesh_np(m):
    return m.points.cpu().numpy(), m.elements.cpu().numpy()


def unpack_bits(b, n):
    return np.unpackbits(np.asarray(b, dtype=np.uint8), bitorder="little")[:n].astype(bool)


def det_of(T, xo, xhat):
    """|det F| of the oracle trafo at xhat"""
    import ctypes as C
    F = np.zeros(9); x = np.zeros(3)
    xh = np.zeros(3); xh[:len(xhat)] = xhat
    xo.lib().xo_trafo_eval(C.byref(T), xh.ctypes.data_as(C.POINTER(C.c_double)),
                           F.ctypes.data_as(C.POINTER(C.c_double)), x.ctypes.data_as(C.POINTER(C.c_double)))
    d = T.dim
    return abs(np.linalg.det(F[:d * d].reshape(d, d))), x[:d].copy()


def integrate_rule_fn(xo, pts, els, rule_fn, f=None):
    """sum over elements of sum_q w_q |det F| f(x_q); rule_fn(T, el) -> (pts, wts)"""
    total = 0.0
    for el in range(els.shape[0]):
        T = xo.make_trafo(pts[els[el]], els[el])
        p, w = rule_fn(T, el)
        for q in range(len(w)):
            d, x = det_of(T, xo, p[q])
            total += w[q] * d * (1.0 if f is None else f(*x))
    return total
