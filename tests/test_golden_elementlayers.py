"""Golden BitArray of the reference's own test (tests/pytests/test_elementlayers.py:18-23): IF elements of
`x - 1/2` on the 7x7 structured triangle mesh plus two layers of vertex- resp. edge-neighbours.  The layer growth
(xfem.utils.AddNeighborhood, python/utils.py) is restated in numpy here; the classification comes from the oracle
(CPU) and from the CUDA kernel (GPU)."""
import json
import os

import numpy as np
import pytest

from ngsxfem_b200 import meshes
from oracle import xoracle as xo
from util import mesh_np, unpack_bits

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "elementlayers.json")))


def grow(marker, els, shared, layers):
    """elements sharing >= `shared` vertices with a marked element, `layers` times"""
    ne = len(els)
    sets = [set(e) for e in els]
    m = marker.copy()
    for _ in range(layers):
        new = m.copy()
        for i in range(ne):
            if m[i]:
                continue
            new[i] = any(m[j] and len(sets[i] & sets[j]) >= shared for j in range(ne))
        m = new
    return m


def check(if_bits, els):
    for key, shared in (("vertex", 1), ("edge", 2)):
        got = grow(if_bits, els, shared, 2)
        ref = np.array([c == "1" for c in GOLD[key]])
        assert np.array_equal(got, ref), key


def test_oracle_matches_golden_bits():
    m = meshes.MakeStructured2DMesh(nx=7, ny=7)
    pts, els = mesh_np(m)
    phi = pts[:, 0] - 0.5
    bn, bp, bi, cr, err, _ = xo.cutinfo_update(2, els, phi)
    assert err == 0
    check(unpack_bits(bi, m.ne), els)


@pytest.mark.gpu
def test_gpu_matches_golden_bits():
    import torch
    from ngsxfem_b200 import ops
    m = meshes.MakeStructured2DMesh(nx=7, ny=7)
    pts, els = mesh_np(m)
    md = m.to("cuda")
    phi = torch.from_numpy(pts[:, 0] - 0.5).to("cuda")
    bn, bp, bi, cr = ops.classify(2, md.elements, phi)
    check(unpack_bits(bi.cpu().numpy().view(np.uint8), m.ne), els)
