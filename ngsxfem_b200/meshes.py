"""Minimal simplicial mesh container + structured generators (input producers, below the hot path).

The reference takes its mesh topology from NGSolve's MeshAccess (third party): element->vertex,
element->edge/face, facet->element tables.  On the accelerated path those tables are plain INPUT
arrays of the C-ABI; this module only produces them for tests / benchmarks (torch tensors, CPU or
CUDA, int32 connectivity, float64 coordinates).  Local topology follows NGSolve's published
ElementTopology tables (tet edges {3,0},{3,1},{3,2},{0,1},{0,2},{1,2}; face i opposite vertex i;
trig edges {2,0},{1,2},{0,1}).
"""
import itertools

import torch

TET_EDGES = [(3, 0), (3, 1), (3, 2), (0, 1), (0, 2), (1, 2)]
TET_FACES = [(3, 1, 2), (3, 2, 0), (3, 0, 1), (0, 2, 1)]
TRIG_EDGES = [(2, 0), (1, 2), (0, 1)]


class Mesh:
    """points: (nv, dim) float64; elements: (ne, dim+1) int32."""

    def __init__(self, points, elements):
        self.points = points.contiguous().to(torch.float64)
        self.elements = elements.contiguous().to(torch.int32)
        self.dim = self.points.shape[1]
        assert self.elements.shape[1] == self.dim + 1
        self._topo = {}
        self.deformation = None

    def SetDeformation(self, deformation):
        """mesh.SetDeformation(gf) (python/lsetcurv.py:358-363): every cut integral / form assembled while it is set uses the
        isoparametric map (forms without their own `deformation=` argument pick it up)"""
        self.deformation = deformation

    def UnsetDeformation(self):
        self.deformation = None

    # ---- sizes (NGSolve names) ----
    @property
    def nv(self):
        return self.points.shape[0]

    @property
    def ne(self):
        return self.elements.shape[0]

    @property
    def device(self):
        return self.points.device

    def to(self, device):
        m = Mesh(self.points.to(device), self.elements.to(device))
        for k, v in self._topo.items():
            m._topo[k] = tuple(t.to(device) for t in v)
        return m

    # ---- topology ----
    def _entities(self, local):
        """unique sorted-vertex entities for the local table; returns (ents (n,k) int32, el_ents (ne,nl) int32)"""
        el = self.elements.to(torch.int64)
        k = len(local[0])
        nl = len(local)
        vs = []
        for ent in local:
            v = torch.stack([el[:, a] for a in ent], dim=1)
            v, _ = torch.sort(v, dim=1)
            vs.append(v)
        vs = torch.stack(vs, dim=1).reshape(-1, k)  # (ne*nl) x k, sorted vertices
        key = vs[:, 0] * self.nv + vs[:, 1]
        if k == 3:  # compress the pair key first: nv^3 would overflow int64 for nv > 2^21
            _, key = torch.unique(key, sorted=True, return_inverse=True)
            key = key * self.nv + vs[:, 2]
        uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
        ents = torch.empty((uniq.shape[0], k), dtype=torch.int32, device=self.device)
        ents[inv] = vs.to(torch.int32)
        return ents, inv.reshape(self.ne, nl).to(torch.int32)

    def edges(self):
        if "edges" not in self._topo:
            if self.dim == 1:
                raise ValueError("no edges in 1D")
            self._topo["edges"] = self._entities(TRIG_EDGES if self.dim == 2 else TET_EDGES)
        return self._topo["edges"]

    def faces(self):
        if "faces" not in self._topo:
            assert self.dim == 3
            self._topo["faces"] = self._entities(TET_FACES)
        return self._topo["faces"]

    def facets(self):
        """(facet vertices, el_facets, facet_els (nf,2) int32 with -1 for boundary)"""
        if "facets" not in self._topo:
            if self.dim == 3:
                ents, el_f = self.faces()
            elif self.dim == 2:
                ents, el_f = self.edges()
            else:
                ents = torch.arange(self.nv, dtype=torch.int32, device=self.device).reshape(-1, 1)
                el_f = self.elements[:, [1, 0]].contiguous()  # facet i opposite... vertices as facets
            nf = ents.shape[0]
            nl = el_f.shape[1]
            elid = torch.arange(self.ne, device=self.device, dtype=torch.int64).repeat_interleave(nl)
            fid = el_f.reshape(-1).to(torch.int64)
            order = torch.argsort(fid * self.ne + elid)
            fs, es = fid[order], elid[order]
            first = torch.ones_like(fs, dtype=torch.bool)
            first[1:] = fs[1:] != fs[:-1]
            fe = torch.full((nf, 2), -1, dtype=torch.int32, device=self.device)
            fe[fs[first], 0] = es[first].to(torch.int32)
            fe[fs[~first], 1] = es[~first].to(torch.int32)
            self._topo["facets"] = (ents, el_f, fe)
        return self._topo["facets"]

    # ---- periodic identification (MeshAccess::GetPeriodicFacet; consumed by GetFacetsWithNeighborTypes, cutinfo.cpp:509-518)
    def SetPeriodicFacets(self, partner):
        """partner: (nfacet,) int32, the identified facet of every facet, the facet itself where there is none
        (NGSolve's GetPeriodicFacet convention)"""
        partner = partner.to(self.device).to(torch.int32).contiguous()
        assert partner.shape == (self.nfacet,)
        self._topo["periodic"] = (partner,)
        self._topo.pop("facet_els_periodic", None)

    def periodic_facets(self):
        return self._topo["periodic"][0] if "periodic" in self._topo else None

    def IdentifyPeriodic(self, shifts, tol=1e-9):
        """identify boundary facets whose vertex sets differ by one of the translations `shifts` (set-up code, host side:
        sorted quantised vertex coordinates as keys)"""
        import numpy as np
        ents, _, fe = self.facets()
        P = self.points.cpu().numpy()
        E = ents.cpu().numpy().reshape(ents.shape[0], -1)
        bnd = np.flatnonzero(fe[:, 1].cpu().numpy() < 0)
        partner = np.arange(E.shape[0], dtype=np.int32)

        def key(X):   # X: (k, dim) vertex coordinates of a facet
            q = np.round(X / tol).astype(np.int64)
            return tuple(map(tuple, q[np.lexsort(q.T[::-1])]))

        table = {key(P[E[f]]): f for f in bnd}
        for sh in shifts:
            sh = np.asarray(sh, dtype=np.float64)
            for f in bnd:
                g = table.get(key(P[E[f]] + sh))
                if g is not None and g != f:
                    partner[f], partner[g] = g, f
        self.SetPeriodicFacets(torch.from_numpy(partner))
        return self

    def facet_els_periodic(self):
        """facet_els with the periodic neighbour filled in the way GetFacetsWithNeighborTypes sees it (cutinfo.cpp:509-518):
        a boundary facet whose identified facet has the LARGER number gets that facet's element as its second element; the
        facet with the larger number stays a boundary facet.  Without identification: facet_els itself."""
        _, _, fe = self.facets()
        partner = self.periodic_facets()
        if partner is None:
            return fe
        if "facet_els_periodic" not in self._topo:
            idx = torch.arange(fe.shape[0], device=fe.device, dtype=torch.int64)
            take = (fe[:, 1] < 0) & (partner.to(torch.int64) > idx)
            fe2 = fe.clone()
            fe2[take, 1] = fe[partner[take].long(), 0]
            self._topo["facet_els_periodic"] = (fe2.contiguous(),)
        return self._topo["facet_els_periodic"][0]

    def boundary_elements(self):
        """(bnd_vtx (nbe, dim) int32, facet ids): the boundary facets as surface elements (VorB = BND), ascending facet
        id.  NGSolve's own boundary-element numbering is third party and not needed: the table is an input."""
        if "bnd" not in self._topo:
            ents, _, fe = self.facets()
            ids = torch.nonzero(fe[:, 1] < 0).reshape(-1)
            self._topo["bnd"] = (ents[ids].contiguous(), ids.to(torch.int32))
        return self._topo["bnd"]

    @property
    def nedge(self):
        return self.edges()[0].shape[0]

    @property
    def nface(self):
        return self.faces()[0].shape[0] if self.dim == 3 else 0

    @property
    def nfacet(self):
        return self.facets()[0].shape[0]


def MakeStructured2DMesh(quads=False, nx=10, ny=10, mapping=None, device="cpu", periodic_x=False, periodic_y=False):
    """Unit square, nx x ny cells, each split into 2 triangles (lower-right, upper-left);
    element id = 2*(i + nx*j) + t.  (NGSolve's ngsolve.meshes.MakeStructured2DMesh is third party;
    its element ordering is not pinned and not needed: all tables are inputs.)"""
    if quads:
        raise NotImplementedError("quadrilateral cut rules (straightcutrule.cpp:206-501) are out of scope")
    xs = torch.linspace(0, 1, nx + 1, dtype=torch.float64, device=device)
    ys = torch.linspace(0, 1, ny + 1, dtype=torch.float64, device=device)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    pts = torch.stack([X.reshape(-1), Y.reshape(-1)], dim=1)
    j, i = torch.meshgrid(torch.arange(ny, device=device), torch.arange(nx, device=device), indexing="ij")
    v00 = (i + (nx + 1) * j).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    t0 = torch.stack([v00, v10, v11], dim=1)
    t1 = torch.stack([v00, v01, v11], dim=1)
    els = torch.stack([t0, t1], dim=1).reshape(-1, 3)
    m = Mesh(pts, els.to(torch.int32))
    if periodic_x or periodic_y:   # ngsolve.meshes.MakeStructured2DMesh(periodic_x=, periodic_y=): identified before the mapping
        m.IdentifyPeriodic([s for s, on in (((1.0, 0.0), periodic_x), ((0.0, 1.0), periodic_y)) if on])
    if mapping is not None:
        m.points = torch.stack(list(mapping(pts[:, 0], pts[:, 1])), dim=1).contiguous()
    return m


def MakeStructured3DMesh(hexes=False, nx=10, ny=None, nz=None, mapping=None, device="cpu", periodic_x=False, periodic_y=False,
                        periodic_z=False):
    """Unit cube, nx x ny x nz cells, each split into 6 Kuhn tets (the 6 monotone vertex paths
    (0,0,0)->(1,1,1)); vertex id = i + (nx+1)*(j + (ny+1)*k); element id = 6*cell + path index
    (SURVEY.md §8d C2)."""
    if hexes:
        raise NotImplementedError("hexahedral cut rules are out of scope")
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    xs = torch.linspace(0, 1, nx + 1, dtype=torch.float64, device=device)
    ys = torch.linspace(0, 1, ny + 1, dtype=torch.float64, device=device)
    zs = torch.linspace(0, 1, nz + 1, dtype=torch.float64, device=device)
    Z, Y, X = torch.meshgrid(zs, ys, xs, indexing="ij")
    pts = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], dim=1)
    k, j, i = torch.meshgrid(torch.arange(nz, device=device), torch.arange(ny, device=device),
                             torch.arange(nx, device=device), indexing="ij")
    base = (i + (nx + 1) * (j + (ny + 1) * k)).reshape(-1)
    stride = [1, nx + 1, (nx + 1) * (ny + 1)]
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [base]
        for a in perm:
            v.append(v[-1] + stride[a])
        tets.append(torch.stack(v, dim=1))
    els = torch.stack(tets, dim=1).reshape(-1, 4)
    m = Mesh(pts, els.to(torch.int32))
    if periodic_x or periodic_y or periodic_z:
        m.IdentifyPeriodic([s for s, on in (((1.0, 0, 0), periodic_x), ((0, 1.0, 0), periodic_y), ((0, 0, 1.0), periodic_z)) if on])
    if mapping is not None:
        m.points = torch.stack(list(mapping(pts[:, 0], pts[:, 1], pts[:, 2])), dim=1).contiguous()
    return m


def MakeStructured1DMesh(nx=10, device="cpu"):
    pts = torch.linspace(0, 1, nx + 1, dtype=torch.float64, device=device).reshape(-1, 1)
    i = torch.arange(nx, device=device)
    return Mesh(pts, torch.stack([i, i + 1], dim=1).to(torch.int32))
