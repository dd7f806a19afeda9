"""Multi-process host logic of the element-partitioned assembly (world_size 2 and 3, gloo on CPU):
slab partition with halo layers, (owner, gid) numbering, interface-row exchange, global ids.
The merged row-distributed matrix must equal the single-process assembly of the whole chain."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _global_reference(n, world, fe_order, scaling="weak"):
    """whole chain (weak) / THE cube (strong) on one process through the oracle; returns dict key -> row (dict key -> value)"""
    import ngsxfem_b200 as xf
    from ngsxfem_b200.pipeline import sphere_levelset
    from oracle.pipeline_ref import oracle_step
    if scaling == "strong":
        world = 1   # the global problem of a strong split is the single cube
    mesh = xf.MakeStructured3DMesh(nx=n, ny=n, nz=n * world)
    nv1 = n + 1
    pts = mesh.points.clone()
    K = torch.arange(mesh.nv, dtype=torch.int64) // (nv1 * nv1)
    pts[:, 2] = K.to(torch.float64) / n
    gmesh = xf.Mesh(pts, mesh.elements)
    V = xf.H1(gmesh, order=fe_order)
    centers = [(.5, .5, .5)] if world == 1 else [(.5, .5, float(j)) for j in range(world + 1)]
    phi = sphere_levelset(pts, centers, 0.4).numpy()
    res, _ = oracle_step(3, gmesh.elements.numpy(), pts.numpy(), V.el_dofs.numpy(), V.ndof, fe_order, phi)
    # keys of the global dofs (same construction as make_slab_partition)
    from ngsxfem_b200.distributed import structured_dof_keys
    NV = nv1 * nv1 * (n * world + 1)
    gid, _ = structured_dof_keys(gmesh, torch.arange(gmesh.nv, dtype=torch.int64), nv1, NV, fe_order,
                                 torch.arange(gmesh.ne, dtype=torch.int64))
    gid = gid.numpy()
    # compressed id -> key
    act = np.zeros(V.ndof, dtype=bool)
    els_act = np.concatenate([np.flatnonzero(np.unpackbits(res["bits"][0], bitorder="little")[:gmesh.ne]),
                              np.flatnonzero(np.unpackbits(res["bits"][2], bitorder="little")[:gmesh.ne])])
    act[V.el_dofs.numpy()[els_act].reshape(-1)] = True
    keys = gid[act]
    return res, keys


def _worker(rank, world, n, fe_order, port, out, gpu=False, scaling="weak"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if gpu:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ngsxfem_b200.distributed import DistributedCutAssembly, make_slab_partition
        if gpu:
            from ngsxfem_b200.distributed import GpuBackend
            part, phi = make_slab_partition(n, world, rank, torch.device("cuda", rank), fe_order=fe_order, scaling=scaling)
            asm = DistributedCutAssembly(part, GpuBackend(), rank, world)
        else:
            from dist_backend import OracleBackend
            part, phi = make_slab_partition(n, world, rank, "cpu", fe_order=fe_order, scaling=scaling)
            asm = DistributedCutAssembly(part, OracleBackend(), rank, world)
        res = asm.step(phi)
        rows_g, rp, cols_g, vals = asm.owned_rows_global(res)
        a, b = part.range_of[rank]
        own_ids = a + torch.nonzero(res["dofmap"][a:b] >= 0).reshape(-1)   # owned active dofs, key order
        cpu = lambda t: t.cpu()
        rows_g, rp, cols_g, vals, own_ids = cpu(rows_g), cpu(rp), cpu(cols_g), cpu(vals), cpu(own_ids)
        torch.save(dict(rows=rows_g, rowptr=rp, cols=cols_g, vals=vals, keys=part.gid.cpu()[own_ids], units=res["units"],
                        ndof_global=res["ndof_global"], n_neg=res["n_neg"], n_if=res["n_if"]),
                   os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,fe_order", [(2, 4, 2), (3, 3, 2), (2, 5, 1), (2, 3, 3), (2, 2, 4)])
def test_slab_partition_merge_equals_global(tmp_path, world, n, fe_order):
    _run_and_check(tmp_path, world, n, fe_order, gpu=False)


@pytest.mark.parametrize("world,n,fe_order", [(2, 6, 2), (3, 7, 2), (4, 8, 1)])
def test_strong_split_of_one_cube_equals_global(tmp_path, world, n, fe_order):
    """scaling="strong": THE unit cube (BASELINE config 2 shape) cut into `world` z-slabs of work-balanced layer counts"""
    _run_and_check(tmp_path, world, n, fe_order, gpu=False, scaling="strong")


@pytest.mark.gpu
@pytest.mark.parametrize("world,n,fe_order,scaling", [(2, 6, 2, "weak"), (2, 5, 1, "weak"), (2, 8, 2, "strong")])
def test_nccl_slab_partition_merge_equals_global(tmp_path, world, n, fe_order, scaling):
    """the same check with the CUDA kernels (interface bookkeeping in merge.cu, static-capacity messages) and NCCL
    send/recv (needs >= 2 GPUs: gpurun --gpus 2)"""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    _run_and_check(tmp_path, world, n, fe_order, gpu=True, scaling=scaling)


def _run_and_check(tmp_path, world, n, fe_order, gpu, scaling="weak"):
    port = 29500 + (os.getpid() % 2000) + world + (7 if gpu else 0)
    mp.spawn(_worker, args=(world, n, fe_order, port, str(tmp_path), gpu, scaling), nprocs=world, join=True)
    ref, keys = _global_reference(n, world, fe_order, scaling)
    parts = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    ndofg = parts[0]["ndof_global"]
    assert ndofg == ref["ndof"] == len(keys)
    assert sum(p["units"] for p in parts) == ref["units"]
    if world > 1 and scaling == "weak":  # weak-scaling construction: every rank sees two half spheres = the same element counts
        assert len({(p["n_neg"], p["n_if"]) for p in parts}) == 1
    # distributed global id -> key
    key_of = np.zeros(ndofg, dtype=np.int64)
    seen = np.zeros(ndofg, dtype=bool)
    for p in parts:
        r = p["rows"].numpy()
        assert not seen[r].any()
        seen[r] = True
        key_of[r] = p["keys"].numpy()
    assert seen.all()
    # reference rows keyed by gid
    ref_id_of_key = {int(k): i for i, k in enumerate(keys)}
    rp, col, val = ref["rowptr"], ref["col"], ref["vals"]
    nnz = 0
    scale = np.abs(val).max()
    for p in parts:
        rows, prp, cols, vals = p["rows"].numpy(), p["rowptr"].numpy(), p["cols"].numpy(), p["vals"].numpy()
        assert (cols >= 0).all()
        for i, g in enumerate(rows):
            ri = ref_id_of_key[int(key_of[g])]
            c_ref = col[rp[ri]:rp[ri + 1]]
            v_ref = val[rp[ri]:rp[ri + 1]]
            c_d = np.array([ref_id_of_key[int(key_of[c])] for c in cols[prp[i]:prp[i + 1]]])
            v_d = vals[prp[i]:prp[i + 1]]
            o = np.argsort(c_d)
            assert np.array_equal(c_d[o], c_ref), f"pattern of row {g}"
            assert np.allclose(v_d[o], v_ref, rtol=1e-12, atol=1e-13 * scale)
            nnz += len(c_ref)
    assert nnz == ref["nnz"]


# ------------------------------------------------------------------ config 4: multi-level-set domain, element-partitioned
def _mlset_values(pts, world):
    """cube-chain ∩ 3 planes ∩ sphere around the interface of ranks 0 and 1 (so that interface rows are active)"""
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    fns = [x + 0.31 * y - 0.83, -0.27 * x + y + 0.11 * (z - 0.6) - 0.79, 0.13 * x - 0.21 * y + (z - 0.6) - 0.81,
           torch.sqrt((x - .4) ** 2 + (y - .45) ** 2 + (z - 1.0) ** 2) - 0.37]
    return torch.stack([torch.where(f.abs() < 1e-14, torch.full_like(f, 1e-14), f) for f in fns]).contiguous()


def _mlset_worker(rank, world, n, fe_order, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from ngsxfem_b200.distributed import DistributedCutAssembly, MlsetGpuBackend, make_slab_partition
        part, _ = make_slab_partition(n, world, rank, dev, fe_order=fe_order)
        be = MlsetGpuBackend([(0, 0, 0, 0)]).bind(part)
        asm = DistributedCutAssembly(part, be, rank, world)
        res = asm.step(_mlset_values(part.vtx_xyz, world))
        rows_g, rp, cols_g, vals = asm.owned_rows_global(res)
        a, b = part.range_of[rank]
        own_ids = a + torch.nonzero(res["dofmap"][a:b] >= 0).reshape(-1)
        torch.save(dict(rows=rows_g.cpu(), rowptr=rp.cpu(), cols=cols_g.cpu(), vals=vals.cpu(), keys=part.gid.cpu()[own_ids.cpu()],
                        units=res["units"], ndof_global=res["ndof_global"]), os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("n,fe_order", [(5, 2), (4, 3)])
def test_nccl_mlset_domain_partitioned_equals_single_gpu(tmp_path, n, fe_order):
    """BASELINE config 4 shape: the multi-level-set domain assembled on 2 GPUs (halo layer, NCCL interface rows) equals the
    single-GPU RestrictedBilinearForm.Assemble of the whole chain, row by row through the global dof keys"""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    import ngsxfem_b200 as xf
    port = 29500 + (os.getpid() % 2000) + 11 + fe_order
    mp.spawn(_mlset_worker, args=(world, n, fe_order, port, str(tmp_path)), nprocs=world, join=True)
    # single-GPU reference on the whole chain
    dev = torch.device("cuda", 0)
    mesh = xf.MakeStructured3DMesh(nx=n, ny=n, nz=n * world, device=dev)
    nv1 = n + 1
    pts = mesh.points.clone()
    K = torch.arange(mesh.nv, device=dev, dtype=torch.int64) // (nv1 * nv1)
    pts[:, 2] = K.to(torch.float64) / n
    gmesh = xf.Mesh(pts, mesh.elements)
    phis = _mlset_values(pts, world)
    lsets = [xf.GridFunction(gmesh, phis[i]) for i in range(4)]
    dts = (0, 0, 0, 0)
    active = xf.MultiLevelsetCutInfo(gmesh, lsets).GetElementsWithContribution(dts)
    V0 = xf.H1(gmesh, order=fe_order)
    V = xf.Restrict(V0, active)
    u, v = V.TnT()
    a = xf.RestrictedBilinearForm(V, element_restriction=active)
    a += (xf.grad(u) * xf.grad(v) + u * v) * xf.dCut(lsets, dts, definedonelements=active)
    a.Assemble()
    from ngsxfem_b200.distributed import structured_dof_keys
    NV = nv1 * nv1 * (n * world + 1)
    gid, _ = structured_dof_keys(gmesh, torch.arange(gmesh.nv, device=dev, dtype=torch.int64), nv1, NV, fe_order,
                                 torch.arange(gmesh.ne, device=dev, dtype=torch.int64))
    keys = gid[V.dofmap >= 0].cpu().numpy()            # compressed id -> key (ascending compression)
    rp, col, val = (t.cpu().numpy() for t in (a.mat.rowptr, a.mat.col, a.mat.vals))
    parts = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    assert parts[0]["ndof_global"] == V.ndof and sum(p["units"] for p in parts) == active.NumSet()
    assert all(p["units"] > 0 for p in parts)
    key_of = np.zeros(V.ndof, dtype=np.int64)
    for p in parts:
        key_of[p["rows"].numpy()] = p["keys"].numpy()
    ref_id = {int(k): i for i, k in enumerate(keys)}
    scale, nnz = np.abs(val).max(), 0
    for p in parts:
        rows, prp, cols, vals = p["rows"].numpy(), p["rowptr"].numpy(), p["cols"].numpy(), p["vals"].numpy()
        for i, g in enumerate(rows):
            ri = ref_id[int(key_of[g])]
            c_d = np.array([ref_id[int(key_of[c])] for c in cols[prp[i]:prp[i + 1]]])
            o = np.argsort(c_d)
            assert np.array_equal(c_d[o], col[rp[ri]:rp[ri + 1]])
            assert np.allclose(vals[prp[i]:prp[i + 1]][o], val[rp[ri]:rp[ri + 1]], rtol=1e-12, atol=1e-13 * scale)
            nnz += prp[i + 1] - prp[i]
    assert nnz == len(col)


# ------------------------------------------------------------------ config 5: one space-time slab, element-partitioned
def _st_worker(rank, world, n, p, kt, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from ngsxfem_b200.distributed import DistributedCutAssembly, SpaceTimeGpuBackend, make_slab_partition
        part, phi_st = make_slab_partition(n, world, rank, dev, fe_order=p, radius=0.3, time_fe_order=kt)
        be = SpaceTimeGpuBackend(p, kt, kt, 2 * kt).bind(part)
        asm = DistributedCutAssembly(part, be, rank, world)
        res = asm.step(phi_st)
        rows_g, rp, cols_g, vals = asm.owned_rows_global(res)
        a, b = part.range_of[rank]
        own_ids = a + torch.nonzero(res["dofmap"][a:b] >= 0).reshape(-1)
        torch.save(dict(rows=rows_g.cpu(), rowptr=rp.cpu(), cols=cols_g.cpu(), vals=vals.cpu(), keys=part.gid.cpu()[own_ids.cpu()],
                        units=res["units"], ndof_global=res["ndof_global"]), os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,kt", [(5, 1, 2), (4, 2, 1)])
def test_nccl_spacetime_slab_partitioned_equals_single_gpu(tmp_path, n, p, kt):
    """BASELINE config 5 shape: one space-time slab (moving spheres, P1/P2 x nodal time) assembled on 2 GPUs (halo layer,
    NCCL interface rows) equals the single-GPU BilinearForm(SpaceTimeFESpace).Assemble of the whole chain, row by row
    through the global space-time dof keys"""
    import math
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    import ngsxfem_b200 as xf
    from ngsxfem_b200.pipeline import sphere_levelset
    port = 29500 + (os.getpid() % 2000) + 23 + p
    mp.spawn(_st_worker, args=(world, n, p, kt, port, str(tmp_path)), nprocs=world, join=True)
    dev = torch.device("cuda", 0)
    mesh = xf.MakeStructured3DMesh(nx=n, ny=n, nz=n * world, device=dev)
    nv1 = n + 1
    pts = mesh.points.clone()
    K = torch.arange(mesh.nv, device=dev, dtype=torch.int64) // (nv1 * nv1)
    pts[:, 2] = K.to(torch.float64) / n
    gmesh = xf.Mesh(pts, mesh.elements)
    tfe = xf.ScalarTimeFE(kt)
    phis = []
    for t in tfe.nodes():
        cy = 0.5 + 0.1 * math.sin(2 * math.pi * 0.25 * t)
        phis.append(sphere_levelset(pts, [(.5, cy, float(j)) for j in range(world + 1)], 0.3))
    ls = xf.SpaceTimeGridFunction(gmesh, tfe, torch.stack(phis))
    hasneg = xf.CutInfo(gmesh, ls, time_order=0).GetElementsOfType(xf.HASNEG)
    Vs = xf.Restrict(xf.H1(gmesh, order=p), hasneg)
    V = xf.SpaceTimeFESpace(Vs, tfe)
    u, v = V.TnT()
    a = xf.BilinearForm(V)
    a += (u * v + xf.dt(u) * v + xf.grad(u) * xf.grad(v)) * xf.dCut(ls, xf.NEG, time_order=2 * kt, definedonelements=hasneg)
    a.Assemble()
    from ngsxfem_b200.distributed import structured_dof_keys
    NV = nv1 * nv1 * (n * world + 1)
    gid, _ = structured_dof_keys(gmesh, torch.arange(gmesh.nv, device=dev, dtype=torch.int64), nv1, NV, p,
                                 torch.arange(gmesh.ne, device=dev, dtype=torch.int64))
    nt = kt + 1
    ks = gid[Vs.dofmap >= 0].cpu().numpy()            # compressed space id -> key
    keys = np.concatenate([ks * nt + j for j in range(nt)])   # compressed space-time id (time-major) -> key
    rp, col, val = (t.cpu().numpy() for t in (a.mat.rowptr, a.mat.col, a.mat.vals))
    parts = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    assert parts[0]["ndof_global"] == V.ndof and sum(q["units"] for q in parts) == hasneg.NumSet()
    assert all(q["units"] > 0 for q in parts)
    key_of = np.zeros(V.ndof, dtype=np.int64)
    for q in parts:
        key_of[q["rows"].numpy()] = q["keys"].numpy()
    ref_id = {int(k): i for i, k in enumerate(keys)}
    scale, nnz = np.abs(val).max(), 0
    for q in parts:
        rows, prp, cols, vals = q["rows"].numpy(), q["rowptr"].numpy(), q["cols"].numpy(), q["vals"].numpy()
        for i, g in enumerate(rows):
            ri = ref_id[int(key_of[g])]
            c_d = np.array([ref_id[int(key_of[c])] for c in cols[prp[i]:prp[i + 1]]])
            o = np.argsort(c_d)
            assert np.array_equal(c_d[o], col[rp[ri]:rp[ri + 1]])
            assert np.allclose(vals[prp[i]:prp[i + 1]][o], val[rp[ri]:rp[ri + 1]], rtol=1e-12, atol=1e-13 * scale)
            nnz += prp[i + 1] - prp[i]
    assert nnz == len(col)


def _st_gloo_worker(rank, world, n, p, kt, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dist_backend import SpaceTimeOracleBackend
        from ngsxfem_b200.distributed import DistributedCutAssembly, make_slab_partition
        part, phi_st = make_slab_partition(n, world, rank, "cpu", fe_order=p, radius=0.3, time_fe_order=kt)
        asm = DistributedCutAssembly(part, SpaceTimeOracleBackend(p, kt, kt, 2 * kt).bind(part), rank, world)
        res = asm.step(phi_st)
        rows_g, rp, cols_g, vals = asm.owned_rows_global(res)
        a, b = part.range_of[rank]
        own_ids = a + torch.nonzero(res["dofmap"][a:b] >= 0).reshape(-1)
        torch.save(dict(rows=rows_g, rowptr=rp, cols=cols_g, vals=vals, keys=part.gid[own_ids], units=res["units"],
                        ndof_global=res["ndof_global"]), os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,p,kt", [(2, 3, 1, 2), (3, 2, 1, 1), (2, 2, 2, 1)])
def test_spacetime_slab_partition_merge_equals_global(tmp_path, world, n, p, kt):
    """config 5 host logic on CPU (gloo): the element-partitioned space-time slab (time-major dof table, keys = space key x
    (k_t+1) + time node) merged over `world` processes equals the single-process oracle assembly of the whole chain"""
    import math
    import ngsxfem_b200 as xf
    from dist_backend import SpaceTimeOracleBackend as B
    from ngsxfem_b200.distributed import structured_dof_keys
    from ngsxfem_b200.pipeline import sphere_levelset
    from oracle import xoracle as xo
    port = 29500 + (os.getpid() % 2000) + 31 + world + p
    mp.spawn(_st_gloo_worker, args=(world, n, p, kt, port, str(tmp_path)), nprocs=world, join=True)
    # single-process reference of the whole chain
    mesh = xf.MakeStructured3DMesh(nx=n, ny=n, nz=n * world)
    nv1 = n + 1
    pts = mesh.points.clone()
    K = torch.arange(mesh.nv, dtype=torch.int64) // (nv1 * nv1)
    pts[:, 2] = K.to(torch.float64) / n
    gmesh = xf.Mesh(pts, mesh.elements)
    nodes = xf.ScalarTimeFE(kt).nodes()
    L = np.stack([sphere_levelset(pts, [(.5, 0.5 + 0.1 * math.sin(2 * math.pi * 0.25 * t), float(j)) for j in range(world + 1)], 0.3).numpy()
                  for t in nodes])
    els, xyz = gmesh.elements.numpy(), pts.numpy()
    neg, pos, cut = B.slab_classes(els, xyz, L, kt)
    act_el = np.flatnonzero(neg | cut).astype(np.int32)
    assert cut.any()
    Vs = xf.H1(gmesh, order=p)
    eld = Vs.el_dofs.numpy()
    act = np.zeros(Vs.ndof, bool)
    act[eld[act_el].reshape(-1)] = True
    nc = int(act.sum())
    dofmap = np.full(Vs.ndof, -1, np.int32)
    dofmap[act] = np.arange(nc, dtype=np.int32)
    nt = kt + 1
    rows_s = dofmap[eld[act_el]]
    table = np.concatenate([rows_s + j * nc for j in range(nt)], axis=1).astype(np.int32)
    mats = B.slab_elmats(els, xyz, L, act_el, p, kt, kt, 2 * kt, B.TERMS)
    rp, col = xo.graph(nc * nt, table)
    val = xo.scatter(nc * nt, table, mats, rp, col)
    NV = nv1 * nv1 * (n * world + 1)
    gid, _ = structured_dof_keys(gmesh, torch.arange(gmesh.nv, dtype=torch.int64), nv1, NV, p, torch.arange(gmesh.ne, dtype=torch.int64))
    ks = gid.numpy()[act]
    keys = np.concatenate([ks * nt + j for j in range(nt)])
    parts = [torch.load(os.path.join(str(tmp_path), f"rank{r}.pt")) for r in range(world)]
    assert parts[0]["ndof_global"] == nc * nt and sum(q["units"] for q in parts) == len(act_el)
    key_of = np.zeros(nc * nt, dtype=np.int64)
    for q in parts:
        key_of[q["rows"].numpy()] = q["keys"].numpy()
    ref_id = {int(k): i for i, k in enumerate(keys)}
    scale, nnz = np.abs(val).max(), 0
    for q in parts:
        rows, prp, cols, vals = q["rows"].numpy(), q["rowptr"].numpy(), q["cols"].numpy(), q["vals"].numpy()
        for i, g in enumerate(rows):
            ri = ref_id[int(key_of[g])]
            c_d = np.array([ref_id[int(key_of[c])] for c in cols[prp[i]:prp[i + 1]]])
            o = np.argsort(c_d)
            assert np.array_equal(c_d[o], col[rp[ri]:rp[ri + 1]])
            assert np.allclose(vals[prp[i]:prp[i + 1]][o], val[rp[ri]:rp[ri + 1]], rtol=1e-12, atol=1e-13 * scale)
            nnz += prp[i + 1] - prp[i]
    assert nnz == len(col)
