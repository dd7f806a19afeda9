#!/usr/bin/env python
"""bench.py — cut-element matrices assembled per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 128] [--impl reference]

Workload (N=1): BASELINE.json configs[1] — unit cube, n^3 x 6 Kuhn tets (n=128: 12 582 912),
sphere level set r=0.4, H1 order 2, (grad u.grad v + u v) dCut(NEG) on HASNEG + u v dCut(IF) on IF,
RestrictedBilinearForm on Restrict(H1, HASNEG).  One STEP = one full pass of the hot path:
CutInfo classification -> Restrict (dof compression) -> cut quadrature -> element matrices ->
sparsity graph -> CSR scatter.  Unit = one element matrix assembled into the CSR (|HASNEG|).

  value : device-resident inputs (mesh tables, dof table, level set already in HBM)
  e2e   : the same pass from pinned HOST buffers, H2D of all inputs and D2H of the CutInfo
          BitArrays / cut ratios and the CSR matrix inside the timed region
  roofline : the dominant kernel of the step (largest share of device time), timed with CUDA
          events on the launching stream inside the timed steps; algorithmic bytes per DESIGN.md
  cpu_baseline : the oracle port (C restatement, OpenMP over the host cores) on a bounded sample
          (the same workload at n=64), rank 0 only.

`--impl reference` times that CPU port alone (ngsxfem itself cannot be built in this image: all
of its sources need NGSolve headers).  Timing: CUDA events, barrier + synchronize on
both sides, max over ranks; inputs/outputs of a step (> 5 GB) exceed the 126 MB L2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cut_element_matrices_per_sec"
UNIT = "matrices/s"
SAMPLE_N = 64


def workload_name(n):
    return f"unit cube {n}^3x6 Kuhn tets ({6 * n ** 3}), sphere r=0.4, H1 order 2, mass+stiffness dCut NEG on HASNEG + mass dCut IF"


# ------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample(n, steps, warmup):
    """time the oracle port on the n^3x6 sample; returns (matrices/s, cores, info)"""
    import numpy as np  # noqa: F401
    import ngsxfem_b200 as xf
    from ngsxfem_b200.pipeline import sphere_levelset
    from oracle.pipeline_ref import oracle_step
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    mesh = xf.MakeStructured3DMesh(nx=n, device="cpu")
    V0 = xf.H1(mesh, order=2)
    phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4).numpy()
    els, pts, eld = mesh.elements.numpy(), mesh.points.numpy(), V0.el_dofs.numpy()
    secs, units, phases = [], 0, {}
    for it in range(warmup + steps):
        r, sec = oracle_step(3, els, pts, eld, V0.ndof, 2, phi, threads=True)
        if it >= warmup:
            secs.append(sec)
            units = r["units"]
            phases = r["phases"]
    tot = sum(secs)
    return units * len(secs) / tot, cores, dict(units_per_step=units, ms_per_step=1e3 * tot / len(secs),
                                                 phases_s={k: round(v, 4) for k, v in phases.items()})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, cores, info = cpu_sample(SAMPLE_N, args.steps, max(1, args.warmup))
    sample = f"same workload at n={SAMPLE_N} ({6 * SAMPLE_N ** 3} tets, {info['units_per_step']} matrices per step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n), "sample": sample, "phases_s": info["phases_s"]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ GPU arm
# algorithmic bytes per unit of the candidate dominant kernels (DESIGN.md "Kernels")
def kernel_bytes(phase, r, nd):
    nn = nd * nd
    if phase == "csr":           # distributed path: plan + pattern + scatter timed as one phase
        return sum(kernel_bytes(k, r, nd) for k in ("csr_adjacency", "csr_pattern_pass", "csr_scatter"))
    if phase == "elmat":         # cut quadrature + uncut + cut element matrices
        return kernel_bytes("elmat_uncut", r, nd) + kernel_bytes("elmat_cut", r, nd)
    if phase == "csr_scatter":   # read every element matrix once + dof rows + the row's columns, write colidx + vals + rowptr
        return r["units"] * (nn * 8 + nd * 4) + r["nnz"] * (8 + 4 + 4) + r["ndof"] * 16
    if phase == "csr_pattern":   # compaction copy: read + write colidx, rowptr
        return r["nnz"] * 8 + r["ndof"] * 16
    if phase == "csr_pattern_pass":  # read dof rows (through the adjacency), write the sorted unique columns + counts
        return r["units"] * nd * (4 + 8) + r["nnz"] * 4 + r["ndof"] * 12
    if phase == "csr_adjacency":     # read the dof rows twice (count, fill), write + sort the transposed table
        return r["units"] * nd * (4 + 4 + 8 + 16) + r["ndof"] * 24
    if phase == "elmat_uncut":   # 16 B vtx ids + 96 B coords in, nd^2 doubles out
        return r["n_neg"] * (16 + 96 + 4 + nn * 8)
    if phase == "elmat_cut":
        return r["n_if"] * (16 + 96 + 32 + 4 + 2 * nn * 8) + sum(r["npts"]) * 32
    if phase == "classify":
        return None
    return None


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import ngsxfem_b200 as xf
    from ngsxfem_b200 import ops
    from ngsxfem_b200.pipeline import CutAssembly, HostBuffers, PhaseTimer, drain_host, sphere_levelset, step_host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the assembly path has no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = ops.default_ctx(local)

    # ---- synthetic inputs (not timed).  N>1: every rank owns one unit-cube slab of the chain
    # [0,1]^2 x [0,N] with its own sphere (weak scaling; element ranges are contiguous per rank)
    n = args.n
    nd = 10
    if world == 1:
        mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
        V0 = xf.H1(mesh, order=2)
        phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
        prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2, ctx=ctx)
        hb = HostBuffers(mesh.elements, mesh.points, V0.el_dofs, phi)
        step_dev = lambda t: prob.step(phi, timer=t)

        def step_e2e():
            # steps are software-pipelined: the H2D + kernels of step k+1 overlap the D2H of step k (alternating landing
            # buffers); every step still copies all its inputs in and all its results out inside the timed region
            return step_host(hb, 3, V0.ndof, 2, dev, ctx=ctx, pipelined=True)
        e2e_finish = lambda: drain_host(hb, dev)
    else:
        # element-partitioned chain of `world` unit cubes, one per rank, halo layer + NCCL interface merge
        import copy
        from ngsxfem_b200.distributed import DistributedCutAssembly, GpuBackend, make_slab_partition
        part, phi = make_slab_partition(n, world, rank, dev, fe_order=2)
        asm = DistributedCutAssembly(part, GpuBackend(ctx), rank, world)
        hb = HostBuffers(part.el_vtx, part.vtx_xyz, part.el_dofs, phi)
        step_dev = lambda t: asm.step(phi, timer=t)
        e2e_finish = None
        asm_e2e = DistributedCutAssembly(part, asm.be, rank, world)

        def step_e2e():
            p2 = copy.copy(part)
            p2.el_vtx = hb.el_vtx.to(dev, non_blocking=True)
            p2.vtx_xyz = hb.vtx_xyz.to(dev, non_blocking=True)
            p2.el_dofs = hb.el_dofs.to(dev, non_blocking=True)
            ph = hb.phi.to(dev, non_blocking=True)
            asm_e2e.part = p2
            r = asm_e2e.step(ph)
            d2h = sum(hb.land(f"bits{i}", b) for i, b in enumerate(r["cls"][:3])) + hb.land("cut_ratio", r["cls"][3])
            rows_g, rp, cols_g, vals = DistributedCutAssembly.owned_rows_global(r)
            d2h += hb.land("rows", rows_g) + hb.land("rowptr", rp) + hb.land("col", cols_g) + hb.land("vals", vals)
            torch.cuda.current_stream().synchronize()
            return r, hb.h2d_bytes(), d2h
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps, warmup, sampler=None, finish=None):
        for _ in range(warmup):
            fn(None)
        if finish:
            finish()
        barrier()
        timers = [PhaseTimer(True) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches
        e0.record()
        res = None
        for k in range(steps):
            res = fn(timers[k])
        if finish:
            finish()   # the launching stream waits for the outstanding copies: they are inside the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, res, timers, ctx.launches - l0

    steps, warmup = args.steps, max(3, args.warmup)
    with ClockSampler(local) as cs:
        ms, res, timers, launches = timed_loop(step_dev, steps, warmup)
    clocks = cs.summary()
    units = res["units"]
    # phase shares (CUDA events on the launching stream, inside the timed steps)
    phases = {}
    for t in timers:
        for k, v in t.result().items():
            phases[k] = phases.get(k, 0.0) + v / steps
    phases_max = None
    if world > 1:  # the slowest rank per phase (interior ranks of the chain have two neighbours)
        keys = sorted(phases)
        pv = torch.tensor([phases[k] for k in keys], device=dev, dtype=torch.float64)
        dist.all_reduce(pv, op=dist.ReduceOp.MAX)
        phases_max = {k: round(float(v), 4) for k, v in zip(keys, pv.tolist())}
    if "npts" not in res:
        res["npts"] = (0, 0)
        res["ndof"] = res["ndof_local"]
    dom = max((k for k in phases if kernel_bytes(k, res, nd)), key=lambda k: phases[k])
    del res["vals"], res["col"], res["rowptr"]
    torch.cuda.empty_cache()

    # ---- e2e: host buffers in, host CSR out, copies inside the timed region
    e2e_steps = max(1, min(steps, 10))
    info = {}

    def e2e_fn(_):
        r, h2d, d2h = step_e2e()
        info.update(h2d=h2d, d2h=d2h)
        return r
    ms_e2e, _, _, _ = timed_loop(e2e_fn, e2e_steps, 2, finish=e2e_finish)   # 2 warm-up steps: both landing sets get pinned

    tot_units = units * world
    if world > 1:
        tu = torch.tensor([units], device=dev, dtype=torch.int64)
        dist.all_reduce(tu)
        tot_units = int(tu.item())
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        kb = kernel_bytes(dom, res, nd)
        ach = kb / (phases[dom] * 1e-3) / 1e9
        # DRAM bytes of that phase's kernels from the committed `ncu --set full` capture of this workload
        # (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch), n=128 / 1 GPU only
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if world == 1 and n == tj.get("n"):
                traffic = tj["phases"].get(dom)
        except Exception:
            pass
        cpu_baseline = None
        if world == 1:  # the CPU arm is timed on rank 0 at N=1 only (other ranks would compete for the host cores)
            cpu_val, cores, cinfo = cpu_sample(SAMPLE_N, 2, 1)
            sample = f"same workload at n={SAMPLE_N} ({6 * SAMPLE_N ** 3} tets, {cinfo['units_per_step']} matrices per step), 2 timed passes"
            cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": tot_units * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n) + ("" if world == 1 else f"; x{world} ranks: chain of {world} unit cubes, "
                       "spheres centred on the slab interfaces (same per-rank element counts), halo layer + NCCL interface-row merge"),
                       "units_per_step_per_gpu": units, "n_neg": res["n_neg"], "n_if": res["n_if"],
                       "ndof": res["ndof"], "nnz": res["nnz"], "cut_points": list(res["npts"]),
                       "l2_policy": "inputs+outputs of a step (>5 GB) exceed the 126 MB L2; no flush needed",
                       "phases_ms": {k: round(v, 4) for k, v in phases.items()},
                       **({"phases_ms_max_over_ranks": phases_max} if phases_max else {})},
            "e2e": {"value": tot_units * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / e2e_steps,
                    "steps": e2e_steps, "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"],
                    "pipelining": ("steps software-pipelined: H2D + kernels of step k+1 overlap the D2H of step k (two sets of pinned "
                                   "landing buffers)") if world == 1 else "none"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kb,
                         "ms_per_launch": phases[dom]},
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
