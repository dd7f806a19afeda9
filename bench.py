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
  roofline : the dominant phase of the step (largest share of device time; CUDA events on the
          launching stream around its launches, in instrumented steps right after the timed loop);
          algorithmic bytes = SURVEY.md §8(d), peak = MEASURED_PEAKS.json
  cpu_baseline : the oracle port (C restatement, OpenMP over the host cores), ONE pass of the same
          workload (n=128: ~10 s), rank 0 at N=1 only.

`--impl reference` times that CPU port alone on the same config (ngsxfem itself cannot be built
in this image: all of its sources need NGSolve headers); it imports nothing of the
product package (inputs from oracle/meshgen.py).  Timing: CUDA events, barrier + synchronize on
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


def workload_name(n):
    return f"unit cube {n}^3x6 Kuhn tets ({6 * n ** 3}), sphere r=0.4, H1 order 2, mass+stiffness dCut NEG on HASNEG + mass dCut IF"


# ------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample(n, steps, warmup, budget_s=150.0):
    """time the oracle port on the n^3x6 workload with all host cores; returns (matrices/s, threads, info).
    Nothing of the product package is imported: the inputs come from oracle/meshgen.py (numpy)."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)   # torchrun exports OMP_NUM_THREADS=1: this arm uses every core
    from oracle import meshgen, xoracle as xo
    from oracle.pipeline_ref import oracle_step
    xo.set_num_threads(cores)
    threads = xo.num_threads()
    pts, els = meshgen.structured_tets(n)
    eld, ndof_full = meshgen.h1_dofs(els, len(pts), 2)
    phi = meshgen.sphere_levelset(pts, (.5, .5, .5), 0.4)
    secs, units, phases = [], 0, {}
    t_begin = time.perf_counter()
    for it in range(warmup + steps):
        r, sec = oracle_step(3, els, pts, eld, ndof_full, 2, phi, threads=True)
        units = r["units"]
        del r["vals"], r["col"], r["rowptr"]
        if it >= warmup:
            secs.append(sec)
            phases = r["phases"]
            if time.perf_counter() - t_begin > budget_s:   # a slow host: stop after the steps that fit the budget
                break
    tot = sum(secs)
    return units * len(secs) / tot, threads, dict(units_per_step=units, ms_per_step=1e3 * tot / len(secs), steps=len(secs),
                                                   phases_s={k: round(v, 4) for k, v in phases.items()})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warmup = max(1, min(args.warmup, 1))
    val, threads, info = cpu_sample(args.n, args.steps, warmup)
    sample = (f"{info['steps']} timed pass(es) of the full workload (n={args.n}: {6 * args.n ** 3} tets, "
              f"{info['units_per_step']} matrices per pass)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": info["steps"], "warmup": warmup, "ms_per_step": info["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n), "sample": sample, "phases_s": info["phases_s"],
                   "note": "CPU restatement of the hot path (oracle/, plain C + OpenMP), not ngsxfem itself"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
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
            time.sleep(0.02)

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
# algorithmic bytes per unit of every phase (SURVEY.md §8(d) / BASELINE.md §4).  The uncut elements'
# matrices are procedural now (never in HBM); the ALGORITHMIC figure of the assembly stays nd^2*8 + 2*nd*4 per unit
# + 12 B per stored nonzero, as §8(d) defines the work.
def kernel_bytes(phase, r, nd, ne=None, ndof_full=None):
    nn = nd * nd
    units, nnz, n_if, n_neg = r["units"], r["nnz"], r["n_if"], r["n_neg"]
    npts = sum(r.get("npts", (0, 0)))
    if phase in ("csr_rows_fused", "csr_scatter"):   # assemble: read every element matrix + its dof ids, write the CSR once
        return units * (nn * 8 + 2 * nd * 4) + nnz * 12
    if phase == "csr_count_pass":   # pattern: dof rows through the transposed table, columns + positions out
        return units * nd * (4 + 8 + 1) + nnz * 4 + r["ndof"] * 12
    if phase == "restrict_plan":    # Restrict + dof rows + transposed table
        return units * nd * (4 + 4 + 8) + (ndof_full or 0) * 8 + r["ndof"] * 12
    if phase == "classify":
        return int((ne or 0) * 56.4)
    if phase == "cutquad":
        return n_if * 2 * 48 + npts * 32
    if phase == "elmat_cut":     # connectivity, level set, coordinates in, one matrix out (+ the stored rules when they are read)
        return n_if * (16 + 96 + 32 + 40 + nn * 8) + (npts * 32 if r.get("rules_stored") else 0)
    if phase == "affine_coefs":
        return n_neg * (16 + 96 + 64)
    if phase == "csr":           # distributed path (two-pass backend): plan + pattern + scatter timed as one phase
        return units * (nn * 8 + 2 * nd * 4) + nnz * 12 + units * nd * (4 + 8 + 1) + nnz * 4
    return None


def fp64_peak():
    """measured FP64 peaks (tools/peak_fp64.py -> profiles/fp64_peak.json); nominal vendor figure otherwise"""
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
        return float(pk["dfma_tflops"]), "profiles/fp64_peak.json dfma_tflops (of measured: DFMA chain on all SMs)"
    except Exception:
        return 37.0, "37 TFLOP/s (of fallback: vendor nominal FP64)"


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

    # ---- synthetic inputs (not timed)
    n = args.n
    nd = 10
    check = None
    conc = False
    if world == 1:
        mesh = xf.MakeStructured3DMesh(nx=n, device=dev)
        V0 = xf.H1(mesh, order=2)
        phi = sphere_levelset(mesh.points, [(.5, .5, .5)], 0.4)
        prob = CutAssembly(3, mesh.elements, mesh.points, V0.el_dofs, V0.ndof, 2, ctx=ctx)
        hb = HostBuffers(mesh.elements, mesh.points, V0.el_dofs, phi)
        ne_loc, ndof_full = mesh.ne, V0.ndof
        # timed: the production configuration (cut-element branch on a side stream); instrumented: one stream, per phase
        # production step: ONE stream.  The cut-element branch on a low-priority side stream (XGPU_CONCURRENT=1) gave 5.44-7.76 ms
        # run to run against a steady 5.52 ms (profiles/r02_notes.md): its kernel fills the register file, so nothing co-resides
        conc = os.environ.get("XGPU_CONCURRENT", "0") == "1"
        step_dev = lambda t: prob.step(phi, timer=t, concurrent=(t is None) and conc)

        def step_e2e():
            # steps are software-pipelined: the H2D + kernels of step k+1 overlap the D2H of step k (alternating landing
            # buffers); every step still copies all its inputs in and all its results out inside the timed region
            return step_host(hb, 3, V0.ndof, 2, dev, ctx=ctx, pipelined=True)
        e2e_finish = lambda: drain_host(hb, dev)
    else:
        # element-partitioned (contiguous element ranges = z-slabs), halo layer + NCCL interface-row merge.
        # weak: chain of `world` unit cubes, one per rank; strong: THE config-2 cube cut into `world` slabs
        import copy
        from ngsxfem_b200.distributed import DistributedCutAssembly, GpuBackend, make_slab_partition, merge_check
        part, phi = make_slab_partition(n, world, rank, dev, fe_order=2, scaling=args.scaling)
        asm = DistributedCutAssembly(part, GpuBackend(ctx), rank, world)
        hb = HostBuffers(part.el_vtx, part.vtx_xyz, part.el_dofs, phi)
        ne_loc, ndof_full = part.el_vtx.shape[0], part.ndof_loc
        step_dev = lambda t: asm.step(phi, timer=t)
        asm_e2e = DistributedCutAssembly(part, asm.be, rank, world)

        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        hb.done = [None, None]

        def step_e2e():
            # software-pipelined like the single-GPU leg: the inputs of step k+1 arrive (copy stream s_in) and its kernels
            # run while the results of step k leave (copy stream s_out, two alternating sets of pinned landing buffers)
            hb.parity = 1 - getattr(hb, "parity", 1)
            par = f"#{hb.parity}"
            if hb.done[hb.parity] is not None:
                hb.done[hb.parity].synchronize()   # this landing set was filled two steps ago
            main = torch.cuda.current_stream(dev)
            p2 = copy.copy(part)
            with torch.cuda.stream(s_in):
                p2.el_vtx = hb.el_vtx.to(dev, non_blocking=True)
                ph = hb.phi.to(dev, non_blocking=True)
                p2.el_dofs = hb.el_dofs.to(dev, non_blocking=True)
                p2.vtx_xyz = hb.vtx_xyz.to(dev, non_blocking=True)
            main.wait_stream(s_in)
            for t_ in (p2.el_vtx, ph, p2.el_dofs, p2.vtx_xyz):
                t_.record_stream(main)
            asm_e2e.part = p2
            r = asm_e2e.step(ph)
            rows_g, rp, cols_g, vals = DistributedCutAssembly.owned_rows_global(r)
            s_out.wait_stream(main)
            with torch.cuda.stream(s_out):
                outs = [(f"bits{i}", b) for i, b in enumerate(r["cls"][:3])] + [("cut_ratio", r["cls"][3]), ("rows", rows_g),
                                                                                 ("rowptr", rp), ("col", cols_g), ("vals", vals)]
                d2h = 0
                for name, t_ in outs:
                    t_.record_stream(s_out)
                    d2h += hb.land(name + par, t_)
                hb.done[hb.parity] = torch.cuda.Event()
                hb.done[hb.parity].record(s_out)
            return r, hb.h2d_bytes(), d2h

        def e2e_finish():
            main = torch.cuda.current_stream(dev)
            main.wait_stream(s_out)
            main.wait_stream(s_in)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps, warmup, finish=None):
        for _ in range(warmup):
            fn(None)
        if finish:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launches + (prob.ctx_hp.launches + prob.ctx2.launches if world == 1 else 0)
        e0.record()
        res = None
        for k in range(steps):
            res = fn(None)
        if finish:
            finish()   # the launching stream waits for the outstanding copies: they are inside the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        l1 = ctx.launches + (prob.ctx_hp.launches + prob.ctx2.launches if world == 1 else 0)
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, res, l1 - l0

    steps, warmup = args.steps, max(3, args.warmup)
    with ClockSampler(local) as cs:
        ms, res, launches = timed_loop(step_dev, steps, warmup)
    clocks = cs.summary()
    units = res["units"]
    # ---- correctness carried by the record (N > 1): the merged, row-distributed matrix against the sum of the ranks'
    # own element contributions, through position-dependent checksums over the global dof keys (distributed.merge_check)
    if world > 1:
        check = merge_check(asm, res, phi)
    # ---- phase shares: instrumented steps (CUDA events around every phase, one stream), right after the timed loop
    nphase = 3
    step_dev(PhaseTimer(True))   # untimed: the one-stream schedule allocates its own pool blocks on first use
    timers = [PhaseTimer(True) for _ in range(nphase)]
    for t in timers:
        res = step_dev(t)
    torch.cuda.synchronize()
    phases = {}
    for t in timers:
        for k, v in t.result().items():
            phases[k] = phases.get(k, 0.0) + v / nphase
    phases_max = None
    if os.environ.get("XGPU_DUMP_PHASES"):   # diagnosis: every rank's own phase times
        with open(os.path.join(ROOT, "gpurun_out", f"phases_rank{rank}_of{world}_{args.scaling}.json"), "w") as f_:
            json.dump({k: round(v, 4) for k, v in phases.items()}, f_)
    if world > 1:  # the slowest rank per phase (interior ranks have two neighbours)
        keys = sorted(phases)
        pv = torch.tensor([phases[k] for k in keys], device=dev, dtype=torch.float64)
        dist.all_reduce(pv, op=dist.ReduceOp.MAX)
        phases_max = {k: round(float(v), 4) for k, v in zip(keys, pv.tolist())}
    if "npts" not in res:
        res["npts"] = (0, 0)
        res["ndof"] = res["ndof_local"]
    elif res["npts"][0] is None:   # cut rules made on the fly: the point counts are reporting only, taken here (untimed)
        res["npts"] = prob.cut_point_counts(phi)
    kb_of = lambda k: kernel_bytes(k, res, nd, ne=ne_loc, ndof_full=ndof_full)
    dom = max((k for k in phases if kb_of(k)), key=lambda k: phases[k])
    for k in ("vals", "col", "rowptr"):
        res.pop(k, None)
    torch.cuda.empty_cache()

    # ---- e2e: host buffers in, host CSR out, copies inside the timed region
    e2e_steps = max(1, min(steps, 10))
    info = {}

    def e2e_fn(_):
        r, h2d, d2h = step_e2e()
        info.update(h2d=h2d, d2h=d2h)
        return r
    ms_e2e, _, _ = timed_loop(e2e_fn, e2e_steps, 3, finish=e2e_finish)   # warm-up steps: both landing sets get pinned

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
        kb = kb_of(dom)
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
        # every phase against the HBM roofline, and the whole step against the summed algorithmic bytes
        per_phase = {}
        tot_bytes = 0
        for k, v in phases.items():
            b_ = kb_of(k)
            if b_:
                tot_bytes += b_
                per_phase[k] = {"ms": round(v, 4), "GB": round(b_ / 1e9, 3), "GB/s": round(b_ / (v * 1e-3) / 1e9, 1),
                                "frac": round(b_ / (v * 1e-3) / 1e9 / peak, 4)}
        step_ms = ms / steps
        roof_step = {"algorithmic_bytes": tot_bytes, "GB/s": tot_bytes / (step_ms * 1e-3) / 1e9,
                     "frac": tot_bytes / (step_ms * 1e-3) / 1e9 / peak}
        # the FP64 contraction of the cut elements (k_elmat_ms): ~950 flop per quadrature point
        roof_fp64 = None
        if "elmat_cut" in phases and sum(res["npts"]) > 0:
            pk64, src64 = fp64_peak()
            fl = sum(res["npts"]) * 950.0
            a64 = fl / (phases["elmat_cut"] * 1e-3) / 1e12
            roof_fp64 = {"kernel": "elmat_cut (k_elmat_cut: cut rule + contraction, one lane per point)", "bound": "fp64", "achieved": a64, "peak": pk64, "unit": "TFLOP/s",
                         "frac": a64 / pk64, "peak_source": src64, "flops_per_launch": fl, "ms_per_launch": phases["elmat_cut"]}
        cpu_baseline = None
        if world == 1 and not args.no_cpu:  # the CPU arm is timed on rank 0 at N=1 only (other ranks would compete for the host cores)
            cpu_val, cores, cinfo = cpu_sample(n, 1, 1, budget_s=60.0)
            sample = (f"1 timed pass of the full workload after 1 warm-up pass (n={n}: {6 * n ** 3} tets, "
                      f"{cinfo['units_per_step']} matrices per pass, {cinfo['ms_per_step'] / 1e3:.1f} s)")
            cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "phases_s": cinfo["phases_s"]}
        wl = workload_name(n)
        if world > 1 and args.scaling == "weak":
            wl += (f"; x{world} ranks: chain of {world} unit cubes, spheres centred on the slab interfaces (same per-rank element "
                   "counts), halo layer + NCCL interface-row merge")
        elif world > 1:
            wl += (f"; the SAME cube cut into {world} z-slabs (contiguous element ranges), halo layer + NCCL interface-row merge")
        line = {
            "metric": METRIC, "value": tot_units * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl,
                       "units_per_step_per_gpu": units, "n_neg": res["n_neg"], "n_if": res["n_if"],
                       "ndof": res["ndof"], "nnz": res["nnz"], "cut_points": list(res["npts"]),
                       "l2_policy": "inputs+outputs of a step (>5 GB) exceed the 126 MB L2; no flush needed",
                       "timed_region": ("production step, one stream" if not conc else "production step (cut-element branch on a low-priority side stream)") if world == 1 else "distributed step",
                       "phases_ms": {k: round(v, 4) for k, v in phases.items()},
                       "phases_note": "instrumented steps on one stream after the timed loop; they sum to slightly more than ms_per_step",
                       **({"phases_ms_max_over_ranks": phases_max} if phases_max else {})},
            "e2e": {"value": tot_units * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / e2e_steps,
                    "steps": e2e_steps, "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"],
                    "pipelining": ("steps software-pipelined: H2D + kernels of step k+1 overlap the D2H of step k (two sets of pinned "
                                   "landing buffers)")},
            "gpu_launches": int(launches),
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kb,
                         "ms_per_launch": phases[dom],
                         "bytes_rule": "SURVEY.md §8(d): units x (nd^2*8 + 2*nd*4) + 12 B per stored nonzero" if dom == "csr_rows_fused" else "DESIGN.md §4"},
            "roofline_phases": per_phase,
            "roofline_step": roof_step,
            "roofline_fp64": roof_fp64,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
        }
        if check is not None:
            line["check"] = check
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if check is not None and not check.get("ok", False):
        raise SystemExit("bench.py: the merged distributed matrix does not match the ranks' own contributions: " + json.dumps(check))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = chain of N unit cubes (one per rank); strong = the config-2 cube cut into N slabs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
