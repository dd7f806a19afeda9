"""BASELINE config 5 element-partitioned: one space-time slab (moving spheres, P1 space x P2 time, 6 n^3 tets per rank,
n = 102 -> 6 367 248) assembled on N GPUs — slab CutInfo, Restrict, space-time rules + element matrices (u v + dt(u) v +
grad u . grad v), local CSR, NCCL interface rows.  Weak scaling on the chain of unit cubes, device-timed, max over ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/perf_config5_dist.py [--n 102] [--steps 5] [--warmup 3]
Prints one JSON line on rank 0 (a tool, not the bench contract: bench.py measures config 2)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngsxfem_b200 import ops
from ngsxfem_b200.distributed import DistributedCutAssembly, SpaceTimeGpuBackend, make_slab_partition
from ngsxfem_b200.pipeline import PhaseTimer

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=102)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
ctx = ops.default_ctx(local)
P, KT = 1, 2
part, phi_st = make_slab_partition(args.n, world, rank, dev, fe_order=P, radius=0.3, time_fe_order=KT)
asm = DistributedCutAssembly(part, SpaceTimeGpuBackend(P, KT, KT, 4, ctx).bind(part), rank, world)
for _ in range(max(3, args.warmup)):
    asm.step(phi_st)
dist.barrier(); torch.cuda.synchronize()
timers = [PhaseTimer(True) for _ in range(args.steps)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ctx.launches
e0.record()
for k in range(args.steps):
    res = asm.step(phi_st, timer=timers[k])
e1.record()
dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
units = torch.tensor([res["units"]], device=dev, dtype=torch.int64)
dist.all_reduce(units)
phases = {}
for t in timers:
    for k, v in t.result().items():
        phases[k] = phases.get(k, 0.0) + v / args.steps
keys = sorted(phases)
pv = torch.tensor([phases[k] for k in keys], device=dev, dtype=torch.float64)
dist.all_reduce(pv, op=dist.ReduceOp.MAX)
# size-independent check: the all-ones vector is the constant 1 (P1 x nodal time): 1^T A 1 = space-time NEG volume summed over
# the owned rows of all ranks (dt and stiffness terms annihilate constants) = world x (ball of radius 0.3), polyhedral
rows_g, rp, cols_g, vals = DistributedCutAssembly.owned_rows_global(res)
tot = vals.sum().reshape(1).clone()
dist.all_reduce(tot)
if rank == 0:
    ball = 4.0 / 3.0 * 3.141592653589793 * 0.3 ** 3
    vol = float(tot.item())
    assert abs(vol - world * ball) < 2e-2 * world * ball, (vol, world * ball)
    print(json.dumps({"workload": f"config 5: space-time slab, P1 x P2, {6 * args.n ** 3} tets per rank, moving spheres r=0.3, "
                                  "u v + dt(u) v + grad u.grad v on NEG, time_order 4",
                      "n_gpus": world, "steps": args.steps, "ms_per_step": float(ms.item()) / args.steps,
                      "matrices_per_s": int(units.item()) * args.steps / (float(ms.item()) * 1e-3),
                      "units_total": int(units.item()), "ndof_global": res["ndof_global"],
                      "sum_of_entries": vol, "expected_volume": world * ball,
                      "phases_ms_max_over_ranks": {k: round(float(v), 4) for k, v in zip(keys, pv.tolist())},
                      "gpu_launches": int(ctx.launches - l0)}), flush=True)
dist.destroy_process_group()
