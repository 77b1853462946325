"""
Panel-cyclic multi-GPU Cholesky vs the replicated one, under torchrun (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_dist_cholesky.py [n ...]
For each size n: a row-sharded Gram matrix of a CheckerBoard sample (n points, forces at the data) is
all-reduced, then solved twice from the same matrix -- replicated (vb200_gram_solve on every rank) and
panel-cyclic (GramSystem.solve_distributed) -- and the forces are compared.  Rank 0 prints one JSON
object per size: device times (max over ranks) and the largest difference.
"""
import json, os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch
import torch.distributed as td
import verde_b200 as vb
from verde_b200 import kernels

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
vb._cabi.require_device().vb200_set_device(local)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
td.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = [int(a) for a in sys.argv[1:]] or [3000, 20000]
damping, mindist = 1e-6, 1e-5
ok = True


def timed(fn):
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    td.all_reduce(dt, op=td.ReduceOp.MAX)
    return float(dt.item()), out


for n in sizes:
    tab = vb.CheckerBoard().scatter(size=n, random_state=0)
    e, nn, d = tab.easting.values, tab.northing.values, tab.scalars.values
    a, b = vb.dist.shard_bounds(n)
    system = kernels.GramSystem(n)
    system.accumulate(0, e[a:b], nn[a:b], d[a:b], None, e, nn, mindist)
    t_allreduce, _ = timed(system.allreduce)
    keep = system.gram.clone()
    res = {"n": n, "world": world, "allreduce_s": t_allreduce}
    for rep in range(2):  # second repetition is the warm one
        system.gram.copy_(keep)
        res["replicated_s"], f_rep = timed(lambda: system.solve(n, damping))
        res["replicated_factor_ms"] = system.info.factor_ms
        system.info = vb._cabi.FitInfo()
        system.gram.copy_(keep)
        for exchange in ("broadcast", "allreduce"):
            vb.dist.set_panel_exchange(exchange)
            system.gram.copy_(keep)
            phases = {}
            res[exchange + "_s"], f_x = timed(lambda: system.solve_distributed(n, damping, phase_times=phases))
            # this rank's device time per phase: "broadcast" includes waiting for the owner's panel
            res[exchange + "_rank0_phase_ms"] = {k: round(v, 2) for k, v in phases.items()}
            res[exchange + "_factor_ms"] = system.info.factor_ms
            system.info = vb._cabi.FitInfo()
            if exchange == "broadcast":
                f_dist = f_x
            else:
                res["allreduce_exchange_bit_identical"] = bool(np.array_equal(f_x, f_dist))
        vb.dist.set_panel_exchange("broadcast")
        res["distributed_s"], res["distributed_factor_ms"] = res["broadcast_s"], res["broadcast_factor_ms"]
    # every rank must hold the same answer
    mine = torch.from_numpy(f_dist).cuda()
    lo, hi = mine.clone(), mine.clone()
    td.all_reduce(lo, op=td.ReduceOp.MIN); td.all_reduce(hi, op=td.ReduceOp.MAX)
    res["ranks_agree"] = bool((lo == hi).all().item())
    res["dist_vs_replicated_rel"] = float(np.max(np.abs(f_dist - f_rep)) / np.max(np.abs(f_rep)))
    res["bit_identical"] = bool(np.array_equal(f_dist, f_rep))
    ok = ok and res["ranks_agree"] and res["dist_vs_replicated_rel"] < 1e-6
    if rank == 0:
        print(json.dumps(res), flush=True)
    del system, keep
td.destroy_process_group()
sys.exit(0 if ok else 1)
