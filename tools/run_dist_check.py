"""
Multi-GPU functional check, launched under torchrun (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_dist_check.py
Compares the sharded paths (row-sharded fit + all-reduce, point-sharded predict + all-gather,
SplineCV job deal + score all-reduce, sharded VectorSpline2D) with the golden vectors of the
reference and with the single-GPU path; rank 0 prints one JSON object.
"""
import json, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch
import torch.distributed as td
import verde_b200 as vb

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
vb._cabi.require_device().vb200_set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"world": world}
g = np.load(os.path.join(ROOT, "tests/golden/spline_small.npz"))
e, n, d, w = g["east"], g["north"], g["data"], g["weights"]
assert vb.dist.sharding_active()
sp = vb.Spline(damping=1e-4).fit((e, n), d, weights=w)
out["spline_force_vs_ref"] = float(np.max(np.abs(sp.force_ - g["force_d1em4_w"])) / np.max(np.abs(g["force_d1em4_w"])))
out["spline_grid_vs_ref"] = float(np.max(np.abs(sp.predict((g["grid_east"], g["grid_north"])) - g["grid_d1em4_w"])) / np.ptp(d))
vb.dist.set_sharding(False)
single = vb.Spline(damping=1e-4).fit((e, n), d, weights=w)
vb.dist.set_sharding("auto")
out["spline_sharded_vs_single"] = float(np.max(np.abs(sp.force_ - single.force_)) / np.max(np.abs(single.force_)))

# the same fit with the panel-cyclic Cholesky forced on (tiny system: 2 outer panels of 1 block column)
vb.dist.set_cholesky("distributed")
spd = vb.Spline(damping=1e-4).fit((e, n), d, weights=w)
vb.dist.set_cholesky("auto")
out["spline_distributed_cholesky_vs_replicated"] = float(np.max(np.abs(spd.force_ - sp.force_)) / np.max(np.abs(sp.force_)))

v = np.load(os.path.join(ROOT, "tests/golden/vector_small.npz"))
vs = vb.VectorSpline2D(damping=1e-3).fit((v["east"], v["north"]), (v["data_east"], v["data_north"]))
out["vector_force_vs_ref"] = float(np.max(np.abs(vs.force_ - v["force_d1em3"])) / np.max(np.abs(v["force_d1em3"])))
pe, pn = vs.predict((v["grid_east"], v["grid_north"]))
out["vector_grid_vs_ref"] = float(max(np.max(np.abs(pe - v["grid_east_d1em3"])), np.max(np.abs(pn - v["grid_north_d1em3"]))))

c = np.load(os.path.join(ROOT, "tests/golden/splinecv_small.npz"))
cv = vb.SplineCV(dampings=list(c["dampings"]), mindists=list(c["mindists"])).fit((c["east"], c["north"]), c["data"])
out["cv_scores_vs_ref"] = float(np.max(np.abs(cv.scores_ - c["scores"])))
out["cv_best"] = [float(cv.mindist_), float(cv.damping_), float(c["best_mindist"]), float(c["best_damping"])]
out["cv_grid_vs_ref"] = float(np.max(np.abs(cv.predict((c["grid_east"], c["grid_north"])) - c["grid"])) / np.ptp(c["data"]))
ok = (out["spline_distributed_cholesky_vs_replicated"] < 1e-9 and out["spline_force_vs_ref"] < 1e-5 and out["spline_grid_vs_ref"] < 1e-6 and out["vector_force_vs_ref"] < 1e-6
      and out["cv_scores_vs_ref"] < 1e-7 and out["cv_best"][:2] == out["cv_best"][2:] and out["cv_grid_vs_ref"] < 1e-6)
out["ok"] = bool(ok)
if rank == 0:
    print(json.dumps(out), flush=True)
td.destroy_process_group()
sys.exit(0 if ok else 1)
