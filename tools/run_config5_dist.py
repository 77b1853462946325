"""
BASELINE config 5 (SplineCV: 6 dampings x 4 mindists x 5 folds on 20 000 points) with the (fold, mindist)
Gram jobs dealt across the ranks of a torchrun launch; rank 0 prints one JSON object.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config5_dist.py
"""
import json, os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")
import torch
import torch.distributed as td
import verde_b200 as vb

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
vb._cabi.require_device().vb200_set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
tab = vb.CheckerBoard().scatter(size=20000, random_state=0)
coords, d = (tab.easting.values, tab.northing.values), tab.scalars.values
dampings = [1e-8, 1e-6, 1e-4, 1e-3, 1e-2, 1e-1]
mindists = [0, 1e-5, 1e-3, 1e-1]
times = []
for rep in range(2):
    cv = vb.SplineCV(dampings=dampings, mindists=mindists)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    t0 = time.perf_counter()
    cv.fit(coords, d)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    times.append(time.perf_counter() - t0)
if rank == 0:
    print(json.dumps({"config": 5, "n_gpus": world, "fit_s": min(times), "fit_s_all": times, "best": [float(cv.mindist_), float(cv.damping_)],
                      "scores": [float(s) for s in cv.scores_]}), flush=True)
if world > 1:
    td.destroy_process_group()
