#!/usr/bin/env python
"""
bench.py -- headline benchmark of the Green's-function spline hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size n]

A STEP is one pass of the hot path over one batch of synthetic CheckerBoard
input: ``Spline(damping=1e-6, mindist=1e-5).fit`` followed by the grid predict.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
  N = 1   50 000 scattered points, forces at the data (50k x 50k FP64 normal
          matrix + Cholesky on one B200), grid 2000 x 2000.
  N > 1   weak scaling towards BASELINE configs[2]: the first 50 000 * N points of
          config 3's 400 000-point CheckerBoard scatter (rows sharded across ranks,
          one NCCL all-reduce of the Gram matrix), weights RandomState(1).uniform(0.1, 10),
          forces at the 50 151 BlockReduce(np.mean, shape=(224, 224)) block centres of
          the 400 000 points (computed on the device), grid 2000 x (2000 * N) points
          sharded across ranks.  N = 8 is configs[2]'s own 400k x 50k system.
          Damping = 1e-6 * sum(weights) / 50 000 (configs[1]'s relative damping, see
          config3_damping: cond(A) <= 4.6e15 at every N; at a flat 1e-6 the weighted
          200k- and 400k-row systems are not positive definite in FP64).
          Outside the timed region every rank also runs the sharded estimator API
          (Spline, VectorSpline2D, SplineCV) on the reference's golden cases and rank 0
          reports the gaps in "parity".

Metric: grid points predicted per second of whole-step time (fit + predict),
whole job.  The JSON line also carries the two quantities BASELINE.json names
separately (``fit_s``, ``predict_gpts_per_s``) and the FP64 roofline fraction.

value  -- inputs resident in HBM (device pointers through the C ABI).
e2e    -- the estimator API (``Spline.fit`` / ``Spline.predict``) with pinned HOST
          arrays: H2D of the inputs and D2H of forces and grid inside the timing.
Timing -- CUDA events on the launching (legacy default) stream, barrier +
          synchronize on both sides, max over ranks.  Inputs are far larger than
          L2 (10 GB Gram matrix, 400+ MB panels), so no explicit L2 flush is used.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings


def _wants_reference_arm(argv):
    return any(a == "reference" or a.endswith("=reference") for a in argv)


def host_cores():
    "Cores this process may run on (the launcher's affinity mask, not the machine's core count)."
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank it starts.  The CPU arm must use ALL host
# cores however it was launched, and OpenBLAS / libgomp read these variables when they are loaded -- so they
# are set here, before numpy is imported (force_cpu_threads() repeats it at run time through threadpoolctl
# for libraries a site hook may have loaded earlier).
if _wants_reference_arm(sys.argv[1:]):
    for _var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_var] = str(host_cores())

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spline_fit_plus_grid_predict_points_per_s@50k_forces"
UNIT = "Mpts/s"
DAMPING, MINDIST = 1e-6, 1e-5
NCU_GRAM_LAUNCH_TRAFFIC_BYTES = 36.87e9  # profiles/r02_ncu_full_summary.json gemm_dynamic_gram_launch (26.85 GB read + 10.02 GB written)
REGION = (0, 5000, -5000, 0)


# ---------------------------------------------------------------------------------------
# synthetic workload (CheckerBoard defaults of verde/synthetic.py:65-118, scatter of
# verde/coordinates.py:183-186: easting then northing from RandomState(seed).uniform)
# ---------------------------------------------------------------------------------------
CONFIG3_POINTS, CONFIG3_BLOCKS = 400_000, (224, 224)


def checkerboard(east, north):
    return 1000.0 * np.sin((2 * np.pi / 2500.0) * east) * np.cos((2 * np.pi / 2500.0) * north)


def make_workload(n_data, n_forces, grid_shape, forces_at_data):
    "configs[1]-style workload: scatter(size=n_data, random_state=0); separate forces from a second scatter."
    rng = np.random.RandomState(0)
    east = rng.uniform(REGION[0], REGION[1], n_data)
    north = rng.uniform(REGION[2], REGION[3], n_data)
    data = checkerboard(east, north)
    if forces_at_data:
        fe, fn = east, north
    else:
        rng1 = np.random.RandomState(1)
        fe = rng1.uniform(REGION[0], REGION[1], n_forces)
        fn = rng1.uniform(REGION[2], REGION[3], n_forces)
    ge = np.linspace(east.min(), east.max(), grid_shape[1])
    gn = np.linspace(north.min(), north.max(), grid_shape[0])
    return east, north, data, None, fe, fn, ge, gn


def config3_points():
    "BASELINE configs[2]'s data: CheckerBoard().scatter(size=400_000, random_state=0) and its weights."
    rng = np.random.RandomState(0)
    east = rng.uniform(REGION[0], REGION[1], CONFIG3_POINTS)
    north = rng.uniform(REGION[2], REGION[3], CONFIG3_POINTS)
    weights = np.random.RandomState(1).uniform(0.1, 10, CONFIG3_POINTS)
    return east, north, checkerboard(east, north), weights


def config3_damping(n_data):
    """
    Damping of the N > 1 workload.  BASELINE configs[2] names no damping, and configs[1]'s 1e-6 does not carry over:
    lambda_max(A) = 9.2e4 * sum(w) grows with the rows and their weights, and with the config's weights the 200 000- and
    400 000-row systems are NOT numerically positive definite in FP64 at 1e-6 (measured: first non-positive pivots 19 829
    and 4 760, profiles/r02l_config3_conditioning.jsonl; the reference's Ridge would fall back to an SVD solve there).
    The damping is therefore scaled with the total weight, damping = 1e-6 * sum(w) / 50 000 -- configs[1]'s RELATIVE
    damping, so that cond(A) <= lambda_max / damping = 4.6e15 at every N, the conditioning configs[1] is quoted at.
    """
    weights = np.random.RandomState(1).uniform(0.1, 10, CONFIG3_POINTS)
    return DAMPING * float(np.sum(weights[:n_data])) / 50_000.0


def make_config3_workload(n_data, grid_shape, block_reduce):
    """
    The first *n_data* points of config 3's scatter (all 400 000 at N = 8) with the config's weights; forces at the
    block centres of ALL 400 000 points -- *block_reduce* is the device BlockReduce on the GPU arm and the numpy
    block oracle on the CPU arm (they agree: tests/test_gpu_blocks.py::test_config3_force_positions).
    """
    east, north, data, weights = config3_points()
    fe, fn = block_reduce(east, north, data)
    ge = np.linspace(east.min(), east.max(), grid_shape[1])
    gn = np.linspace(north.min(), north.max(), grid_shape[0])
    return east[:n_data], north[:n_data], data[:n_data], weights[:n_data], fe, fn, ge, gn


class ClockSampler:
    "nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[0]))
                mx.append(float(row[1]))
                power.append(float(row[2]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(names, row[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy/C restatement of the reference) on a bounded sample
# ---------------------------------------------------------------------------------------
def cpu_reference_step(sample_n, sample_grid, full, weighted=False):
    """
    Time the reference's CPU algorithm (oracle/ restatement: OpenMP Green's kernels +
    numpy/scipy normal equations, the same BLAS/LAPACK calls sklearn makes) on a
    bounded sample and extrapolate to the full workload with the cost model of
    SURVEY.md 8d: Jacobian ~ n*m, scaler ~ n*m, Gram ~ n*m^2, Cholesky ~ m^3/3,
    predict ~ N*M.
    """
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import scipy.linalg
    import verde_oracle as vo

    east, north, data, _, _, _, _, _ = make_workload(sample_n, sample_n, (8, 8), True)
    t0 = time.perf_counter()
    jac = vo.jacobian(east, north, east, north, MINDIST)
    t_jac = time.perf_counter() - t0
    t0 = time.perf_counter()
    scale = vo.column_scale(jac)
    jac /= scale
    t_scale = time.perf_counter() - t0
    t0 = time.perf_counter()
    if weighted:  # sklearn/linear_model/_base.py:223-278: rows of X and y rescaled by sqrt(w) (one more pass over J)
        sw = np.sqrt(np.random.RandomState(1).uniform(0.1, 10, sample_n))
        jac *= sw[:, None]
        data = data * sw
    gram = jac.T @ jac
    rhs = jac.T @ data
    t_gram = time.perf_counter() - t0
    t0 = time.perf_counter()
    gram.flat[:: sample_n + 1] += DAMPING
    coef = scipy.linalg.solve(gram, rhs, assume_a="pos", overwrite_a=True) / scale
    t_solve = time.perf_counter() - t0
    ge = np.linspace(east.min(), east.max(), sample_grid)
    gn = np.linspace(north.min(), north.max(), sample_grid)
    gx, gy = np.meshgrid(ge, gn)
    t0 = time.perf_counter()
    vo.predict(gx, gy, east, north, MINDIST, coef)
    t_pred = time.perf_counter() - t0

    n, m, npts = full
    s = float(sample_n)
    fit_full = (t_jac + t_scale) * (n * m) / (s * s) + t_gram * (n * m * m) / s**3 + t_solve * (m / s) ** 3
    pred_rate = sample_grid * sample_grid * s / t_pred  # pair evaluations / s
    pred_full = npts * m / pred_rate
    return {
        "sample_seconds": t_jac + t_scale + t_gram + t_solve + t_pred,
        "fit_s_extrapolated": fit_full,
        "predict_s_extrapolated": pred_full,
        "predict_gpairs_per_s": pred_rate * 1e-9,
        "value": npts / (fit_full + pred_full) * 1e-6,
        "phases_s": {"jacobian": t_jac, "scaler": t_scale, "gram": t_gram, "solve": t_solve, "predict": t_pred},
    }


def force_cpu_threads():
    """
    Make BLAS/LAPACK and the oracle's OpenMP kernels use every host core regardless of the launcher
    (torchrun sets OMP_NUM_THREADS=1) and report the thread counts actually in force.
    """
    cores = host_cores()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import verde_oracle as vo

    vo._lib()  # load the OpenMP Green's kernels first so that threadpoolctl sees their libgomp
    pools = {}
    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=cores)
        for pool in threadpoolctl.threadpool_info():
            pools["{}:{}".format(pool.get("user_api"), pool.get("internal_api"))] = pool.get("num_threads")
    except ImportError:  # pragma: no cover
        pools["threadpoolctl"] = "missing"
    return cores, pools


def reference_sample_plan(args):
    """
    The CPU arm times a bounded sample (n = m = --cpu-sample, ~35 s on 16 cores at the default 20 000)
    and repeats it while the wall-clock budget lasts: at least 3, at most --steps samples; the value printed
    is the MEDIAN over those samples.
    """
    return max(1, min(args.steps, 3)), args.steps, float(args.cpu_budget)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n = args.size * world
    npts = args.grid * args.grid * world
    cores, pools = force_cpu_threads()
    weighted = world > 1
    if world > 1:
        # the same force positions the GPU arm uses: block centres of config 3's 400 000 points (numpy block oracle)
        import block_oracle as bo

        ce, cn, _, _ = config3_points()
        _, labels = bo.block_split((ce, cn), shape=CONFIG3_BLOCKS)
        m = int(np.unique(labels).size)
    else:
        m = args.forces if args.forces else n
    sample_n, sample_grid = args.cpu_sample, 150
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_step(min(1500, sample_n), 40, (n, m, npts), weighted)
    at_least, at_most, budget = reference_sample_plan(args)
    times, samples = [], []
    t_begin = time.perf_counter()
    while len(samples) < at_most:
        t0 = time.perf_counter()
        samples.append(cpu_reference_step(sample_n, sample_grid, (n, m, npts), weighted))
        times.append(time.perf_counter() - t0)
        if len(samples) >= at_least and (time.perf_counter() - t_begin) + times[-1] > budget:
            break
    order = sorted(range(len(samples)), key=lambda i: samples[i]["value"])
    mid = samples[order[len(order) // 2]]  # the median sample (upper median for an even count)
    values = [x["value"] for x in samples]
    sample = ("oracle (C/OpenMP Green's kernels + numpy/scipy BLAS/LAPACK normal equations) on n=m={} "
              "points and a {}x{} grid; extrapolated to n={}, m={}, {} grid points with n*m, n*m^2, "
              "m^3/3 and N*M scaling; median of {} samples").format(sample_n, sample_grid, sample_grid, n, m, npts,
                                                                   len(samples))
    line = {
        "impl": "reference", "metric": METRIC, "value": mid["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (mid["fit_s_extrapolated"] + mid["predict_s_extrapolated"]),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world, args.size, m, args.grid),
        "fit_s": mid["fit_s_extrapolated"], "predict_gpts_per_s": npts / mid["predict_s_extrapolated"] * 1e-9,
        "cpu_baseline": {"value": mid["value"], "unit": UNIT, "cores": cores, "threads": pools, "kind": "port",
                         "sample": sample, "samples_timed": len(samples), "sample_wall_s": statistics.median(times),
                         "values_min_max": [min(values), max(values)], "phases_s": mid["phases_s"],
                         "extrapolated": True},
        "e2e": {"value": mid["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, n_rank=50_000, m=50_000, grid=2000):
    if world > 1:
        full = (n_rank * world, grid) == (CONFIG3_POINTS, 2000)
        return {
            "workload": ("{}: Spline(damping=1e-6*sum(w)/50000, mindist=1e-5), the first {} of config 3's 400000 CheckerBoard points "
                         "(rows sharded over {} ranks), weights RandomState(1).uniform(0.1, 10), {} forces at the "
                         "BlockReduce(mean, shape=(224, 224)) centres of the 400000 points, grid {}x{} sharded".format(
                             "BASELINE configs[2]'s 400k x 50k system" if full else "weak-scaled towards BASELINE configs[2]",
                             n_rank * world, world, m, grid, grid * world)),
            "n_data": n_rank * world, "n_forces": m, "grid_points": grid * grid * world, "weights": True,
            "damping": config3_damping(n_rank * world), "mindist": MINDIST,
            "damping_rule": "1e-6 * sum(weights) / 50000: configs[1]'s relative damping, cond(A) <= 4.6e15 at every N "
                            "(at 1e-6 the weighted 200k/400k-row systems are not positive definite in FP64)",
            "parallelism": "rows+grid sharded x{}".format(world),
            "l2": "inputs larger than L2 (10 GB Gram matrix streamed); no explicit flush",
        }
    base = "BASELINE configs[1]" if (n_rank, m, grid) == (50_000, 50_000, 2000) else "REDUCED-SIZE variant of configs[1]"
    if m != n_rank:
        base = ("BASELINE configs[2] on one GPU" if (n_rank, m, grid) == (400_000, 50_000, 10_000)
                else "configs[2]-style system (separate forces)")
    return {
        "workload": "{}: Spline(damping=1e-6, mindist=1e-5), CheckerBoard scatter of {} points, {}, grid {}x{}".format(
            base, n_rank, "forces at data" if m == n_rank else "{} separate forces".format(m), grid, grid),
        "n_data": n_rank, "n_forces": m, "grid_points": grid * grid, "weights": False,
        "damping": DAMPING, "mindist": MINDIST, "parallelism": "rows+grid sharded x1",
        "l2": "inputs larger than L2 (10 GB Gram matrix streamed); no explicit flush",
    }


# ---------------------------------------------------------------------------------------
# N > 1: the sharded estimator API against the reference's golden vectors (outside the timed region)
# ---------------------------------------------------------------------------------------
def multi_gpu_parity(vb, world):
    """
    Every rank runs the SAME estimator calls a user would (torch.distributed initialised -> Spline/VectorSpline2D
    fits are row-sharded with one Gram all-reduce and the panel-cyclic Cholesky where it applies, predicts are
    point-sharded, SplineCV deals its (fold, mindist) jobs round-robin) and compares with tests/golden/*.npz,
    which were produced by the UNMODIFIED reference (oracle/make_golden*.py).  Collective: all ranks must call it.
    """
    golden = os.path.join(ROOT, "tests", "golden")
    eps = np.finfo(np.float64).eps
    out = {"world": world, "against": "golden vectors of the unmodified reference (tests/golden)"}

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # small goldens: n = 240 weighted Spline, 150-station VectorSpline2D, SplineCV 4 x 3 x 5 on 300 points
        g = np.load(os.path.join(golden, "spline_small.npz"))
        e, n, d, w = g["east"], g["north"], g["data"], g["weights"]
        sp = vb.Spline(damping=1e-4).fit((e, n), d, weights=w)
        out["spline_small_force"] = rel(sp.force_, g["force_d1em4_w"])
        out["spline_small_grid"] = float(np.max(np.abs(sp.predict((g["grid_east"], g["grid_north"])) - g["grid_d1em4_w"])) / np.ptp(d))
        vb.dist.set_sharding(False)
        single = vb.Spline(damping=1e-4).fit((e, n), d, weights=w)
        vb.dist.set_sharding("auto")
        out["sharded_vs_single_force"] = rel(sp.force_, single.force_)
        vb.dist.set_cholesky("distributed")  # tiny system: panel-cyclic factorisation forced on
        spd = vb.Spline(damping=1e-4).fit((e, n), d, weights=w)
        vb.dist.set_cholesky("auto")
        out["distributed_vs_replicated_cholesky_force"] = rel(spd.force_, sp.force_)
        # Vector([Spline, Spline]) under sharding: one row-sharded Gram matrix + one factor for both components,
        # one point-sharded Green's evaluation for both predictions; component 0 against the reference golden,
        # component 1 against its own (sharded) independent fit
        d2 = np.cos(e / 700.0) * 300.0 + n * 0.05
        vec = vb.Vector([vb.Spline(damping=1e-4), vb.Spline(damping=1e-4)]).fit((e, n), (d, d2), weights=(w, w))
        alone = vb.Spline(damping=1e-4).fit((e, n), d2, weights=w)
        grid = (g["grid_east"], g["grid_north"])
        pred = vec.predict(grid)
        out["vector_of_splines_shared_factor"] = {
            "component0_force_vs_golden": rel(vec.components[0].force_, g["force_d1em4_w"]),
            "component1_force_vs_independent_fit": rel(vec.components[1].force_, alone.force_),
            "component0_grid_vs_golden": float(np.max(np.abs(pred[0] - g["grid_d1em4_w"])) / np.ptp(d)),
            "component1_grid_vs_independent_fit": float(np.max(np.abs(pred[1] - alone.predict(grid))) / np.ptp(d2)),
            "shared": bool(vec._splines_share_forces())}
        v = np.load(os.path.join(golden, "vector_small.npz"))
        vs = vb.VectorSpline2D(damping=1e-3).fit((v["east"], v["north"]), (v["data_east"], v["data_north"]))
        out["vector_small_force"] = rel(vs.force_, v["force_d1em3"])
        c = np.load(os.path.join(golden, "splinecv_small.npz"))
        cv = vb.SplineCV(dampings=list(c["dampings"]), mindists=list(c["mindists"])).fit((c["east"], c["north"]), c["data"])
        out["cv_small_scores"] = float(np.max(np.abs(cv.scores_ - c["scores"])))
        out["cv_small_best_matches"] = bool((cv.mindist_, cv.damping_) == (float(c["best_mindist"]), float(c["best_damping"])))

        # config-3 shape: 40 000 rows sharded x 5 041 BlockReduce-centre forces, weighted (cond 1.9e11 twin)
        g3 = np.load(os.path.join(golden, "cfg3_40k.npz"))
        tab = vb.CheckerBoard().scatter(size=40_000, random_state=0)
        e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
        w = np.random.RandomState(1).uniform(0.1, 10, e.size)
        fc = (g3["force_east"], g3["force_north"])
        full = vb.grid_coordinates((e.min(), e.max(), n.min(), n.max()), shape=(600, 600))
        sub = (full[0][::8, ::8], full[1][::8, ::8])
        sp = vb.Spline(damping=1e-2, mindist=1e-5, force_coords=fc).fit((e, n), d, weights=w)
        cond = float(g3["cond_d1em2"])
        out["cfg3_40k"] = {"cond": cond, "force": rel(sp.force_, g3["force_d1em2"]),
                           "grid_over_range": float(np.max(np.abs(sp.predict(sub) - g3["grid_sub_d1em2"])) / np.ptp(d)),
                           "force_tol": max(1e-9, 10 * cond * eps), "cond_est": sp.fit_info_.get("cond_est"),
                           "backward_err": sp.fit_info_.get("backward_err")}

        # config-4 shape: VectorSpline2D on 3 000 stations (6 000 x 6 000), row-sharded, panel-cyclic Cholesky
        g4 = np.load(os.path.join(golden, "cfg4_3k.npz"))
        region = (0, 500e3, 0, 500e3)
        se, sn = vb.scatter_points(region, 3000, random_state=0)
        ue = 10 * np.sin(2 * np.pi * se / 3e5) * np.cos(2 * np.pi * sn / 4.5e5) + 3
        un = -7 * np.cos(2 * np.pi * se / 4.5e5) * np.sin(2 * np.pi * sn / 3e5) + 1
        vs = vb.VectorSpline2D(poisson=0.5, mindist=10e3, damping=1e-3).fit((se, sn), (ue, un))
        full = vb.grid_coordinates(region, shape=(200, 200))
        ge, gn = vs.predict((full[0][::5, ::5], full[1][::5, ::5]))
        cond = float(g4["cond_d1em3"])
        out["cfg4_3k"] = {"cond": cond, "force": rel(vs.force_, g4["force_d1em3"]),
                          "grid_over_range": max(float(np.max(np.abs(ge - g4["grid_east_sub_d1em3"])) / np.ptp(ue)),
                                                 float(np.max(np.abs(gn - g4["grid_north_sub_d1em3"])) / np.ptp(un))),
                          "force_tol": max(1e-9, 10 * cond * eps), "cond_est": vs.fit_info_.get("cond_est"),
                          "backward_err": vs.fit_info_.get("backward_err")}

        # config-5 shape: SplineCV 6 dampings x 4 mindists x 5 folds on 2 000 points, 20 Gram jobs dealt to the ranks
        g5 = np.load(os.path.join(golden, "cfg5_2k.npz"))
        tab = vb.CheckerBoard().scatter(size=2000, random_state=0)
        e, n, d = tab.easting.values, tab.northing.values, tab.scalars.values
        cv = vb.SplineCV(dampings=list(g5["dampings"]), mindists=list(g5["mindists"])).fit((e, n), d)
        diff = np.abs(cv.scores_ - g5["scores"]).reshape(len(g5["mindists"]), len(g5["dampings"]))
        out["cfg5_2k"] = {"scores_max_diff_by_damping": {str(float(x)): float(diff[:, k].max()) for k, x in enumerate(g5["dampings"])},
                          "best_matches": bool((cv.mindist_, cv.damping_) == (float(g5["best_mindist"]), float(g5["best_damping"]))),
                          "jobs_this_rank": int(cv.sweep_info_["jobs"])}
    vos = out["vector_of_splines_shared_factor"]
    ok = (out["spline_small_force"] < 1e-5 and out["spline_small_grid"] < 1e-6 and out["vector_small_force"] < 1e-6
          and vos["component0_force_vs_golden"] < 1e-5 and vos["component1_force_vs_independent_fit"] < 1e-5
          and vos["component0_grid_vs_golden"] < 1e-6 and vos["component1_grid_vs_independent_fit"] < 1e-7
          and out["distributed_vs_replicated_cholesky_force"] < 1e-9 and out["cv_small_scores"] < 1e-7
          and out["cv_small_best_matches"] and out["cfg3_40k"]["force"] <= out["cfg3_40k"]["force_tol"]
          and out["cfg3_40k"]["grid_over_range"] < 1e-6 and out["cfg4_3k"]["force"] <= out["cfg4_3k"]["force_tol"]
          and out["cfg4_3k"]["grid_over_range"] < 1e-6 and out["cfg5_2k"]["best_matches"]
          and max(v for k, v in out["cfg5_2k"]["scores_max_diff_by_damping"].items() if float(k) >= 1e-4) < 1e-7)
    out["ok"] = bool(ok)
    return out


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=50_000, help="data points per rank (default: the BASELINE 50k)")
    ap.add_argument("--grid", type=int, default=2000, help="grid is GRID x GRID per rank")
    ap.add_argument("--forces", type=int, default=None,
                    help="number of forces (default: forces at the data on one GPU, 50000 separate forces on several); "
                         "--gpus 1 --size 400000 --forces 50000 --grid 10000 is BASELINE configs[2] on ONE GPU")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="n=m of the CPU baseline sample")
    ap.add_argument("--cpu-budget", type=float, default=200.0,
                    help="--impl reference: wall-clock seconds after which no further sample is started")
    ap.add_argument("--cholesky", default="auto", choices=["auto", "replicated", "distributed"],
                    help="N > 1: replicated m x m Cholesky on every rank, or panel-cyclic over the ranks")
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed steps of the end-to-end arm (at most --steps)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-estimator parity block")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE",
                    help="library tunable (vb200_set_option), e.g. chol_lookahead=0 for a launch list whose DMMA launches "
                         "do not overlap panel kernels (ncu serialises concurrent kernels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # Libraries (NCCL's "NCCL version ..." banner) write to STDOUT; the contract is ONE JSON line there.
    # Park the real stdout and point fd 1 at stderr until the line is ready.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as td

    import verde_b200 as vb
    from verde_b200 import _cabi, kernels

    torch.cuda.set_device(local_rank)
    lib = _cabi.require_device()
    lib.vb200_set_device(local_rank)
    options = {}
    for item in args.option:
        key, _, val = item.partition("=")
        _cabi.check(lib.vb200_set_option(key.encode(), float(val)), "set_option " + item)
        options[key] = float(val)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints "NCCL version ..." on STDOUT at NCCL_DEBUG=VERSION; keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_rank = args.size
    n_total = n_rank * world
    damping = config3_damping(n_total) if world > 1 else DAMPING
    grid_shape = (args.grid, args.grid * world)
    npts_total = grid_shape[0] * grid_shape[1]
    if world > 1:
        def device_block_centres(ce, cn, cd):
            "force positions of config 3: BlockReduce on the device (vb_block.cu), as the config takes them"
            (be, bn), _ = vb.BlockReduce(np.mean, shape=CONFIG3_BLOCKS, center_coordinates=True).filter((ce, cn), cd)
            return be, bn

        vb.dist.set_sharding(False)  # the block reduction of the 400 000 points is replicated setup work
        east, north, data, weights, fe, fn, ge, gn = make_config3_workload(n_total, grid_shape, device_block_centres)
        vb.dist.set_sharding("auto")
        m = int(fe.size)
        forces_at_data = False
    else:
        m = args.forces if args.forces else args.size
        forces_at_data = m == n_total and not args.forces
        east, north, data, weights, fe, fn, ge, gn = make_workload(n_total, m, grid_shape, forces_at_data=forces_at_data)
    gx, gy = np.meshgrid(ge, gn)
    gx, gy = gx.ravel(), gy.ravel()
    a, b = vb.dist.shard_bounds(n_total, rank, world)
    pa, pb = vb.dist.shard_bounds(npts_total, rank, world)
    dev = torch.device("cuda", local_rank)

    def to_dev(x):
        return None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    d_e, d_n, d_d = to_dev(east[a:b]), to_dev(north[a:b]), to_dev(data[a:b])
    d_w = to_dev(None if weights is None else weights[a:b])
    d_fe, d_fn = to_dev(fe), to_dev(fn)
    d_gx, d_gy = to_dev(gx[pa:pb]), to_dev(gy[pa:pb])
    d_force = torch.empty(m, dtype=torch.float64, device=dev)
    d_out = torch.empty(pb - pa, dtype=torch.float64, device=dev)
    system = kernels.GramSystem(m) if world > 1 else None
    vb.dist.set_cholesky(args.cholesky)
    panel_tiles = vb.dist.panel_tiles(world)
    dist_chol = world > 1 and vb.dist.use_distributed_cholesky((m + 127) // 128, panel_tiles, world)
    info = _cabi.FitInfo()
    phase = {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def device_step():
        "fit + predict with everything resident in HBM (device pointers through the C ABI)"
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        if world == 1:
            rc = lib.vb200_spline_fit(_cabi.ptr(d_e), _cabi.ptr(d_n), _cabi.ptr(d_d), None, b - a, _cabi.ptr(d_fe),
                                      _cabi.ptr(d_fn), m, MINDIST, damping, _cabi.ptr(d_force), ctypes.byref(info))
            _cabi.check(rc, "bench fit")
        else:
            ctypes.memset(ctypes.byref(info), 0, ctypes.sizeof(info))
            rc = lib.vb200_gram_accumulate(0, _cabi.ptr(d_e), _cabi.ptr(d_n), _cabi.ptr(d_d), _cabi.ptr(d_w), b - a,
                                           _cabi.ptr(d_fe), _cabi.ptr(d_fn), m, MINDIST, 0.0,
                                           _cabi.ptr(system.gram), _cabi.ptr(system.stats), 0, ctypes.byref(info))
            _cabi.check(rc, "bench gram")
            system.allreduce()
            if dist_chol:
                system.info = info
                system.solve_distributed(n_total, damping, panel_tiles=panel_tiles, out=d_force)
            else:
                rc = lib.vb200_gram_solve(_cabi.ptr(system.gram), _cabi.ptr(system.stats), m, n_total, damping, 0,
                                          _cabi.ptr(d_force), ctypes.byref(info))
                _cabi.check(rc, "bench solve")
                system.info = info
            if info.lambda_max > 0:  # diagnostics on: backward error of the row-sharded system (m doubles all-reduced)
                system.backward_error(0, d_e, d_n, d_d, d_w, d_fe, d_fn, MINDIST, 0.0, d_force, n_total, damping)
        ev[1].record()
        rc = lib.vb200_predict(_cabi.ptr(d_gx), _cabi.ptr(d_gy), pb - pa, _cabi.ptr(d_fe), _cabi.ptr(d_fn),
                               _cabi.ptr(d_force), m, MINDIST, _cabi.ptr(d_out))
        _cabi.check(rc, "bench predict")
        ev[2].record()
        torch.cuda.synchronize()
        phase["fit_ms"] = ev[0].elapsed_time(ev[1])
        phase["predict_ms"] = ev[1].elapsed_time(ev[2])
        phase["info"] = info.as_dict()

    def timed(fn, steps):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(ms, op=td.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ---------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.vb200_launch_count()
    fit_ms, pred_ms, gemm_ms, gemm_flops, gemm_launches = [], [], 0.0, 0.0, 0
    shared_ms, shared_flops, shared_launches = 0.0, 0.0, 0

    def counted_step():
        nonlocal gemm_ms, gemm_flops, gemm_launches, shared_ms, shared_flops, shared_launches
        device_step()
        fit_ms.append(phase["fit_ms"])
        pred_ms.append(phase["predict_ms"])
        gemm_ms += phase["info"]["gemm_ms"]
        gemm_flops += phase["info"]["gemm_flops"]
        gemm_launches += phase["info"]["gemm_launches"]
        shared_ms += phase["info"]["shared_gemm_ms"]
        shared_flops += phase["info"]["shared_gemm_flops"]
        shared_launches += phase["info"]["shared_gemm_launches"]

    total_ms = timed(counted_step, args.steps)
    launches = lib.vb200_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = npts_total / (ms_per_step * 1e-3) * 1e-6

    # ---- end-to-end arm: estimator API, pinned host arrays ----------------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(x):
            t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
            return t.numpy()

        h_e, h_n, h_d = pinned(east), pinned(north), pinned(data)
        h_w = None if weights is None else pinned(weights)
        h_gx, h_gy = pinned(gx), pinned(gy)
        force_coords = None if forces_at_data else (pinned(fe), pinned(fn))
        result = {}

        def api_step():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", FutureWarning)
                spline = vb.Spline(damping=damping, mindist=MINDIST, force_coords=force_coords)
            spline.fit((h_e, h_n), h_d, weights=h_w)
            result["grid"] = spline.predict((h_gx, h_gy))
            result["force"] = spline.force_

        # two warm-up steps: the sharded predict returns page-locked arrays from torch's caching host allocator, and
        # with the previous result still referenced while the next is made, two blocks alternate from the third step on
        for _ in range(max(1, min(args.warmup, 2))):
            api_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_ms = timed(api_step, e2e_steps) / e2e_steps
        h2d = 8 * ((b - a) * (3 if weights is None else 4) + (0 if forces_at_data else 2 * m) + 2 * m + 2 * (pb - pa) + m)
        d2h = 8 * (m + (pb - pa))
        e2e = {"value": npts_total / (e2e_ms * 1e-3) * 1e-6, "unit": UNIT, "ms_per_step": e2e_ms, "e2e_steps": e2e_steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "verde_b200.Spline.fit + Spline.predict (host numpy, pinned)"}
        # parity spot check of the two arms against each other (same forces, same grid)
        dev_force = d_force.cpu().numpy()
        e2e["force_vs_device_arm"] = float(np.max(np.abs(result["force"] - dev_force)) / np.max(np.abs(dev_force)))

    parity = None
    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(vb, world)

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    peak = lib.vb200_measure_fp64_peak_tflops(0)
    achieved = gemm_flops / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
    fit_s = statistics.mean(fit_ms) * 1e-3
    pred_s = statistics.mean(pred_ms) * 1e-3
    n_pairs = (pb - pa) * m
    algorithmic_fit_flops = (b - a) * m * (m + 1) + m**3 / 3.0 / (world if dist_chol else 1)  # this rank's share
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(world, n_rank, m, args.grid),  # identical to the reference arm's config
        "cholesky": ("single GPU" if world == 1 else
                     "panel-cyclic over {} ranks, outer panels of {} block columns, look-ahead, panel exchange: {}".format(
                         world, panel_tiles, {"peer": "copy-engine pushes into peer memory (CUDA IPC over NVLink) + flag words",
                                 "lookahead": "ncclBroadcast on a side stream", "serial": "ncclBroadcast"}[vb.dist.panel_pipeline()])
                     if dist_chol else
                     "replicated on every rank"),
        "fit_s": fit_s, "predict_s": pred_s, "predict_gpts_per_s": (pb - pa) * world / pred_s * 1e-9,
        "predict_gpairs_per_s": n_pairs / pred_s * 1e-9,
        "fit_phases_ms": {k: phase["info"][k] for k in ("gram_ms", "factor_ms", "solve_ms", "gemm_ms")},
        "fit_fp64_frac": algorithmic_fit_flops / fit_s * 1e-12 / peak if peak > 0 else None,
        "roofline": {
            "kernel": "vb_gemm_dynamic_kernel (persistent FP64 DMMA 128x128 tile kernel, TMA-bulk/mbarrier pipeline, dynamic "
                      "tile queue: Gram accumulation + Cholesky updates)",
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak > 0 else None,
            "traffic": NCU_GRAM_LAUNCH_TRAFFIC_BYTES if (n_rank, m) == (50_000, 50_000) else None,
            "traffic_note": "ncu dram__bytes_read+write of ONE Gram-accumulation launch of this kernel (76 636 tiles x K=2048 = "
                            "5.14e12 flop, 142.9 ms; algorithmic HBM bytes 21e9 = read+write the 10 GB triangle + the 0.8 GB "
                            "panel once); profiles/r02_ncu_full_summary.json gemm_dynamic_gram_launch",
            "peak_source": "FP64 DMMA.8x8x4 issue microbenchmark run in this process "
                           "(vb200_measure_fp64_peak_tflops; MEASURED_PEAKS.json has no FP64 entry)",
            "launches_per_step": gemm_launches / args.steps,
            "algorithmic_flops_per_step": gemm_flops / args.steps,
            "share_of_step": (gemm_ms + shared_ms) / total_ms,
            # look-ahead Cholesky: these launches of the same kernel ran on SMs - R of the SMs BESIDE the factorisation of
            # the next panel (a second grid joins their tile queue when the panel is done); their spans time a kernel
            # that owns part of the machine, so achieved/frac above are over the launches that own the GPU
            "beside_panel_factorisation": {
                "launches_per_step": shared_launches / args.steps, "flops_per_step": shared_flops / args.steps,
                "ms_per_step": shared_ms / args.steps,
                "achieved": shared_flops / (shared_ms * 1e-3) * 1e-12 if shared_ms > 0 else None,
                "achieved_all_launches": ((gemm_flops + shared_flops) / ((gemm_ms + shared_ms) * 1e-3) * 1e-12
                                          if gemm_ms + shared_ms > 0 else None)},
        },
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "diag_min_max": [phase["info"]["diag_min"], phase["info"]["diag_max"]],
        # SURVEY 8b out_diag{cond, backward_err}: computed inside every timed step (vb_diag.cu)
        "conditioning": {"cond_est": phase["info"]["cond_est"], "cond_upper": phase["info"]["cond_upper"],
                         "lambda_max": phase["info"]["lambda_max"], "lambda_min_est": phase["info"]["lambda_min_est"],
                         "backward_err": phase["info"]["backward_err"], "diag_ms_per_step": phase["info"]["diag_ms"],
                         "power_its": phase["info"]["power_its"], "inverse_its": phase["info"]["inverse_its"],
                         "note": "A = D G D + damping I (the matrix the reference's Ridge factors); cond_est <= cond_2(A) <= "
                                 "cond_upper; backward_err = ||Js^T W (d - Js c) - damping c|| / (lambda_max ||c|| + ||Js^T W d||)"},
    }
    if options:
        line["options"] = options
    if parity is not None:
        line["parity"] = parity
    if not args.no_cpu_baseline and world == 1:
        cores, pools = force_cpu_threads()
        cpu = cpu_reference_step(args.cpu_sample, 150, (n_total, m, npts_total))
        line["cpu_baseline"] = {
            "value": cpu["value"], "unit": UNIT, "cores": cores, "threads": pools, "kind": "port",
            "sample": ("oracle on n=m={} + 150x150 grid ({:.1f} s of CPU work), extrapolated to the full workload with "
                       "n*m, n*m^2, m^3/3, N*M scaling").format(args.cpu_sample, cpu["sample_seconds"]),
            "fit_s_extrapolated": cpu["fit_s_extrapolated"], "predict_gpairs_per_s": cpu["predict_gpairs_per_s"],
        }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
