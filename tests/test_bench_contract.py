"""
bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the
keys the driver reads, on the same metric / unit / config as the GPU arm.
"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_reference_arm(*extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--cpu-sample", "1200", *extra]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [line for line in proc.stdout.splitlines() if line.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    import bench

    line = run_reference_arm()
    assert line["impl"] == "reference"
    assert line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["dtype"] == "f64" and line["data"] == "synthetic"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cfg = line["config"]
    assert "workload" in cfg and cfg["n_data"] == 50_000 and cfg["n_forces"] == 50_000 and cfg["grid_points"] == 4_000_000
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and "n=m=1200" in cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                           "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_workload_generator_matches_checkerboard():
    import numpy as np

    import bench
    import verde_b200 as vb

    east, north, data, weights, fe, fn, ge, gn = bench.make_workload(3000, 3000, (10, 12), True)
    assert weights is None
    tab = vb.CheckerBoard().scatter(size=3000, random_state=0)
    np.testing.assert_array_equal(east, tab.easting.values)
    np.testing.assert_array_equal(north, tab.northing.values)
    np.testing.assert_allclose(data, tab.scalars.values, rtol=1e-13, atol=1e-10)
    assert fe is east and ge.size == 12 and gn.size == 10


def test_reference_arm_uses_all_cores_under_torchrun_environment():
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must still use every host core, on rank 0 of
    a 2-rank launch, on config 3's weak-scaled workload (weights, BlockReduce-centre forces), and say so."""
    import bench

    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "1",
           "--cpu-sample", "1000", "--cpu-budget", "1"]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip())
    cpu = line["cpu_baseline"]
    assert cpu["cores"] == bench.host_cores()
    assert cpu["threads"] and all(v == cpu["cores"] for v in cpu["threads"].values()), cpu["threads"]
    assert cpu["samples_timed"] == 3 and "median of 3 samples" in cpu["sample"]
    assert cpu["values_min_max"][0] <= line["value"] <= cpu["values_min_max"][1]
    cfg = line["config"]
    assert line["n_gpus"] == 2 and cfg["n_data"] == 100_000 and cfg["n_forces"] == 50_151 and cfg["weights"] is True
    assert cfg["grid_points"] == 8_000_000
    # the N > 1 damping follows the total weight (config 2's relative damping), identically in both arms
    assert abs(cfg["damping"] - bench.config3_damping(100_000)) < 1e-20 and "damping_rule" in cfg
    assert cfg == bench.workload_config(2, 50_000, 50_151, 2000)


def test_config3_damping_keeps_the_relative_damping_of_config_2():
    import numpy as np

    import bench

    w = np.random.RandomState(1).uniform(0.1, 10, 400_000)
    for n in (100_000, 200_000, 400_000):
        np.testing.assert_allclose(bench.config3_damping(n), 1e-6 * w[:n].sum() / 50_000, rtol=1e-14)
    assert 3.9e-5 < bench.config3_damping(400_000) < 4.2e-5


def test_config3_workload_generator():
    import numpy as np

    import bench

    calls = []

    def fake_reduce(e, n, d):
        calls.append(e.size)
        return e[:7], n[:7]

    east, north, data, weights, fe, fn, ge, gn = bench.make_config3_workload(1000, (4, 8), fake_reduce)
    assert calls == [400_000] and east.size == north.size == data.size == weights.size == 1000 and fe.size == 7
    rng = np.random.RandomState(0)
    np.testing.assert_array_equal(east, rng.uniform(0, 5000, 400_000)[:1000])
    np.testing.assert_array_equal(weights, np.random.RandomState(1).uniform(0.1, 10, 400_000)[:1000])
    assert ge.size == 8 and gn.size == 4
